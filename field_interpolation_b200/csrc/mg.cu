// Geometric multigrid preconditioner for the lattice normal equations, and the CG driver that uses it.
//
// Why: A^T A = S + P is a fourth-order operator for the default model_2 smoothness; Jacobi-preconditioned CG
// needs O(n^2) iterations on an n^D lattice (43,000 at 256^3 to reach 1e-6, measured).  The reference fights this
// with a coarse-to-fine initial guess (src/sdf_field.cpp:251-304): solve the same problem re-assembled on a
// coarser lattice and upscale.  A V-cycle applies that idea to the *error* on every level, every iteration:
//   * levels: sizes halved (ceil) until the lattice is tiny;
//   * coarse operators by re-discretisation, exactly the way the reference builds its coarse problem — the same
//     points in the coarse lattice's coordinates (positions * (n_c - 1) / (n_f - 1)) and the same smoothness
//     model — with the weights rescaled so that the coarse energy approximates the Galerkin energy P^T A P:
//     value rows unchanged, gradient rows / 2 per level, order-k smoothness rows * 2^((D - 2k) / 2) per level;
//   * prolongation = upscale_field's multilinear, align-corners interpolation (field_interpolation.cpp:431-485),
//     restriction = its transpose;
//   * smoother: Chebyshev polynomial in D^-1 A over [lambda_max / ratio, lambda_max] (symmetric, no dots);
//   * coarsest level: dense inverse, computed once on the host.
// The V-cycle runs in fp32 whatever the arithmetic of the outer CG (a fixed linear operator, symmetric up to
// rounding), and is captured in one CUDA graph.
#include <algorithm>
#include <cmath>
#include <vector>

#include "solver.hpp"

namespace fi {

namespace {

constexpr int kThreads = 256;

int vgrid(int64_t n) { return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + kThreads - 1) / kThreads, static_cast<int64_t>(sm_count()) * 8))); }

// ---- level construction ------------------------------------------------------------------------------------------
__global__ void coarsen_points_kernel(int D, int64_t n, const float* __restrict__ pos, const float* __restrict__ gw, float sx, float sy, float sz,
                                      float gscale, float* __restrict__ opos, float* __restrict__ ogw)
{
	const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	for (int d = 0; d < D; ++d) { opos[i * D + d] = pos[i * D + d] * (d == 0 ? sx : (d == 1 ? sy : sz)); }
	ogw[i] = gw[i] * gscale;
}

// ---- smoother ----------------------------------------------------------------------------------------------------
// One Chebyshev step: res -= q (when q is given), d = a d + b M^-1 res, e += d.
__global__ void __launch_bounds__(kThreads) cheb_step_kernel(int64_t n, float* __restrict__ res, const float* __restrict__ q, float* __restrict__ d,
                                                             const float* __restrict__ minv, float* __restrict__ e, float a, float b, int e_is_zero)
{
	const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
		float r = res[i];
		if (q) {
			r -= q[i];
			res[i] = r;
		}
		const float dn = (a != 0.0f ? a * d[i] : 0.0f) + b * minv[i] * r;
		d[i]           = dn;
		e[i]           = e_is_zero ? dn : e[i] + dn;
	}
}

// res = r - q
__global__ void __launch_bounds__(kThreads) residual_sub_kernel(int64_t n, const float* __restrict__ r, const float* __restrict__ q, float* __restrict__ res)
{
	const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) { res[i] = r[i] - q[i]; }
}

// ---- transfers -----------------------------------------------------------------------------------------------------
struct Xfer  // fine <-> coarse geometry of one level pair
{
	int    D;
	int    nf[kMaxDim], nc[kMaxDim];
	double s[kMaxDim];  // coarse position of fine node i = i * s  (upscale_field: coord * (small - 1) / (large - 1))
};

// e_f += P e_c : every fine node gathers its 2^D coarse neighbours.
__global__ void __launch_bounds__(kThreads) prolong_add_kernel(Xfer x, int64_t nfine, const float* __restrict__ ec, float* __restrict__ ef)
{
	const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (i >= nfine) { return; }
	int64_t rem = i;
	int     base[kMaxDim] = {0, 0, 0};
	float   fr[kMaxDim]   = {0, 0, 0};
	for (int d = 0; d < x.D; ++d) {
		const int c = static_cast<int>(rem % x.nf[d]);
		rem /= x.nf[d];
		const double t = c * x.s[d];
		int          b = static_cast<int>(t);
		if (b > x.nc[d] - 1) { b = x.nc[d] - 1; }
		base[d] = b;
		fr[d]   = static_cast<float>(t - b);
	}
	float acc = 0.0f;
	for (int corner = 0; corner < (1 << x.D); ++corner) {
		float   w   = 1.0f;
		int64_t idx = 0, str = 1;
		bool    ok  = true;
		for (int d = 0; d < x.D; ++d) {
			const int bit = (corner >> d) & 1;
			const int c   = base[d] + bit;
			w *= bit ? fr[d] : 1.0f - fr[d];
			ok = ok && c < x.nc[d];
			idx += str * c;
			str *= x.nc[d];
		}
		if (ok && w != 0.0f) { acc += w * ec[idx]; }
	}
	ef[i] += acc;
}

// r_c = P^T res_f : every coarse node gathers the fine nodes whose interpolation stencil contains it, with exactly
// the weights prolong_add_kernel uses (same double-precision position arithmetic), so that R = P^T.
__global__ void __launch_bounds__(kThreads) restrict_kernel(Xfer x, int64_t ncoarse, const float* __restrict__ rf, float* __restrict__ rc)
{
	const int64_t I = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (I >= ncoarse) { return; }
	int64_t rem = I;
	int     cnt[kMaxDim] = {1, 1, 1};
	int     idx[kMaxDim][6];
	float   w[kMaxDim][6];
	for (int d = 0; d < kMaxDim; ++d) {
		idx[d][0] = 0;
		w[d][0]   = 1.0f;
	}
	for (int d = 0; d < x.D; ++d) {
		const int C = static_cast<int>(rem % x.nc[d]);
		rem /= x.nc[d];
		int first = 0, last = x.nf[d] - 1;
		if (x.s[d] > 0.0) {  // fine nodes i with |i * s - C| < 1, plus a node of slack either side
			first = max(first, static_cast<int>(floor((C - 1) / x.s[d])) - 1);
			last  = min(last, static_cast<int>(ceil((C + 1) / x.s[d])) + 1);
		}
		int n = 0;
		for (int i = first; i <= last && n < 6; ++i) {
			const double t = i * x.s[d];
			int          b = static_cast<int>(t);
			if (b > x.nc[d] - 1) { b = x.nc[d] - 1; }
			const float fr = static_cast<float>(t - b);
			float       wt = 0.0f;
			if (b == C) { wt = 1.0f - fr; } else if (b + 1 == C) { wt = fr; }
			if (wt != 0.0f) {
				idx[d][n] = i;
				w[d][n]   = wt;
				++n;
			}
		}
		cnt[d] = n;
	}
	float         acc = 0.0f;
	const int64_t sy = x.nf[0], sz = static_cast<int64_t>(x.nf[0]) * x.nf[1];
	for (int k = 0; k < cnt[2]; ++k) {
		for (int j = 0; j < cnt[1]; ++j) {
			const float   wyz = w[2][k] * w[1][j];
			const int64_t row = idx[2][k] * sz + idx[1][j] * sy;
			for (int i = 0; i < cnt[0]; ++i) { acc += wyz * w[0][i] * rf[row + idx[0][i]]; }
		}
	}
	rc[I] = acc;
}

// ---- coarsest level: e = Ainv r (dense, one block per row) --------------------------------------------------------
__global__ void dense_matvec_kernel(int n, const float* __restrict__ A, const float* __restrict__ r, float* __restrict__ e)
{
	__shared__ double red[32];
	const int row = blockIdx.x;
	double    acc = 0.0;
	for (int j = threadIdx.x; j < n; j += blockDim.x) { acc += static_cast<double>(A[static_cast<size_t>(row) * n + j]) * static_cast<double>(r[j]); }
	acc = block_sum(acc, red);
	if (threadIdx.x == 0) { e[row] = static_cast<float>(acc); }
}

__global__ void unit_vector_kernel(int n, int k, float* v)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { v[i] = i == k ? 1.0f : 0.0f; }
}

// ---- power iteration helpers ----------------------------------------------------------------------------------------
__global__ void hash_fill_kernel(int64_t n, float* v)
{
	const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
		uint64_t h = static_cast<uint64_t>(i) * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull;
		h ^= h >> 31;
		h *= 0xBF58476D1CE4E5B9ull;
		h ^= h >> 29;
		v[i] = static_cast<float>(static_cast<double>(h >> 11) * (1.0 / 9007199254740992.0)) - 0.5f;
	}
}

// w = M^-1 q; out[0] = w.w, out[1] = v.w
__global__ void __launch_bounds__(kThreads) power_step_kernel(int64_t n, const float* __restrict__ v, const float* __restrict__ q,
                                                              const float* __restrict__ minv, float* __restrict__ w, double* out, double* partial,
                                                              unsigned* ticket)
{
	__shared__ double red[32];
	double            acc[2] = {0, 0};
	const int64_t     stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
		const float wi = minv[i] * q[i];
		w[i]           = wi;
		acc[0] += static_cast<double>(wi) * wi;
		acc[1] += static_cast<double>(v[i]) * wi;
	}
	acc[0] = block_sum(acc[0], red);
	acc[1] = block_sum(acc[1], red);
	grid_sum<2>(acc, partial, ticket, red, [&](const double(&tot)[2]) {
		out[0] = tot[0];
		out[1] = tot[1];
	});
}

__global__ void scale_kernel(int64_t n, float* v, float a)
{
	const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) { v[i] *= a; }
}

// ---- outer CG kernels -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) mg_residual_kernel(int64_t n, const T* __restrict__ b, const T* __restrict__ q, T* __restrict__ r, float* __restrict__ r32,
                                                               double* out, double* partial, unsigned* ticket)
{
	__shared__ double red[32];
	double            acc[2] = {0, 0};
	const int64_t     stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
		const T bi = b[i], ri = bi - q[i];
		r[i]   = ri;
		r32[i] = static_cast<float>(ri);
		acc[0] += static_cast<double>(ri) * static_cast<double>(ri);
		acc[1] += static_cast<double>(bi) * static_cast<double>(bi);
	}
	acc[0] = block_sum(acc[0], red);
	acc[1] = block_sum(acc[1], red);
	grid_sum<2>(acc, partial, ticket, red, [&](const double(&tot)[2]) {
		out[0] = tot[0];
		out[1] = tot[1];
	});
}

// x += alpha p, r -= alpha q, r32 = float(r); out[0] = r.r
template <typename T>
__global__ void __launch_bounds__(kThreads) mg_update_kernel(int64_t n, T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p, const T* __restrict__ q,
                                                             float* __restrict__ r32, T alpha, double* out, double* partial, unsigned* ticket)
{
	__shared__ double red[32];
	double            acc[1] = {0};
	const int64_t     stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
		x[i] += alpha * p[i];
		const T ri = r[i] - alpha * q[i];
		r[i]   = ri;
		r32[i] = static_cast<float>(ri);
		acc[0] += static_cast<double>(ri) * static_cast<double>(ri);
	}
	acc[0] = block_sum(acc[0], red);
	grid_sum<1>(acc, partial, ticket, red, [&](const double(&tot)[1]) { out[0] = tot[0]; });
}

// out[0] = r.z
template <typename T>
__global__ void __launch_bounds__(kThreads) mg_dot_kernel(int64_t n, const T* __restrict__ r, const float* __restrict__ z, double* out, double* partial,
                                                          unsigned* ticket)
{
	__shared__ double red[32];
	double            acc[1] = {0};
	const int64_t     stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
		acc[0] += static_cast<double>(r[i]) * static_cast<double>(z[i]);
	}
	acc[0] = block_sum(acc[0], red);
	grid_sum<1>(acc, partial, ticket, red, [&](const double(&tot)[1]) { out[0] = tot[0]; });
}

// p = z + beta p
template <typename T>
__global__ void __launch_bounds__(kThreads) mg_direction_kernel(int64_t n, const float* __restrict__ z, T* __restrict__ p, T beta)
{
	const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) { p[i] = static_cast<T>(z[i]) + beta * p[i]; }
}

void invert_spd(std::vector<double>& a, int n)  // in place, Gauss-Jordan with partial pivoting (n <= ~1000, setup only)
{
	std::vector<double> inv(static_cast<size_t>(n) * n, 0.0);
	for (int i = 0; i < n; ++i) { inv[static_cast<size_t>(i) * n + i] = 1.0; }
	for (int c = 0; c < n; ++c) {
		int    piv  = c;
		double best = std::fabs(a[static_cast<size_t>(c) * n + c]);
		for (int r = c + 1; r < n; ++r) {
			if (std::fabs(a[static_cast<size_t>(r) * n + c]) > best) {
				best = std::fabs(a[static_cast<size_t>(r) * n + c]);
				piv  = r;
			}
		}
		FI_REQUIRE(best > 0.0, FI_ERR_INVALID, "multigrid: the coarsest operator is singular");
		if (piv != c) {
			for (int k = 0; k < n; ++k) {
				std::swap(a[static_cast<size_t>(piv) * n + k], a[static_cast<size_t>(c) * n + k]);
				std::swap(inv[static_cast<size_t>(piv) * n + k], inv[static_cast<size_t>(c) * n + k]);
			}
		}
		const double d = 1.0 / a[static_cast<size_t>(c) * n + c];
		for (int k = 0; k < n; ++k) {
			a[static_cast<size_t>(c) * n + k] *= d;
			inv[static_cast<size_t>(c) * n + k] *= d;
		}
		for (int r = 0; r < n; ++r) {
			if (r == c) { continue; }
			const double f = a[static_cast<size_t>(r) * n + c];
			if (f == 0.0) { continue; }
			for (int k = 0; k < n; ++k) {
				a[static_cast<size_t>(r) * n + k] -= f * a[static_cast<size_t>(c) * n + k];
				inv[static_cast<size_t>(r) * n + k] -= f * inv[static_cast<size_t>(c) * n + k];
			}
		}
	}
	a.swap(inv);
}

}  // namespace

struct Multigrid::Level
{
	Geom                             g;
	Operator<float>*                 op = nullptr;   // level 0: the caller's operator; coarser: owned below
	std::unique_ptr<Operator<float>> owned;
	PointStore                       pts;
	DevBuf<float>                    r, e, res, d, q;  // r / e of level 0 are the caller's vectors
	double                           lmax = 0;
	Xfer                             to_coarser;     // this level (fine) -> next level (coarse)
};

Multigrid::Multigrid()  = default;
Multigrid::~Multigrid()
{
	if (exec) { cudaGraphExecDestroy(exec); }
}

std::unique_ptr<Multigrid> build_multigrid(Operator<float>& fine, const ModelAccum& model, const PointStore& pts, const MgOptions& opt, cudaStream_t s)
{
	TraceScope trace("build_multigrid");
	FI_REQUIRE(!fine.g.sharded(), FI_ERR_UNSUPPORTED, "the multigrid preconditioner runs on one GPU");
	auto mg     = std::make_unique<Multigrid>();
	mg->opt     = opt;
	const int D = fine.g.ndim;
	// levels
	{
		auto l0 = std::make_unique<Multigrid::Level>();
		l0->g   = fine.g;
		l0->op  = &fine;
		mg->levels.push_back(std::move(l0));
	}
	while (static_cast<int>(mg->levels.size()) < 20) {
		const Geom& gf = mg->levels.back()->g;
		int32_t     nc[kMaxDim] = {1, 1, 1};
		int64_t     cells = 1;
		bool        shrunk = false;
		int         smallest = 1 << 30;
		for (int d = 0; d < D; ++d) {
			nc[d] = (gf.size[d] + 1) / 2;
			shrunk = shrunk || nc[d] < gf.size[d];
			cells *= nc[d];
			if (gf.size[d] > 1) { smallest = std::min(smallest, gf.size[d]); }
		}
		// stop once the current level is small enough for the dense solve
		if (gf.N <= opt.coarsest_cells || !shrunk || smallest <= 2) { break; }
		auto lv = std::make_unique<Multigrid::Level>();
		lv->g   = make_geom(D, nc);
		mg->levels.push_back(std::move(lv));
	}
	const int L = static_cast<int>(mg->levels.size());
	FI_REQUIRE(mg->levels.back()->g.N <= 4096, FI_ERR_UNSUPPORTED, "multigrid: coarsest level too large for the dense solve");
	// coarse operators by re-discretisation
	for (int l = 1; l < L; ++l) {
		Multigrid::Level& lv = *mg->levels[l];
		ModelAccum        m  = model;
		for (int k = 0; k <= 4; ++k) {
			const double f = std::pow(2.0, static_cast<double>(D - 2 * k) * l);  // squared weights scale by rho^(D-2k) per level
			for (int a = 0; a < 5; ++a) {
				for (int b = 0; b < 5; ++b) { m.cc[k][a][b] *= f; }
			}
		}
		m.gs_sq *= std::pow(2.0, static_cast<double>(D - 4) * l);
		const int64_t n = pts.count;
		if (n > 0) {
			lv.pts.pos.resize(static_cast<size_t>(n) * D);
			lv.pts.grad.resize(static_cast<size_t>(n) * D);
			lv.pts.value.resize(n);
			lv.pts.vw.resize(n);
			lv.pts.gw.resize(n);
			lv.pts.kind.resize(n);
			lv.pts.count = n;
			float sc[kMaxDim] = {1, 1, 1};
			for (int d = 0; d < D; ++d) {
				sc[d] = fine.g.size[d] > 1 ? static_cast<float>(static_cast<double>(lv.g.size[d] - 1) / static_cast<double>(fine.g.size[d] - 1)) : 0.0f;
			}
			FI_LAUNCH(coarsen_points_kernel, div_up(n, kThreads), kThreads, 0, s, D, n, pts.pos.data(), pts.gw.data(), sc[0], sc[1], sc[2],
			          static_cast<float>(std::pow(0.5, l)), lv.pts.pos.data(), lv.pts.gw.data());
			FI_CUDA(cudaMemcpyAsync(lv.pts.grad.data(), pts.grad.data(), static_cast<size_t>(n) * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
			FI_CUDA(cudaMemcpyAsync(lv.pts.value.data(), pts.value.data(), n * sizeof(float), cudaMemcpyDeviceToDevice, s));
			FI_CUDA(cudaMemcpyAsync(lv.pts.vw.data(), pts.vw.data(), n * sizeof(float), cudaMemcpyDeviceToDevice, s));
			FI_CUDA(cudaMemcpyAsync(lv.pts.kind.data(), pts.kind.data(), n * sizeof(uint8_t), cudaMemcpyDeviceToDevice, s));
		}
		HostRows none;
		lv.owned = build_operator<float>(lv.g, m, lv.pts, none, s);
		lv.op    = lv.owned.get();
		lv.op->use_fast = kStencilAuto;
	}
	// work vectors and transfer geometry
	for (int l = 0; l < L; ++l) {
		Multigrid::Level& lv = *mg->levels[l];
		const size_t      n  = static_cast<size_t>(lv.g.N);
		if (l > 0) {
			lv.r.resize(n);
			lv.e.resize(n);
		}
		if (l < L - 1) {
			lv.res.resize(n);
			lv.d.resize(n);
			lv.q.resize(n);
			Xfer& x = lv.to_coarser;
			x.D     = D;
			for (int d = 0; d < kMaxDim; ++d) {
				x.nf[d] = lv.g.size[d];
				x.nc[d] = mg->levels[l + 1]->g.size[d];
				x.s[d]  = x.nf[d] > 1 ? static_cast<double>(x.nc[d] - 1) / static_cast<double>(x.nf[d] - 1) : 0.0;
			}
		}
	}
	// largest eigenvalue of D^-1 A per smoothed level: power iteration
	{
		DevBuf<double>   out(2), partial(static_cast<size_t>(2) * (static_cast<size_t>(sm_count()) * 8 + 8));
		DevBuf<unsigned> ticket(1);
		ticket.zero(s);
		for (int l = 0; l < L - 1; ++l) {
			Multigrid::Level& lv = *mg->levels[l];
			const int64_t     n  = lv.g.N;
			float *           v = lv.d.data(), *w = lv.res.data();
			FI_LAUNCH(hash_fill_kernel, vgrid(n), kThreads, 0, s, n, v);
			double lam = 1.0;
			for (int it = 0; it < opt.power_iterations; ++it) {
				lv.op->apply(v, lv.q.data(), nullptr, nullptr, s);
				FI_LAUNCH(power_step_kernel, vgrid(n), kThreads, 0, s, n, v, lv.q.data(), lv.op->minv.data(), w, out.data(), partial.data(), ticket.data());
				double h[2];
				FI_CUDA(cudaMemcpyAsync(h, out.data(), sizeof(h), cudaMemcpyDeviceToHost, s));
				FI_CUDA(cudaStreamSynchronize(s));
				const double nw = std::sqrt(h[0]);
				FI_REQUIRE(nw > 0 && std::isfinite(nw), FI_ERR_INVALID, "multigrid: power iteration broke down");
				lam = nw;  // |D^-1 A v| with |v| = 1 (after the first round)
				FI_LAUNCH(scale_kernel, vgrid(n), kThreads, 0, s, n, w, static_cast<float>(1.0 / nw));
				std::swap(v, w);
			}
			lv.lmax = lam * 1.1;  // the power iteration approaches lambda_max from below
		}
	}
	// dense inverse of the coarsest operator
	{
		Multigrid::Level& lc = *mg->levels[L - 1];
		const int         n  = static_cast<int>(lc.g.N);
		mg->nc               = n;
		if (L == 1) {
			lc.r.resize(n);  // scratch for the unit vectors (level 0 owns no vectors otherwise)
			lc.e.resize(n);
		}
		std::vector<double> A(static_cast<size_t>(n) * n);
		std::vector<float>  col(n);
		for (int k = 0; k < n; ++k) {
			FI_LAUNCH(unit_vector_kernel, div_up(n, kThreads), kThreads, 0, s, n, k, lc.r.data());
			lc.op->apply(lc.r.data(), lc.e.data(), nullptr, nullptr, s);
			FI_CUDA(cudaMemcpyAsync(col.data(), lc.e.data(), n * sizeof(float), cudaMemcpyDeviceToHost, s));
			FI_CUDA(cudaStreamSynchronize(s));
			for (int i = 0; i < n; ++i) { A[static_cast<size_t>(i) * n + k] = col[i]; }
		}
		for (int i = 0; i < n; ++i) {  // symmetrise (the operator is symmetric up to fp32 rounding of the atomics)
			for (int j = i + 1; j < n; ++j) {
				const double v = 0.5 * (A[static_cast<size_t>(i) * n + j] + A[static_cast<size_t>(j) * n + i]);
				A[static_cast<size_t>(i) * n + j] = A[static_cast<size_t>(j) * n + i] = v;
			}
		}
		invert_spd(A, n);
		std::vector<float> Af(A.size());
		for (size_t i = 0; i < A.size(); ++i) { Af[i] = static_cast<float>(A[i]); }
		mg->coarse_inv.resize(Af.size());
		FI_CUDA(cudaMemcpyAsync(mg->coarse_inv.data(), Af.data(), Af.size() * sizeof(float), cudaMemcpyHostToDevice, s));
		FI_CUDA(cudaStreamSynchronize(s));
	}
	return mg;
}

namespace {

// nu Chebyshev steps on A e = r at one level.  e_zero: e starts at zero (pre-smoothing).  On return lv.res holds the
// residual *before* the last correction d (so r - A e = res - A d).
void smooth(Multigrid::Level& lv, const MgOptions& opt, const float* r, float* e, bool e_zero, cudaStream_t s)
{
	const int64_t n     = lv.g.N;
	const double  lmax  = lv.lmax, lmin = lmax / opt.cheb_ratio;
	const double  theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
	if (e_zero) {
		FI_CUDA(cudaMemcpyAsync(lv.res.data(), r, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
	} else {
		lv.op->apply(e, lv.q.data(), nullptr, nullptr, s);
		FI_LAUNCH(residual_sub_kernel, vgrid(n), kThreads, 0, s, n, r, lv.q.data(), lv.res.data());
	}
	double rho = 1.0 / sigma;
	FI_LAUNCH(cheb_step_kernel, vgrid(n), kThreads, 0, s, n, lv.res.data(), static_cast<const float*>(nullptr), lv.d.data(), lv.op->minv.data(), e, 0.0f,
	          static_cast<float>(1.0 / theta), e_zero ? 1 : 0);
	for (int k = 1; k < opt.nu; ++k) {
		const double rho_new = 1.0 / (2.0 * sigma - rho);
		lv.op->apply(lv.d.data(), lv.q.data(), nullptr, nullptr, s);
		FI_LAUNCH(cheb_step_kernel, vgrid(n), kThreads, 0, s, n, lv.res.data(), static_cast<const float*>(lv.q.data()), lv.d.data(), lv.op->minv.data(), e,
		          static_cast<float>(rho_new * rho), static_cast<float>(2.0 * rho_new / delta), 0);
		rho = rho_new;
	}
}

void vcycle_level(Multigrid& mg, int l, const float* r, float* e, cudaStream_t s)
{
	const int L = static_cast<int>(mg.levels.size());
	Multigrid::Level& lv = *mg.levels[l];
	if (l == L - 1) {
		FI_LAUNCH(dense_matvec_kernel, mg.nc, 128, 0, s, mg.nc, mg.coarse_inv.data(), r, e);
		return;
	}
	Multigrid::Level& lc = *mg.levels[l + 1];
	smooth(lv, mg.opt, r, e, true, s);
	// residual after the last correction, restricted
	lv.op->apply(lv.d.data(), lv.q.data(), nullptr, nullptr, s);
	FI_LAUNCH(residual_sub_kernel, vgrid(lv.g.N), kThreads, 0, s, lv.g.N, lv.res.data(), lv.q.data(), lv.res.data());
	FI_LAUNCH(restrict_kernel, div_up(lc.g.N, kThreads), kThreads, 0, s, lv.to_coarser, lc.g.N, lv.res.data(), lc.r.data());
	vcycle_level(mg, l + 1, lc.r.data(), lc.e.data(), s);
	FI_LAUNCH(prolong_add_kernel, div_up(lv.g.N, kThreads), kThreads, 0, s, lv.to_coarser, lv.g.N, lc.e.data(), e);
	smooth(lv, mg.opt, r, e, false, s);
}

}  // namespace

void Multigrid::vcycle(const float* r, float* z, cudaStream_t s)
{
	if (exec && graph_r == r && graph_z == z) {
		FI_CUDA(cudaGraphLaunch(exec, s));
		count_launch(static_cast<int>(graph_launches));
		return;
	}
	if (exec) {
		cudaGraphExecDestroy(exec);
		exec = nullptr;
	}
	cudaGraph_t   graph  = nullptr;
	const int64_t before = g_launches;
	FI_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
	try {
		vcycle_level(*this, 0, r, z, s);
	} catch (...) {
		cudaStreamEndCapture(s, &graph);
		if (graph) { cudaGraphDestroy(graph); }
		throw;
	}
	FI_CUDA(cudaStreamEndCapture(s, &graph));
	graph_launches = g_launches - before;
	g_launches     = before;
	FI_CUDA(cudaGraphInstantiate(&exec, graph, 0));
	cudaGraphDestroy(graph);
	graph_r = r;
	graph_z = z;
	FI_CUDA(cudaGraphLaunch(exec, s));
	count_launch(static_cast<int>(graph_launches));
}

// CG on A x = b preconditioned by one V-cycle per iteration.  Scalars travel through the host (three small reads
// per iteration against several milliseconds of device work).  Same stopping rule as pcg_solve.
template <typename T>
PcgResult mgpcg_solve(Operator<T>& op, Multigrid& mg, const T* b, T* x, double tol, long long max_iter, cudaStream_t s)
{
	TraceScope    trace("mgpcg_solve");
	const int64_t n   = op.g.N;
	const T*      rhs = b ? b : op.atb.data();
	if (max_iter <= 0) { max_iter = 2 * n; }
	if (!(tol > 0)) { tol = std::is_same<T, float>::value ? 1.1920929e-07 : 2.220446049250313e-16; }
	DevBuf<T>        r(n), p(n), q(n);
	DevBuf<float>    r32(n), z(n);
	DevBuf<double>   out(2), partial(static_cast<size_t>(2) * (static_cast<size_t>(sm_count()) * 8 + 8));
	DevBuf<unsigned> ticket(1);
	ticket.zero(s);
	cudaEvent_t e0, e1;
	FI_CUDA(cudaEventCreate(&e0));
	FI_CUDA(cudaEventCreate(&e1));
	FI_CUDA(cudaEventRecord(e0, s));
	auto read2 = [&](double* h) {
		FI_CUDA(cudaMemcpyAsync(h, out.data(), 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
		FI_CUDA(cudaStreamSynchronize(s));
	};
	const int grid = vgrid(n);
	PcgResult res;
	double    h[2] = {0, 0};
	op.apply(x, q.data(), nullptr, nullptr, s);
	{
		auto k = mg_residual_kernel<T>;
		FI_LAUNCH(k, grid, kThreads, 0, s, n, rhs, q.data(), r.data(), r32.data(), out.data(), partial.data(), ticket.data());
	}
	read2(h);
	double       rr = h[0];
	const double bb = h[1];
	res.zero_rhs         = bb == 0.0;
	res.initial_residual = bb > 0 ? std::sqrt(rr / bb) : 0.0;
	const double target  = tol * tol * bb;
	long long    it      = 0;
	if (res.zero_rhs) {
		FI_CUDA(cudaMemsetAsync(x, 0, n * sizeof(T), s));
		rr = 0;
	} else if (rr > target) {
		double rz = 0;
		p.zero(s);
		while (it < max_iter) {
			mg.vcycle(r32.data(), z.data(), s);
			{
				auto k = mg_dot_kernel<T>;
				FI_LAUNCH(k, grid, kThreads, 0, s, n, r.data(), z.data(), out.data(), partial.data(), ticket.data());
			}
			read2(h);
			const double rz_new = h[0];
			if (!(rz_new > 0.0) || !std::isfinite(rz_new)) { break; }  // breakdown: keep the last iterate
			const double beta = it == 0 ? 0.0 : rz_new / rz;
			rz                = rz_new;
			{
				auto k = mg_direction_kernel<T>;
				FI_LAUNCH(k, grid, kThreads, 0, s, n, z.data(), p.data(), static_cast<T>(beta));
			}
			op.apply(p.data(), q.data(), out.data(), nullptr, s);
			read2(h);
			const double pq = h[0];
			if (!(pq > 0.0) || !std::isfinite(pq)) { break; }
			const double alpha = rz / pq;
			{
				auto k = mg_update_kernel<T>;
				FI_LAUNCH(k, grid, kThreads, 0, s, n, x, r.data(), p.data(), q.data(), r32.data(), static_cast<T>(alpha), out.data(), partial.data(),
				          ticket.data());
			}
			read2(h);
			rr = h[0];
			++it;
			if (rr <= target) { break; }
		}
	}
	FI_CUDA(cudaEventRecord(e1, s));
	FI_CUDA(cudaEventSynchronize(e1));
	float ms = 0;
	FI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	res.solve_ms      = ms;
	res.loop_ms       = ms;
	res.iterations    = it;
	res.rel_residual  = bb > 0 ? std::sqrt(rr / bb) : 0.0;
	res.converged     = res.zero_rhs || rr <= target;
	res.true_residual = res.rel_residual;
	if (!res.zero_rhs) {
		double trr = 0, tbb = 0;
		residual<T>(op, rhs, x, nullptr, &trr, &tbb, s);
		res.true_residual = tbb > 0 ? std::sqrt(trr / tbb) : 0.0;
	}
	return res;
}

template PcgResult mgpcg_solve<float>(Operator<float>&, Multigrid&, const float*, float*, double, long long, cudaStream_t);
template PcgResult mgpcg_solve<double>(Operator<double>&, Multigrid&, const double*, double*, double, long long, cudaStream_t);

}  // namespace fi
