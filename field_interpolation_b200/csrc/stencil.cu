// Matrix-free smoothness operator S = (model rows)^T (model rows)  — north-star item (b).
//
// The reference materialises one row per lattice node, axis and order (add_model_constraint,
// field_interpolation/field_interpolation.cpp:243-316, driven by add_field_constraints :326-341) and lets
// Eigen form A^T A (make_square, sparse_linear.cpp:105-113).  Those rows are constant-coefficient forward
// differences, so their normal matrix is, per axis d,
//     T_d = sum_k w_k^2 D_k^T D_k        D_k = k-th forward difference restricted to rows coord+k < size
// (a 9-diagonal banded matrix that is constant away from the two ends), plus the gradient-smoothness cross
// terms 2 w_gs^2 sum_{d<o} (D_1^T D_1)_d (x) (D_1^T D_1)_o  (:303-315 emits every pair twice).
// S p is applied directly from p with per-axis coefficient rows selected by a 9-way "row class"
// (4 classes at each end + interior); nothing is stored per node.
#include <algorithm>

#include "internal.hpp"

namespace fi {

namespace {

constexpr int kThreads = 256;

__host__ __device__ __forceinline__ int row_class(int i, int n)
{
	return n <= 9 ? i : (i < 4 ? i : (i >= n - 4 ? i - n + 9 : 4));
}

template <typename T>
struct DevTables
{
	T   band[kMaxDim][9][9];
	T   gs2;
	int radius;
	bool any;
};

template <typename T>
DevTables<T> to_dev(const StencilTables& t)
{
	DevTables<T> d;
	for (int a = 0; a < kMaxDim; ++a) {
		for (int c = 0; c < 9; ++c) {
			for (int k = 0; k < 9; ++k) { d.band[a][c][k] = static_cast<T>(t.band[a][c][k]); }
		}
	}
	d.gs2    = static_cast<T>(t.gs2);
	d.radius = t.radius;
	d.any    = t.any;
	return d;
}

// (D_1^T D_1)[i][i+o] for the restricted first difference on an axis of n nodes.
__device__ __forceinline__ int lap1(int i, int n, int o)
{
	if (o == 0) { return (i > 0 ? 1 : 0) + (i < n - 1 ? 1 : 0); }
	const int j = i + o;
	return (j >= 0 && j < n) ? -1 : 0;
}

__device__ __forceinline__ void coords_of(const Geom& g, int64_t index, int* c)
{
	// lattice coordinates of local cell `index` (the slowest axis runs over the locally stored planes)
	c[0] = c[1] = c[2] = 0;
	for (int d = 0; d < g.ndim; ++d) {
		if (d == g.ndim - 1) {
			c[d] = static_cast<int>(index) + (d == 2 ? g.zoff : 0);
		} else {
			c[d] = static_cast<int>(index % g.size[d]);
			index /= g.size[d];
		}
	}
}

// diag += diagonal of S.  One block per (x segment, kDiagRows y rows, local z plane): the coordinates come from the block
// index — the div / mod chain per node of the first version made this 8 B/cell kernel run at a tenth of the memory roofline.
constexpr int kDiagRows = 8;

template <typename T>
__global__ void __launch_bounds__(kThreads) diagonal_kernel(Geom g, DevTables<T> tab, T* __restrict__ diag, T* __restrict__ minv)
{
	const int nx = g.size[0], ny = g.ndim >= 2 ? g.size[1] : 1;
	const int zl = blockIdx.z;                      // local plane (3D), 0 otherwise
	const int z  = g.ndim == 3 ? zl + g.zoff : 0;   // lattice plane
	const bool outside = g.ndim == 3 && (z < 0 || z >= g.size[2]);  // slab planes beyond the lattice: the diagonal stays zero
	T   zpart = 0;
	int lz    = 0;
	if (g.ndim == 3 && !outside) {
		zpart = tab.band[2][row_class(z, g.size[2])][4];
		lz    = lap1(z, g.size[2], 0);
	}
	for (int r = 0; r < kDiagRows; ++r) {
		const int y = blockIdx.y * kDiagRows + r;
		if (y >= ny) { break; }
		T   rest = zpart;  // what does not depend on x
		int ly   = 0;
		if (g.ndim >= 2) {
			rest += tab.band[1][row_class(y, ny)][4];
			ly = lap1(y, ny, 0);
			if (tab.gs2 != T(0)) { rest += tab.gs2 * static_cast<T>(ly * lz); }
		}
		const int64_t at = (static_cast<int64_t>(zl) * ny + y) * nx;
		for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nx; x += gridDim.x * blockDim.x) {
			T d = diag[at + x];
			if (!outside && tab.any) {
				T acc = rest + tab.band[0][row_class(x, nx)][4];
				if (tab.gs2 != T(0) && g.ndim >= 2) { acc += tab.gs2 * static_cast<T>(lap1(x, nx, 0) * (ly + lz)); }
				d += acc;
				diag[at + x] = d;
			}
			// Eigen's DiagonalPreconditioner: 1 / diag, 1 where the diagonal is zero
			if (minv) { minv[at + x] = d != T(0) ? T(1) / d : T(1); }
		}
	}
}

// Reference-shaped kernel for every dimension, order and weight combination: one thread per node, neighbours
// read through L1/L2.  Also the parity baseline of the specialised kernels.
template <typename T>
__global__ void __launch_bounds__(kThreads) stencil_generic_kernel(Geom g, DevTables<T> tab, const T* __restrict__ p, T* __restrict__ q,
                                                                   double* dot_out, double* partial, unsigned* ticket, const int* done)
{
	__shared__ T      band[kMaxDim][9][9];
	__shared__ double red[32];
	if (done && *done) { return; }
	for (int k = threadIdx.x; k < kMaxDim * 81; k += blockDim.x) { (&band[0][0][0])[k] = (&tab.band[0][0][0])[k]; }
	__syncthreads();
	const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	double        mine[1] = {0.0};
	if (i < g.N) {
		int c[kMaxDim];
		coords_of(g, i, c);
		T         acc = 0;
		const int R   = tab.radius;
		const int tl  = g.tile;  // >= 2: block-Jacobi mask (taps into another tile are dropped)
		for (int d = 0; d < g.ndim; ++d) {
			const T* row = band[d][row_class(c[d], g.size[d])];
			for (int o = -R; o <= R; ++o) {
				const T coef = row[o + 4];
				if (coef == T(0)) { continue; }
				if (tl && (c[d] + o < 0 || (c[d] + o) / tl != c[d] / tl)) { continue; }
				acc += coef * p[i + o * g.stride[d]];
			}
		}
		if (tl) { acc += static_cast<T>(g.tile_reg) * p[i]; }
		if (tab.gs2 != T(0)) {
			for (int d = 0; d < g.ndim; ++d) {
				for (int o = d + 1; o < g.ndim; ++o) {
					for (int a = -1; a <= 1; ++a) {
						const int la = lap1(c[d], g.size[d], a);
						if (la == 0) { continue; }
						for (int b = -1; b <= 1; ++b) {
							const int lb = lap1(c[o], g.size[o], b);
							if (lb == 0) { continue; }
							if (tl && ((c[d] + a) / tl != c[d] / tl || (c[o] + b) / tl != c[o] / tl)) { continue; }
							acc += tab.gs2 * static_cast<T>(la * lb) * p[i + a * g.stride[d] + b * g.stride[o]];
						}
					}
				}
			}
		}
		q[i]    = acc;
		mine[0] = static_cast<double>(p[i]) * static_cast<double>(acc);
	}
	if (dot_out) {
		mine[0] = block_sum(mine[0], red);
		grid_sum<1>(mine, partial, ticket, red, [&](const double(&tot)[1]) { *dot_out = tot[0]; });
	}
}

}  // namespace

StencilTables make_tables(const Geom& g, const ModelAccum& m)
{
	StencilTables t;
	std::memset(&t, 0, sizeof(t));
	t.gs2    = 2.0 * m.gs_sq;
	t.radius = 0;
	for (int k = 0; k <= 4; ++k) {
		if (m.on[k]) { t.radius = k; }
	}
	t.any = t.gs2 > 0;
	for (int k = 0; k <= 4; ++k) { t.any = t.any || m.on[k]; }
	for (int d = 0; d < g.ndim; ++d) {
		const int n = g.size[d];
		for (int cls = 0; cls < 9; ++cls) {
			int i;  // representative row of this class
			if (n <= 9) {
				if (cls >= n) { continue; }
				i = cls;
			} else {
				i = cls < 4 ? cls : (cls == 4 ? 4 : n - 9 + cls);
			}
			for (int k = 0; k <= 4; ++k) {
				if (!m.on[k]) { continue; }
				// rows j of D_k exist for 0 <= j <= n-k-1 and touch nodes j..j+k
				for (int j = std::max(0, i - k); j <= std::min(i, n - k - 1); ++j) {
					for (int mm = 0; mm <= k; ++mm) {
						const int o = j + mm - i;  // neighbour offset
						t.band[d][cls][o + 4] += m.cc[k][i - j][mm];
					}
				}
			}
		}
	}
	if (g.ndim > 1 && t.gs2 > 0 && t.radius < 1) { t.radius = 1; }
	return t;
}

template <typename T>
void stencil_diagonal(const Geom& g, const StencilTables& t, T* d_diag, T* d_minv, cudaStream_t s)
{
	if (!t.any && !d_minv) { return; }
	auto kern = diagonal_kernel<T>;
	const dim3 grid(std::min(div_up(g.size[0], kThreads), 64), g.ndim >= 2 ? div_up(g.size[1], kDiagRows) : 1, g.ndim == 3 ? g.nzl : 1);
	FI_REQUIRE(grid.y <= 65535 && grid.z <= 65535, FI_ERR_RANGE, "lattice too large along y or z for the diagonal kernel's grid");
	FI_LAUNCH(kern, grid, kThreads, 0, s, g, to_dev<T>(t), d_diag, d_minv);
}

int stencil_partial_slots(const Geom& g) { return div_up(g.N, kThreads) + 8; }

template <typename T>
bool stencil_fast_3d(const Geom& g, const StencilTables& t, const T* p, T* q, double* d_dot_out, double* d_partial,
                     unsigned* d_ticket, const int* d_done, cudaStream_t s);  // stencil_fast.cu
template <typename T>
bool stencil_tma_3d(const Geom& g, const StencilTables& t, const T* p, T* q, double* d_dot_out, double* d_partial,
                    unsigned* d_ticket, const int* d_done, cudaStream_t s);  // stencil_tma.cu
template <typename T>
bool stencil_tma_2d(const Geom& g, const StencilTables& t, const T* p, T* q, double* d_dot_out, double* d_partial,
                    unsigned* d_ticket, const int* d_done, cudaStream_t s);  // stencil_2d.cu

template <typename T>
void stencil_apply(const Geom& g, const StencilTables& t, const T* p, T* q, double* d_dot_out, double* d_partial,
                   unsigned* d_ticket, const int* d_done, int mode, cudaStream_t s)
{
	if (g.tile) { mode = kStencilGeneric; }  // only the generic kernel knows the tile mask
	if (mode == kStencilAuto && g.ndim == 2 && stencil_tma_2d<T>(g, t, p, q, d_dot_out, d_partial, d_ticket, d_done, s)) { return; }
	if (mode == kStencilAuto && stencil_tma_3d<T>(g, t, p, q, d_dot_out, d_partial, d_ticket, d_done, s)) { return; }
	if (mode != kStencilGeneric && stencil_fast_3d<T>(g, t, p, q, d_dot_out, d_partial, d_ticket, d_done, s)) { return; }
	// the generic kernel runs over every stored node and would read beyond a slab's halo planes
	FI_REQUIRE(!g.sharded(), FI_ERR_UNSUPPORTED, "a slab-sharded lattice needs the TMA-staged 3D stencil kernel (x size a multiple of 16 bytes and >= 32, y size >= 8)");
	auto kern = stencil_generic_kernel<T>;
	FI_LAUNCH(kern, div_up(g.N, kThreads), kThreads, 0, s, g, to_dev<T>(t), p, q, d_dot_out, d_partial, d_ticket, d_done);
}

template void stencil_diagonal<float>(const Geom&, const StencilTables&, float*, float*, cudaStream_t);
template void stencil_diagonal<double>(const Geom&, const StencilTables&, double*, double*, cudaStream_t);
template void stencil_apply<float>(const Geom&, const StencilTables&, const float*, float*, double*, double*, unsigned*, const int*, int, cudaStream_t);
template void stencil_apply<double>(const Geom&, const StencilTables&, const double*, double*, double*, double*, unsigned*, const int*, int, cudaStream_t);

}  // namespace fi
