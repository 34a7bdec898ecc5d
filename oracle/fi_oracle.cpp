// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Single-threaded CPU restatement ("port") of the hot path of
// emilk/field_interpolation.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this; the product (libfi_b200.so) never links or calls it.
//
// PARITY PINNING.  The reference ships no tests, golden vectors or fixtures for this path
// (SURVEY.md §4).  The ASSEMBLY half below is pinned against the reference itself: oracle/_ref
// (the reference's own field_interpolation.cpp compiled unmodified, see oracle/Makefile) is run in
// the build container, tests/test_oracle_vs_ref.py compares both bit-for-bit, and
// tests/golden/make_golden.py freezes reference outputs as fixtures that travel to the GPU box.
// The SOLVE half restates sparse_linear.cpp + Eigen 3 (un-vendored, unpinned system dependency:
// reference CMakeLists.txt:4,10 / build.sh:84-89; not installed here, no network), i.e. Eigen's
// published setFromTriplets / sparse product / DiagonalPreconditioner / BiCGSTAB algorithms.  That half
// is pinned only against scipy fp64 direct solves of the same normal equations: "parity unpinned"
// with respect to Eigen's float rounding.
//
// Every function cites the reference lines it follows (paths relative to /root/reference).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

constexpr int kMaxDim = 3;  // field_interpolation.hpp:44

struct System  // LinearEquation, sparse_linear.hpp:18-22, as struct-of-arrays
{
	std::vector<int>   row, col;
	std::vector<float> val;
	std::vector<float> rhs;

	int next_row() const { return static_cast<int>(rhs.size()); }
	void put(int r, int c, float v) { row.push_back(r); col.push_back(c); val.push_back(v); }
};

struct Lattice  // LatticeField, field_interpolation.hpp:97-114 (x fastest)
{
	System eq;
	int    ndim = 0;
	int    size[kMaxDim]   = {1, 1, 1};
	int    stride[kMaxDim] = {1, 1, 1};

	Lattice(int nd, const int* sz) : ndim(nd)
	{
		int s = 1;
		for (int d = 0; d < nd; ++d) { size[d] = sz[d]; stride[d] = s; s *= sz[d]; }
	}
	int64_t unknowns() const
	{
		int64_t n = 1;
		for (int d = 0; d < ndim; ++d) { n *= size[d]; }
		return n;
	}
};

struct OraWeights  // Weights, field_interpolation.hpp:75-95
{
	float data_pos, data_gradient, model_0, model_1, model_2, model_3, model_4, gradient_smoothness;
	int   value_kernel, gradient_kernel;
};

// sparse_linear.cpp:34-50 — the generic row appender.
void append_row(System* eq, float weight, float rhs, int n, const int* cols, const float* coef)
{
	if (weight == 0) { return; }
	const int r = eq->next_row();
	bool any = false;
	for (int i = 0; i < n; ++i) {
		if (coef[i] == 0) { continue; }
		eq->put(r, cols[i], coef[i] * weight);
		any = true;
	}
	if (any) { eq->rhs.push_back(rhs * weight); }
}

// field_interpolation.cpp:15-55 — corners of the cell containing `pos`, compacted to the ones inside.
// `margin` is the reference's extra_bound.
int corner_weights(const int* size, const int* stride, int ndim, const float* pos, int margin, int* out_index, float* out_w)
{
	int   base[kMaxDim];
	float frac[kMaxDim];
	for (int d = 0; d < ndim; ++d) {
		base[d] = static_cast<int>(std::floor(pos[d]));
		frac[d] = pos[d] - static_cast<float>(base[d]);
	}
	int kept = 0;
	for (int corner = 0; corner < (1 << ndim); ++corner) {
		int   index = 0;
		float w     = 1.0f;
		bool  ok    = true;
		for (int d = 0; d < ndim; ++d) {
			const int bit = (corner >> d) & 1;
			const int c   = base[d] + bit;
			index += stride[d] * c;
			w *= bit ? frac[d] : 1.0f - frac[d];
			ok = ok && (0 <= c) && (c + margin < size[d]);
		}
		if (ok) { out_index[kept] = index; out_w[kept] = w; ++kept; }
	}
	return kept;
}

// field_interpolation.cpp:57-80.  Zero-valued corner coefficients ARE emitted (no add_equation here).
bool value_row(Lattice* f, const float* pos, float value, float cw)
{
	if (cw == 0) { return false; }
	int   idx[8];
	float k[8];
	const int n = corner_weights(f->size, f->stride, f->ndim, pos, 0, idx, k);
	if (n == 0) { return false; }
	const int r   = f->eq.next_row();
	float     sum = 0;
	for (int i = 0; i < n; ++i) {
		const float c = k[i] * cw;
		f->eq.put(r, idx[i], c);
		sum += c;
	}
	f->eq.rhs.push_back(sum * value);
	return true;
}

// field_interpolation.cpp:82-107.
bool value_row_nearest(Lattice* f, const float* pos, const float* grad, float value, float w)
{
	int   node = 0;
	float along = 0;
	for (int d = 0; d < f->ndim; ++d) {
		const int nd = static_cast<int>(std::round(pos[d]));
		if (nd < 0 || f->size[d] <= nd) { return false; }
		along += (pos[d] - static_cast<float>(nd)) * grad[d];
		node += nd * f->stride[d];
	}
	const float one = 1.0f;
	append_row(&f->eq, w, value - along, 1, &node, &one);
	return true;
}

// field_interpolation.cpp:110-121.
int containing_cell(const Lattice& f, const float* pos)
{
	int index = 0;
	for (int d = 0; d < f.ndim; ++d) {
		const int c = static_cast<int>(std::floor(pos[d]));
		if (!(0 <= c && c + 1 < f.size[d])) { return -1; }
		index += c * f.stride[d];
	}
	return index;
}

// field_interpolation.cpp:123-240.  kernel: 0 nearest-neighbour, 1 cell edges, 2 linear interpolation.
// Returns 1/0 like the reference's bool, -1 for an unknown kernel (the reference aborts there, :238).
int gradient_rows(Lattice* f, const float* pos, const float* grad, float cw, int kernel)
{
	if (cw == 0) { return 0; }
	const int D = f->ndim;
	if (kernel == 0) {  // :134-149
		const int cell = containing_cell(*f, pos);
		if (cell < 0) { return 0; }
		for (int d = 0; d < D; ++d) {
			const int   cols[2] = {cell, cell + f->stride[d]};
			const float coef[2] = {-1.0f, +1.0f};
			append_row(&f->eq, cw, grad[d], 2, cols, coef);
		}
		return 1;
	}
	if (kernel == 1) {  // :150-187
		const int cell = containing_cell(*f, pos);
		if (cell < 0) { return 0; }
		const int corners = 1 << D;
		for (int d = 0; d < D; ++d) {
			const int   r    = f->eq.next_row();
			const float term = cw * 2.0f / static_cast<float>(corners);
			for (int c = 0; c < corners; ++c) {
				int node = cell;
				for (int a = 0; a < D; ++a) { node += f->stride[a] * ((c >> a) & 1); }
				const float sign = ((c >> d) & 1) ? +1.0f : -1.0f;
				f->eq.put(r, node, sign * term);
			}
			f->eq.rhs.push_back(cw * grad[d]);
		}
		return 1;
	}
	if (kernel == 2) {  // :188-236; duplicate columns inside a row stay separate triplets
		float shifted[kMaxDim];
		for (int d = 0; d < D; ++d) { shifted[d] = pos[d] - 0.5f; }
		int   idx[8];
		float k[8];
		const int n = corner_weights(f->size, f->stride, D, shifted, 1, idx, k);
		if (n == 0) { return 0; }
		for (int d = 0; d < D; ++d) {
			const int r   = f->eq.next_row();
			float     sum = 0;
			for (int i = 0; i < n; ++i) {
				const float c = k[i] * cw;
				f->eq.put(r, idx[i], -c);
				f->eq.put(r, idx[i] + f->stride[d], +c);
				sum += c;
			}
			f->eq.rhs.push_back(sum * grad[d]);
		}
		return 1;
	}
	return -1;
}

// field_interpolation.cpp:243-316 — smoothness rows anchored at one lattice node along one axis.
void model_rows_at(Lattice* f, const OraWeights& w, const int* coord, int index, int d)
{
	static const float binom[5][5] = {
		{1, 0, 0, 0, 0}, {-1, 1, 0, 0, 0}, {1, -2, 1, 0, 0}, {1, -3, 3, -1, 0}, {1, -4, 6, -4, 1}};
	const float order_w[5] = {w.model_0, w.model_1, w.model_2, w.model_3, w.model_4};
	const int   n = f->size[d], s = f->stride[d], c = coord[d];
	for (int k = 0; k <= 4; ++k) {
		if (!(order_w[k] > 0 && 0 <= c && c + k < n)) { continue; }
		int cols[5];
		for (int m = 0; m <= k; ++m) { cols[m] = index + m * s; }
		append_row(&f->eq, order_w[k], 0.0f, k + 1, cols, binom[k]);
	}
	if (w.gradient_smoothness > 0 && 0 <= c && c + 1 < n) {  // :303-315
		for (int o = 0; o < f->ndim; ++o) {
			if (o == d || coord[o] + 1 >= f->size[o]) { continue; }
			const int   so      = f->stride[o];
			const int   cols[4] = {index, index + s, index + so, index + so + s};
			const float coef[4] = {-1.0f, +1.0f, +1.0f, -1.0f};
			append_row(&f->eq, w.gradient_smoothness, 0.0f, 4, cols, coef);
		}
	}
}

// field_interpolation.cpp:318-341.
void model_rows(Lattice* f, const OraWeights& w)
{
	const int64_t N = f->unknowns();
	for (int64_t index = 0; index < N; ++index) {
		int     coord[kMaxDim];
		int64_t rest = index;
		for (int d = 0; d < f->ndim; ++d) { coord[d] = static_cast<int>(rest % f->size[d]); rest /= f->size[d]; }
		for (int d = 0; d < f->ndim; ++d) { model_rows_at(f, w, coord, static_cast<int>(index), d); }
	}
}

// field_interpolation.cpp:343-371.
void point_rows(Lattice* f, float vw, int vkernel, float gw, int gkernel, int64_t npts, const float* pos,
                const float* normals, const float* pw)
{
	const int D = f->ndim;
	for (int64_t i = 0; i < npts; ++i) {
		const float  w = pw ? pw[i] : 1.0f;
		const float* p = pos + i * D;
		const float* g = normals ? normals + i * D : nullptr;
		if (vkernel == 0) {
			if (!g) { return; }  // the reference CHECK-aborts on null normals (:361)
			value_row_nearest(f, p, g, 0.0f, w * vw);
		} else {
			value_row(f, p, 0.0f, w * vw);
		}
		if (g) { gradient_rows(f, p, g, w * gw, gkernel); }
	}
}

// ------------------------------------------------------------------------------------------------
// Solve half: sparse_linear.cpp:59-113 (+ Eigen).  Column-compressed matrix with sorted inner indices.

template <typename T>
struct Csc
{
	int64_t              nrows = 0, ncols = 0;
	std::vector<int64_t> ptr;  // ncols + 1
	std::vector<int>     idx;
	std::vector<T>       val;
};

// Eigen::SparseMatrix::setFromTriplets: duplicates summed, inner indices sorted.  The float path keeps
// explicit zeros (sparse_linear.cpp:59-70); the double path drops them first (:84-86).
template <typename T>
Csc<T> csc_from_triplets(int64_t nt, const int* r, const int* c, const float* v, int64_t nrows, int64_t ncols, bool drop_zeros)
{
	// Pass 1: row-major bucket (this is what Eigen does: build the transposed matrix first, then
	// transpose-assign, which sorts inner indices), summing duplicates in input order.
	std::vector<int64_t> rptr(nrows + 1, 0);
	for (int64_t i = 0; i < nt; ++i) {
		if (drop_zeros && v[i] == 0.0f) { continue; }
		++rptr[r[i] + 1];
	}
	for (int64_t i = 0; i < nrows; ++i) { rptr[i + 1] += rptr[i]; }
	std::vector<int>     rcol(rptr[nrows]);
	std::vector<T>       rval(rptr[nrows]);
	std::vector<int64_t> fill(rptr.begin(), rptr.end() - 1);
	for (int64_t i = 0; i < nt; ++i) {
		if (drop_zeros && v[i] == 0.0f) { continue; }
		const int64_t at = fill[r[i]]++;
		rcol[at] = c[i];
		rval[at] = static_cast<T>(v[i]);
	}
	// collapse duplicates inside each row (first occurrence keeps the slot, later ones add in order)
	std::vector<int64_t> seen(ncols, -1);
	std::vector<int64_t> rlen(nrows, 0);
	for (int64_t row = 0; row < nrows; ++row) {
		const int64_t b = rptr[row], e = rptr[row + 1];
		int64_t       w = b;
		for (int64_t k = b; k < e; ++k) {
			const int col = rcol[k];
			if (seen[col] >= b) {
				rval[seen[col]] += rval[k];
			} else {
				seen[col] = w;
				rcol[w]   = col;
				rval[w]   = rval[k];
				++w;
			}
		}
		rlen[row] = w - b;
	}
	// Pass 2: transpose into column-compressed form; rows are visited in ascending order so inner
	// indices come out sorted.
	Csc<T> A;
	A.nrows = nrows;
	A.ncols = ncols;
	A.ptr.assign(ncols + 1, 0);
	for (int64_t row = 0; row < nrows; ++row) {
		for (int64_t k = rptr[row]; k < rptr[row] + rlen[row]; ++k) { ++A.ptr[rcol[k] + 1]; }
	}
	for (int64_t j = 0; j < ncols; ++j) { A.ptr[j + 1] += A.ptr[j]; }
	A.idx.resize(A.ptr[ncols]);
	A.val.resize(A.ptr[ncols]);
	std::vector<int64_t> cfill(A.ptr.begin(), A.ptr.end() - 1);
	for (int64_t row = 0; row < nrows; ++row) {
		for (int64_t k = rptr[row]; k < rptr[row] + rlen[row]; ++k) {
			const int64_t at = cfill[rcol[k]]++;
			A.idx[at] = static_cast<int>(row);
			A.val[at] = rval[k];
		}
	}
	return A;
}

template <typename T>
Csc<T> transpose(const Csc<T>& A)
{
	Csc<T> B;
	B.nrows = A.ncols;
	B.ncols = A.nrows;
	B.ptr.assign(B.ncols + 1, 0);
	for (int i : A.idx) { ++B.ptr[i + 1]; }
	for (int64_t j = 0; j < B.ncols; ++j) { B.ptr[j + 1] += B.ptr[j]; }
	B.idx.resize(A.idx.size());
	B.val.resize(A.val.size());
	std::vector<int64_t> fill(B.ptr.begin(), B.ptr.end() - 1);
	for (int64_t j = 0; j < A.ncols; ++j) {
		for (int64_t k = A.ptr[j]; k < A.ptr[j + 1]; ++k) {
			const int64_t at = fill[A.idx[k]]++;
			B.idx[at] = static_cast<int>(j);
			B.val[at] = A.val[k];
		}
	}
	return B;
}

// make_square, sparse_linear.cpp:105-113: AtA = Aᵀ·A with Eigen's conservative column-by-column
// product (structural: numerically-zero results are kept; columns sorted afterwards).
template <typename T>
Csc<T> normal_matrix(const Csc<T>& A)
{
	const Csc<T> At = transpose(A);  // column r of At = row r of A
	Csc<T> M;
	M.nrows = M.ncols = A.ncols;
	M.ptr.assign(A.ncols + 1, 0);
	std::vector<T>       acc(A.ncols, T(0));
	std::vector<char>    mark(A.ncols, 0);
	std::vector<int>     touched;
	for (int64_t j = 0; j < A.ncols; ++j) {
		touched.clear();
		for (int64_t k = A.ptr[j]; k < A.ptr[j + 1]; ++k) {
			const int r = A.idx[k];
			const T   y = A.val[k];
			for (int64_t m = At.ptr[r]; m < At.ptr[r + 1]; ++m) {
				const int i = At.idx[m];
				if (!mark[i]) { mark[i] = 1; acc[i] = At.val[m] * y; touched.push_back(i); }
				else          { acc[i] += At.val[m] * y; }
			}
		}
		std::sort(touched.begin(), touched.end());
		for (int i : touched) {
			M.idx.push_back(i);
			M.val.push_back(acc[i]);
			mark[i] = 0;
		}
		M.ptr[j + 1] = static_cast<int64_t>(M.idx.size());
	}
	return M;
}

template <typename T>
struct Normal  // AtA (symmetric, stored column-compressed = row-compressed) and Atb
{
	Csc<T>         M;
	std::vector<T> atb;
	std::vector<T> inv_diag;  // Eigen DiagonalPreconditioner: 1/d, or 1 where d == 0
};

template <typename T>
void spmv(const Csc<T>& M, const T* x, T* y)  // symmetric ⇒ treat columns as rows (gather form)
{
	for (int64_t j = 0; j < M.ncols; ++j) {
		T s = 0;
		for (int64_t k = M.ptr[j]; k < M.ptr[j + 1]; ++k) { s += M.val[k] * x[M.idx[k]]; }
		y[j] = s;
	}
}

template <typename T>
T dot(const std::vector<T>& a, const std::vector<T>& b)
{
	T s = 0;
	for (size_t i = 0; i < a.size(); ++i) { s += a[i] * b[i]; }
	return s;
}

template <typename T>
Normal<T>* build_normal(int64_t nt, const int* r, const int* c, const float* v, int64_t nrows, const float* rhs, int64_t ncols,
                        bool drop_zeros)
{
	auto*        N = new Normal<T>();
	const Csc<T> A = csc_from_triplets<T>(nt, r, c, v, nrows, ncols, drop_zeros);
	N->M           = normal_matrix(A);
	N->atb.assign(ncols, T(0));
	for (int64_t j = 0; j < ncols; ++j) {  // Atb = Aᵀ·rhs (:120,159,196)
		T s = 0;
		for (int64_t k = A.ptr[j]; k < A.ptr[j + 1]; ++k) { s += A.val[k] * static_cast<T>(rhs[A.idx[k]]); }
		N->atb[j] = s;
	}
	N->inv_diag.assign(ncols, T(1));
	for (int64_t j = 0; j < ncols; ++j) {
		for (int64_t k = N->M.ptr[j]; k < N->M.ptr[j + 1]; ++k) {
			if (N->M.idx[k] == j && N->M.val[k] != T(0)) { N->inv_diag[j] = T(1) / N->M.val[k]; }
		}
	}
	return N;
}

// Eigen 3.3 BiCGSTAB body (Eigen/src/IterativeLinearSolvers/BiCGSTAB.h, restated from the published
// algorithm) with the diagonal preconditioner; wiring and defaults follow
// sparse_linear.cpp:186-212 and :427-440 (max_iterations<=0 ⇒ 2·n, tolerance<=0 ⇒ epsilon).
template <typename T>
void bicgstab(const Normal<T>& N, T* x_io, int64_t max_iter, T tol, int64_t* iters_out, T* err_out)
{
	const int64_t  n = N.M.ncols;
	if (max_iter <= 0) { max_iter = 2 * n; }
	if (!(tol > 0)) { tol = std::numeric_limits<T>::epsilon(); }
	std::vector<T> x(x_io, x_io + n), r(n), r0(n), v(n, 0), p(n, 0), y(n), z(n), s(n), t(n), tmp(n);
	auto residual = [&]() {
		spmv(N.M, x.data(), tmp.data());
		for (int64_t i = 0; i < n; ++i) { r[i] = N.atb[i] - tmp[i]; }
	};
	residual();
	r0 = r;
	T       r0_sq  = dot(r0, r0);
	const T rhs_sq = dot(N.atb, N.atb);
	if (rhs_sq == 0) {
		std::fill(x_io, x_io + n, T(0));
		*iters_out = 0;
		*err_out   = 0;
		return;
	}
	T       rho = 1, alpha = 1, w = 1;
	const T tol2 = tol * tol * rhs_sq;
	const T eps2 = std::numeric_limits<T>::epsilon() * std::numeric_limits<T>::epsilon();
	int64_t i = 0, restarts = 0;
	while (dot(r, r) > tol2 && i < max_iter) {
		const T rho_old = rho;
		rho             = dot(r0, r);
		if (std::abs(rho) < eps2 * r0_sq) {
			residual();
			r0  = r;
			rho = r0_sq = dot(r, r);
			if (restarts++ == 0) { i = 0; }
		}
		const T beta = (rho / rho_old) * (alpha / w);
		for (int64_t k = 0; k < n; ++k) { p[k] = r[k] + beta * (p[k] - w * v[k]); }
		for (int64_t k = 0; k < n; ++k) { y[k] = N.inv_diag[k] * p[k]; }
		spmv(N.M, y.data(), v.data());
		alpha = rho / dot(r0, v);
		for (int64_t k = 0; k < n; ++k) { s[k] = r[k] - alpha * v[k]; }
		for (int64_t k = 0; k < n; ++k) { z[k] = N.inv_diag[k] * s[k]; }
		spmv(N.M, z.data(), t.data());
		const T tt = dot(t, t);
		w          = tt > T(0) ? dot(t, s) / tt : T(0);
		for (int64_t k = 0; k < n; ++k) { x[k] += alpha * y[k] + w * z[k]; }
		for (int64_t k = 0; k < n; ++k) { r[k] = s[k] - w * t[k]; }
		++i;
	}
	*err_out   = std::sqrt(dot(r, r) / rhs_sq);
	*iters_out = i;
	std::copy(x.begin(), x.end(), x_io);
}

// Jacobi-preconditioned conjugate gradients on AtA with Eigen's stopping rule ‖r‖ ≤ tol·‖Atb‖: the CPU
// statement of the algorithm the CUDA path runs (north-star item (c)), used for like-for-like parity.
template <typename T>
void pcg(const Normal<T>& N, T* x_io, int64_t max_iter, T tol, int64_t* iters_out, T* err_out)
{
	const int64_t  n = N.M.ncols;
	if (max_iter <= 0) { max_iter = 2 * n; }
	std::vector<T> x(x_io, x_io + n), r(n), z(n), p(n), q(n);
	spmv(N.M, x.data(), q.data());
	for (int64_t i = 0; i < n; ++i) { r[i] = N.atb[i] - q[i]; }
	const T bb = dot(N.atb, N.atb);
	if (bb == 0) {
		std::fill(x_io, x_io + n, T(0));
		*iters_out = 0;
		*err_out   = 0;
		return;
	}
	for (int64_t i = 0; i < n; ++i) { z[i] = N.inv_diag[i] * r[i]; }
	p       = z;
	T rho   = dot(r, z);
	T rr    = dot(r, r);
	int64_t it = 0;
	while (rr > tol * tol * bb && it < max_iter) {
		spmv(N.M, p.data(), q.data());
		const T pq = dot(p, q);
		if (!(pq > 0)) { break; }
		const T alpha = rho / pq;
		for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
		for (int64_t i = 0; i < n; ++i) { z[i] = N.inv_diag[i] * r[i]; }
		const T rho_new = dot(r, z);
		rr              = dot(r, r);
		const T beta    = rho_new / rho;
		rho             = rho_new;
		for (int64_t i = 0; i < n; ++i) { p[i] = z[i] + beta * p[i]; }
		++it;
	}
	*iters_out = it;
	*err_out   = std::sqrt(rr / bb);
	std::copy(x.begin(), x.end(), x_io);
}

// jacobi_iterations, sparse_linear.cpp:214-241: x ← w·(Atb − R x)/D + (1 − w)·x with R = AtA − diag.
template <typename T>
void jacobi(const Normal<T>& N, T* x_io, int iterations, T weight)
{
	const int64_t  n = N.M.ncols;
	std::vector<T> x(x_io, x_io + n), tmp(n), diag(n, T(0));
	for (int64_t j = 0; j < n; ++j) {
		for (int64_t k = N.M.ptr[j]; k < N.M.ptr[j + 1]; ++k) {
			if (N.M.idx[k] == j) { diag[j] = N.M.val[k]; }
		}
	}
	for (int it = 0; it < iterations; ++it) {
		for (int64_t j = 0; j < n; ++j) {
			T s = 0;
			for (int64_t k = N.M.ptr[j]; k < N.M.ptr[j + 1]; ++k) {
				if (N.M.idx[k] != j) { s += N.M.val[k] * x[N.M.idx[k]]; }
			}
			tmp[j] = N.atb[j] - s;
		}
		for (int64_t j = 0; j < n; ++j) { x[j] = weight * tmp[j] / diag[j] + (T(1) - weight) * x[j]; }
	}
	std::copy(x.begin(), x.end(), x_io);
}

// tile_solver_square, sparse_linear.cpp:246-390.  Each tile's matrix is dense here (tile_size^D unknowns) and
// factorised by a plain Cholesky in T — Eigen's SimplicialLLT computes the same factor up to ordering and
// rounding.  Returns the number of failed tiles (:338,:364-365: a failed tile keeps the guess).
template <typename T>
int tile_solve(const Normal<T>& N, const T* guess, int ndim, const int* sizes, int tile_size, T* out)
{
	const int64_t        n = N.M.ncols;
	std::vector<int>     num_tiles(ndim);
	int64_t              tiles_total = 1, per_tile = 1;
	for (int d = 0; d < ndim; ++d) {  // :258-266
		num_tiles[d] = (sizes[d] + tile_size - 1) / tile_size;
		tiles_total *= num_tiles[d];
		per_tile *= tile_size;
	}
	auto locate = [&](int64_t full, int64_t* tile, int64_t* in_tile) {  // calc_tile_and_index, :270-293
		int64_t t = 0, i = 0, ts = 1, is = 1;
		for (int d = 0; d < ndim; ++d) {
			const int64_t x = full % sizes[d];
			t += (x / tile_size) * ts;
			i += (x % tile_size) * is;
			full /= sizes[d];
			ts *= num_tiles[d];
			is *= tile_size;
		}
		*tile    = t;
		*in_tile = i;
	};
	struct Entry { int64_t r, c; T v; };
	std::vector<std::vector<Entry>> trip(tiles_total);
	std::vector<std::vector<T>>     rhs(tiles_total, std::vector<T>(per_tile, T(0)));
	for (int64_t full = 0; full < n; ++full) {  // :313-318
		int64_t t, i;
		locate(full, &t, &i);
		rhs[t][i] = N.atb[full];
	}
	for (int64_t k = 0; k < n; ++k) {  // :320-336, every stored entry (both triangles)
		for (int64_t e = N.M.ptr[k]; e < N.M.ptr[k + 1]; ++e) {
			const int64_t row = N.M.idx[e], col = k;
			const T       v   = N.M.val[e];
			int64_t       rt, ri, ct, ci;
			locate(row, &rt, &ri);
			locate(col, &ct, &ci);
			if (rt == ct) {
				trip[rt].push_back({ri, ci, v});
			} else {
				rhs[rt][ri] -= v * guess[col];
				rhs[ct][ci] -= v * guess[row];
			}
		}
	}
	std::copy(guess, guess + n, out);
	int              failures = 0;
	std::vector<T>   A(static_cast<size_t>(per_tile * per_tile)), y(per_tile);
	for (int64_t t = 0; t < tiles_total; ++t) {
		if (trip[t].empty()) { continue; }  // only the regularisation: skipped (:345-348)
		std::fill(A.begin(), A.end(), T(0));
		for (int64_t i = 0; i < per_tile; ++i) { A[i * per_tile + i] = T(1e-6f); }  // :306-309
		for (const Entry& e : trip[t]) { A[e.r * per_tile + e.c] += e.v; }
		bool ok = true;  // Cholesky A = L L^T in place (lower triangle)
		for (int64_t j = 0; j < per_tile && ok; ++j) {
			T d = A[j * per_tile + j];
			for (int64_t k = 0; k < j; ++k) { d -= A[j * per_tile + k] * A[j * per_tile + k]; }
			if (!(d > T(0))) { ok = false; break; }
			d = std::sqrt(d);
			A[j * per_tile + j] = d;
			for (int64_t i = j + 1; i < per_tile; ++i) {
				T s = A[i * per_tile + j];
				for (int64_t k = 0; k < j; ++k) { s -= A[i * per_tile + k] * A[j * per_tile + k]; }
				A[i * per_tile + j] = s / d;
			}
		}
		if (!ok) { ++failures; continue; }
		for (int64_t i = 0; i < per_tile; ++i) {
			T s = rhs[t][i];
			for (int64_t k = 0; k < i; ++k) { s -= A[i * per_tile + k] * y[k]; }
			y[i] = s / A[i * per_tile + i];
		}
		for (int64_t i = per_tile - 1; i >= 0; --i) {
			T s = y[i];
			for (int64_t k = i + 1; k < per_tile; ++k) { s -= A[k * per_tile + i] * y[k]; }
			y[i] = s / A[i * per_tile + i];
		}
		for (int64_t i = 0; i < per_tile; ++i) {  // :369-386: scatter back, unknowns outside the lattice dropped
			int64_t tc = t, ic = i, stride = 1, full = 0;
			bool    inside = true;
			for (int d = 0; d < ndim; ++d) {
				const int64_t x = (tc % num_tiles[d]) * tile_size + ic % tile_size;
				inside          = inside && x < sizes[d];
				full += x * stride;
				tc /= num_tiles[d];
				ic /= tile_size;
				stride *= sizes[d];
			}
			if (inside) { out[full] = y[i]; }
		}
	}
	return failures;
}


} // namespace

extern "C" {

void* ora_field_create(int ndim, const int* sizes) { return new Lattice(ndim, sizes); }
void  ora_field_destroy(void* f) { delete static_cast<Lattice*>(f); }

void ora_add_field_constraints(void* f, const OraWeights* w) { model_rows(static_cast<Lattice*>(f), *w); }

int ora_add_value_constraint(void* f, const float* pos, float value, float weight)
{
	return value_row(static_cast<Lattice*>(f), pos, value, weight) ? 1 : 0;
}

int ora_add_value_constraint_nearest_neighbor(void* f, const float* pos, const float* gradient, float value, float weight)
{
	return value_row_nearest(static_cast<Lattice*>(f), pos, gradient, value, weight) ? 1 : 0;
}

int ora_add_gradient_constraint(void* f, const float* pos, const float* gradient, float weight, int kernel)
{
	return gradient_rows(static_cast<Lattice*>(f), pos, gradient, weight, kernel);
}

void ora_add_points(void* f, float value_weight, int value_kernel, float gradient_weight, int gradient_kernel,
                    int num_points, const float* positions, const float* normals, const float* point_weights)
{
	point_rows(static_cast<Lattice*>(f), value_weight, value_kernel, gradient_weight, gradient_kernel, num_points,
	           positions, normals, point_weights);
}

void ora_add_equation(void* f, float weight, float rhs, int num_pairs, const int* columns, const float* values)
{
	append_row(&static_cast<Lattice*>(f)->eq, weight, rhs, num_pairs, columns, values);
}

// field_interpolation.cpp:373-400: model rows first, then the point rows.
void* ora_sdf_from_points(int ndim, const int* sizes, const OraWeights* w, int num_points, const float* positions,
                          const float* normals, const float* point_weights)
{
	auto* f = new Lattice(ndim, sizes);
	model_rows(f, *w);
	point_rows(f, w->data_pos, w->value_kernel, w->data_gradient, w->gradient_kernel, num_points, positions, normals,
	           point_weights);
	return f;
}

int64_t ora_num_rows(void* f) { return static_cast<int64_t>(static_cast<Lattice*>(f)->eq.rhs.size()); }
int64_t ora_num_triplets(void* f) { return static_cast<int64_t>(static_cast<Lattice*>(f)->eq.val.size()); }

void ora_copy_system(void* f, int* rows, int* cols, float* values, float* rhs)
{
	const System& eq = static_cast<Lattice*>(f)->eq;
	const size_t  nt = eq.val.size();
	if (nt) {
		std::memcpy(rows, eq.row.data(), nt * sizeof(int));
		std::memcpy(cols, eq.col.data(), nt * sizeof(int));
		std::memcpy(values, eq.val.data(), nt * sizeof(float));
	}
	if (!eq.rhs.empty()) { std::memcpy(rhs, eq.rhs.data(), eq.rhs.size() * sizeof(float)); }
}

// upscale_field, field_interpolation.cpp:431-485 (multilinear, align-corners; mul then div in fp32 at :462).
void ora_upscale_field(const float* small_field, int ndim, const int* small_sizes, const int* large_sizes, float* out)
{
	int     sstride[kMaxDim] = {1, 1, 1};
	int64_t big              = 1;
	{
		int s = 1;
		for (int d = 0; d < ndim; ++d) { sstride[d] = s; s *= small_sizes[d]; big *= large_sizes[d]; }
	}
	for (int64_t li = 0; li < big; ++li) {
		float   pos[kMaxDim];
		int64_t rest = li;
		for (int d = 0; d < ndim; ++d) {
			const int c = static_cast<int>(rest % large_sizes[d]);
			rest /= large_sizes[d];
			pos[d] = static_cast<float>(c) * (static_cast<float>(small_sizes[d]) - 1.0f) / (static_cast<float>(large_sizes[d]) - 1.0f);
		}
		int       idx[8];
		float     k[8];
		const int n = corner_weights(small_sizes, sstride, ndim, pos, 0, idx, k);
		float wsum = 0, fsum = 0;
		for (int i = 0; i < n; ++i) { wsum += k[i]; fsum += k[i] * small_field[idx[i]]; }
		out[li] = (wsum == 0) ? 0.0f : fsum / wsum;
	}
}

// generate_error_map, field_interpolation.cpp:402-429.
void ora_generate_error_map(int64_t num_triplets, const int* rows, const int* cols, const float* values,
                            int64_t num_unknowns, const float* solution, int64_t num_rows, const float* rhs, float* out)
{
	std::vector<float> err(rhs, rhs + num_rows), sq(num_rows, 0.0f);
	for (int64_t i = 0; i < num_triplets; ++i) {
		err[rows[i]] -= solution[cols[i]] * values[i];
		sq[rows[i]] += values[i] * values[i];
	}
	for (float& e : err) { e *= e; }
	std::fill(out, out + num_unknowns, 0.0f);
	for (int64_t i = 0; i < num_triplets; ++i) {
		if (sq[rows[i]] != 0) { out[cols[i]] += (values[i] * values[i]) / sq[rows[i]] * err[rows[i]]; }
	}
}

// ---- iso-surface helpers the demo runs right after every solve ------------------
// emilib::marching_squares, third_party/emilib/emilib/marching_squares.cpp:11-134.  `iso` is row-major
// width x height; a corner is "outside" when its value is >= 0 (:21-24); cells are visited y-major (:15-16) and
// emit 0, 1 or 2 directed segments x0 y0 x1 y1.  The reference is one switch over 14 cases (:36-129); here the
// same cases are a table of directed edge pairs over the four crossing points
//   L = (x, y + tl/(tl-bl))  T = (x + tl/(tl-tr), y)  R = (x+1, y + tr/(tr-br))  B = (x + bl/(bl-br), y+1)   (:31-34).
// Returns the number of floats; writes them when out != nullptr.
int64_t ora_marching_squares(int64_t width, int64_t height, const float* iso, float* out)
{
	enum { L, T, R, B };
	// config = br<<3 | bl<<2 | tr<<1 | tl (:26); up to two (from, to) pairs, -1 terminated
	static const int kSeg[16][4] = {
		{-1, -1, -1, -1}, {L, T, -1, -1}, {T, R, -1, -1}, {L, R, -1, -1}, {B, L, -1, -1}, {B, T, -1, -1}, {T, L, B, R}, {B, R, -1, -1},
		{R, B, -1, -1},   {L, T, R, B},   {T, B, -1, -1}, {L, B, -1, -1}, {R, L, -1, -1}, {R, T, -1, -1}, {T, L, -1, -1}, {-1, -1, -1, -1}};
	int64_t n = 0;
	for (int64_t y = 0; y + 1 < height; ++y) {
		for (int64_t x = 0; x + 1 < width; ++x) {
			const float tl = iso[x + width * y], tr = iso[x + 1 + width * y];
			const float bl = iso[x + width * (y + 1)], br = iso[x + 1 + width * (y + 1)];
			const int config = (br >= 0.0f ? 8 : 0) | (bl >= 0.0f ? 4 : 0) | (tr >= 0.0f ? 2 : 0) | (tl >= 0.0f ? 1 : 0);
			const int* seg = kSeg[config];
			if (seg[0] < 0) { continue; }
			const float fx = static_cast<float>(x), fy = static_cast<float>(y);
			const float px[4] = {fx + 0.0f, fx + tl / (tl - tr), fx + 1.0f, fx + bl / (bl - br)};
			const float py[4] = {fy + tl / (tl - bl), fy + 0.0f, fy + tr / (tr - br), fy + 1.0f};
			for (int k = 0; k < 4 && seg[k] >= 0; k += 2) {
				if (out) {
					out[n + 0] = px[seg[k]];
					out[n + 1] = py[seg[k]];
					out[n + 2] = px[seg[k + 1]];
					out[n + 3] = py[seg[k + 1]];
				}
				n += 4;
			}
		}
	}
	return n;
}

// emilib::calc_area, marching_squares.cpp:136-150: shoelace sum in double over the segments in order, halved, as float.
float ora_calc_area(int64_t num_segments, const float* xy)
{
	double twice = 0;
	for (int64_t i = 0; i < num_segments; ++i) {
		const double ax = xy[4 * i], ay = xy[4 * i + 1], bx = xy[4 * i + 2], by = xy[4 * i + 3];
		twice += ax * by - bx * ay;
	}
	return static_cast<float>(twice / 2);
}

// emath::catmull_rom, third_party/emath/emath/math.hpp:308-316, in the reference's fp32 evaluation order.
static float catmull_rom_f32(float t, float p0, float p1, float p2, float p3)
{
	const float a = p0 * t * ((2.0f - t) * t - 1.0f);
	const float b = p1 * (t * t * (3.0f * t - 5.0f) + 2.0f);
	const float c = p2 * t * ((4.0f - 3.0f * t) * t + 1.0f);
	const float d = p3 * (t - 1.0f) * t * t;
	return 0.5f * (a + b + c + d);
}

// bicubic_upsample, src/sdf_field.cpp:555-603: large = upsample * small - upsample + 1 per axis; sample (lx, ly) sits at
// cell (lx / upsample, ly / upsample) with t = (l % upsample) / upsample; 4 x 4 neighbourhood with clamped reads (:565-570),
// Catmull-Rom along x for each of the four rows, then along y (:587-594).  out: large_width * large_height floats.
void ora_bicubic_upsample(int width, int height, const float* values, int upsample, float* out)
{
	const int64_t lw = static_cast<int64_t>(upsample) * width - upsample + 1, lh = static_cast<int64_t>(upsample) * height - upsample + 1;
	auto at = [&](int x, int y) {
		x = std::min(std::max(x, 0), width - 1);
		y = std::min(std::max(y, 0), height - 1);
		return values[static_cast<int64_t>(y) * width + x];
	};
	for (int64_t ly = 0; ly < lh; ++ly) {
		for (int64_t lx = 0; lx < lw; ++lx) {
			const float tx = static_cast<float>(lx % upsample) / static_cast<float>(upsample);
			const float ty = static_cast<float>(ly % upsample) / static_cast<float>(upsample);
			const int   sx = static_cast<int>(lx / upsample), sy = static_cast<int>(ly / upsample);
			float row[4];
			for (int j = 0; j < 4; ++j) { row[j] = catmull_rom_f32(tx, at(sx - 1, sy - 1 + j), at(sx, sy - 1 + j), at(sx + 1, sy - 1 + j), at(sx + 2, sy - 1 + j)); }
			out[ly * lw + lx] = catmull_rom_f32(ty, row[0], row[1], row[2], row[3]);
		}
	}
}

// iso_surface, src/sdf_field.cpp:605-614: marching squares of (values - iso).
int64_t ora_iso_surface(int width, int height, const float* values, float iso, float* out)
{
	std::vector<float> shifted(static_cast<size_t>(width) * height);
	for (size_t i = 0; i < shifted.size(); ++i) { shifted[i] = values[i] - iso; }
	return ora_marching_squares(width, height, shifted.data(), out);
}

// ---- solve half -------------------------------------------------------------------------------
// precision: 0 = float (as_sparse_matrix_float, keeps explicit zeros), 1 = double (drops them).
void* ora_normal_create(int64_t nt, const int* rows, const int* cols, const float* values, int64_t num_rows,
                        const float* rhs, int64_t num_columns, int precision)
{
	if (precision == 0) { return build_normal<float>(nt, rows, cols, values, num_rows, rhs, num_columns, false); }
	return build_normal<double>(nt, rows, cols, values, num_rows, rhs, num_columns, true);
}

void ora_normal_destroy(void* n, int precision)
{
	if (precision == 0) { delete static_cast<Normal<float>*>(n); } else { delete static_cast<Normal<double>*>(n); }
}

int64_t ora_normal_nnz(void* n, int precision)
{
	return precision == 0 ? static_cast<int64_t>(static_cast<Normal<float>*>(n)->M.val.size())
	                      : static_cast<int64_t>(static_cast<Normal<double>*>(n)->M.val.size());
}

void ora_normal_copy(void* n, int precision, int64_t* indptr, int* indices, double* data, double* atb)
{
	auto dump = [&](const auto& N) {
		std::copy(N.M.ptr.begin(), N.M.ptr.end(), indptr);
		std::copy(N.M.idx.begin(), N.M.idx.end(), indices);
		for (size_t i = 0; i < N.M.val.size(); ++i) { data[i] = static_cast<double>(N.M.val[i]); }
		for (size_t i = 0; i < N.atb.size(); ++i) { atb[i] = static_cast<double>(N.atb[i]); }
	};
	if (precision == 0) { dump(*static_cast<Normal<float>*>(n)); } else { dump(*static_cast<Normal<double>*>(n)); }
}

void ora_bicgstab_f32(void* n, float* x, int64_t max_iter, float tol, int64_t* iters, float* err)
{
	bicgstab(*static_cast<Normal<float>*>(n), x, max_iter, tol, iters, err);
}
void ora_bicgstab_f64(void* n, double* x, int64_t max_iter, double tol, int64_t* iters, double* err)
{
	bicgstab(*static_cast<Normal<double>*>(n), x, max_iter, tol, iters, err);
}
void ora_pcg_f32(void* n, float* x, int64_t max_iter, float tol, int64_t* iters, float* err)
{
	pcg(*static_cast<Normal<float>*>(n), x, max_iter, tol, iters, err);
}
void ora_pcg_f64(void* n, double* x, int64_t max_iter, double tol, int64_t* iters, double* err)
{
	pcg(*static_cast<Normal<double>*>(n), x, max_iter, tol, iters, err);
}
void ora_jacobi_f32(void* n, float* x, int iterations, float weight)
{
	jacobi(*static_cast<Normal<float>*>(n), x, iterations, weight);
}
int ora_tile_solve_f32(void* n, const float* guess, int ndim, const int* sizes, int tile_size, float* out)
{
	return tile_solve(*static_cast<Normal<float>*>(n), guess, ndim, sizes, tile_size, out);
}
int ora_tile_solve_f64(void* n, const double* guess, int ndim, const int* sizes, int tile_size, double* out)
{
	return tile_solve(*static_cast<Normal<double>*>(n), guess, ndim, sizes, tile_size, out);
}
void ora_apply_f64(void* n, const double* x, double* y)
{
	const auto& N = *static_cast<Normal<double>*>(n);
	spmv(N.M, x, y);
}
void ora_apply_f32(void* n, const float* x, float* y)
{
	const auto& N = *static_cast<Normal<float>*>(n);
	spmv(N.M, x, y);
}

} // extern "C"
