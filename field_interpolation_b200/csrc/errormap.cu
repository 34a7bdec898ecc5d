// generate_error_map (reference field_interpolation/field_interpolation.cpp:402-429) on the GPU: the squared
// residual of every row of A x = b is split over the row's columns in proportion to the squared coefficients —
// the heat map the reference demo draws after every solve (src/sdf_field.cpp:350).
//
// Works on any triplet list (rows need not be grouped).  Three passes: per triplet accumulate A x and the sum of
// squared coefficients per row; per row square the error; per triplet scatter the blame.  Accumulation is by
// fp32 atomics, so sums agree with the reference's sequential fp32 sums to rounding, not bit for bit.
#include "internal.hpp"

namespace fi {

namespace {

constexpr int kThreads = 256;

__global__ void rows_accumulate_kernel(int64_t nt, const fi_triplet* __restrict__ t, const float* __restrict__ x, int64_t n, int64_t nrows,
                                       float* __restrict__ err, float* __restrict__ sumsq, int* bad)
{
	const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (k >= nt) { return; }
	const fi_triplet e = t[k];
	if (e.row < 0 || e.row >= nrows || e.col < 0 || e.col >= n) {
		*bad = 1;
		return;
	}
	atomicAdd(&err[e.row], -(x[e.col] * e.value));
	atomicAdd(&sumsq[e.row], e.value * e.value);
}

__global__ void square_kernel(int64_t nrows, float* __restrict__ err)
{
	const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (r < nrows) { err[r] = err[r] * err[r]; }
}

__global__ void blame_kernel(int64_t nt, const fi_triplet* __restrict__ t, int64_t n, int64_t nrows, const float* __restrict__ err,
                             const float* __restrict__ sumsq, float* __restrict__ heat)
{
	const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
	if (k >= nt) { return; }
	const fi_triplet e = t[k];
	if (e.row < 0 || e.row >= nrows || e.col < 0 || e.col >= n) { return; }
	const float s = sumsq[e.row];
	if (s != 0.0f) { atomicAdd(&heat[e.col], (e.value * e.value) / s * err[e.row]); }
}

}  // namespace

void error_map(int64_t nt, const fi_triplet* h_trips, int64_t n, const float* h_x, int64_t nrows, const float* h_rhs, float* h_out)
{
	cudaStream_t       s = nullptr;
	DevBuf<fi_triplet> t(std::max<int64_t>(nt, 1));
	DevBuf<float>      x(std::max<int64_t>(n, 1)), err(std::max<int64_t>(nrows, 1)), sumsq(std::max<int64_t>(nrows, 1)), heat(std::max<int64_t>(n, 1));
	DevBuf<int>        bad(1);
	if (nt > 0) { FI_CUDA(cudaMemcpyAsync(t.data(), h_trips, nt * sizeof(fi_triplet), cudaMemcpyHostToDevice, s)); }
	if (n > 0) { FI_CUDA(cudaMemcpyAsync(x.data(), h_x, n * sizeof(float), cudaMemcpyHostToDevice, s)); }
	if (nrows > 0) { FI_CUDA(cudaMemcpyAsync(err.data(), h_rhs, nrows * sizeof(float), cudaMemcpyHostToDevice, s)); }
	sumsq.zero(s);
	heat.zero(s);
	bad.zero(s);
	if (nt > 0) {
		FI_LAUNCH(rows_accumulate_kernel, div_up(nt, kThreads), kThreads, 0, s, nt, t.data(), x.data(), n, nrows, err.data(), sumsq.data(), bad.data());
	}
	if (nrows > 0) { FI_LAUNCH(square_kernel, div_up(nrows, kThreads), kThreads, 0, s, nrows, err.data()); }
	if (nt > 0) { FI_LAUNCH(blame_kernel, div_up(nt, kThreads), kThreads, 0, s, nt, t.data(), n, nrows, err.data(), sumsq.data(), heat.data()); }
	int h_bad = 0;
	FI_CUDA(cudaMemcpyAsync(&h_bad, bad.data(), sizeof(int), cudaMemcpyDeviceToHost, s));
	if (n > 0) { FI_CUDA(cudaMemcpyAsync(h_out, heat.data(), n * sizeof(float), cudaMemcpyDeviceToHost, s)); }
	FI_CUDA(cudaStreamSynchronize(s));
	FI_REQUIRE(h_bad == 0, FI_ERR_INVALID, "triplet row or column out of range");
}

}  // namespace fi
