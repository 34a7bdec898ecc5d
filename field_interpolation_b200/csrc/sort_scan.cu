// Device-wide exclusive scan and a stable LSD radix sort (8-bit digits), hand-written for the assembly
// stage: points are sorted by the key of the lattice cell that contains them so that the scatter kernel can
// reduce contributions inside a warp before touching memory (north-star item (a)).
#include "internal.hpp"

namespace fi {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems   = 8;
constexpr int kScanTile    = kScanThreads * kScanItems;

// Block-local exclusive scan of a tile; writes the tile total to sums[blockIdx.x].
__global__ void scan_tiles(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint64_t* __restrict__ sums,
                           int64_t n)
{
	__shared__ uint64_t warp_tot[kScanThreads / 32];
	const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
	uint64_t v[kScanItems];
	uint64_t run = 0;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		const int64_t k = base + i;
		const uint64_t x = k < n ? in[k] : 0;
		v[i] = run;
		run += x;
	}
	// warp-inclusive scan of per-thread totals
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint64_t  inc = run;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint64_t y = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) { inc += y; }
	}
	if (lane == 31) { warp_tot[wid] = inc; }
	__syncthreads();
	uint64_t warp_off = 0;
	for (int w = 0; w < wid; ++w) { warp_off += warp_tot[w]; }
	const uint64_t thread_off = warp_off + inc - run;
#pragma unroll
	for (int i = 0; i < kScanItems; ++i) {
		const int64_t k = base + i;
		if (k < n) { out[k] = v[i] + thread_off; }
	}
	if (threadIdx.x == kScanThreads - 1) { sums[blockIdx.x] = thread_off + run; }
}

__global__ void scan_add_offsets(uint64_t* __restrict__ out, const uint64_t* __restrict__ offs, int64_t n)
{
	const uint64_t off = offs[blockIdx.x];
	const int64_t  base = static_cast<int64_t>(blockIdx.x) * kScanTile;
	for (int i = threadIdx.x; i < kScanTile; i += kScanThreads) {
		const int64_t k = base + i;
		if (k < n) { out[k] += off; }
	}
}

__global__ void scan_total(const uint64_t* in_last, const uint64_t* out_last, uint64_t* total)
{
	*total = *in_last + *out_last;
}

void scan_rec(const uint64_t* in, uint64_t* out, int64_t n, cudaStream_t s, std::vector<DevBuf<uint64_t>>& keep)
{
	const int nb = div_up(n, kScanTile);
	keep.emplace_back(static_cast<size_t>(nb));
	uint64_t* sums = keep.back().data();
	FI_LAUNCH(scan_tiles, nb, kScanThreads, 0, s, in, out, sums, n);
	if (nb > 1) {
		scan_rec(sums, sums, nb, s, keep);
		FI_LAUNCH(scan_add_offsets, nb, kScanThreads, 0, s, out, sums, n);
	}
}

}  // namespace

void exclusive_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, uint64_t* total_dev, cudaStream_t s)
{
	if (n <= 0) {
		if (total_dev) { FI_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(uint64_t), s)); }
		return;
	}
	// `in` may alias `out`: the total needs in[n-1] before it is overwritten.
	DevBuf<uint64_t> last_in(1);
	if (total_dev) { FI_CUDA(cudaMemcpyAsync(last_in.data(), in + (n - 1), sizeof(uint64_t), cudaMemcpyDeviceToDevice, s)); }
	std::vector<DevBuf<uint64_t>> keep;
	scan_rec(in, out, n, s, keep);
	if (total_dev) { FI_LAUNCH(scan_total, 1, 1, 0, s, last_in.data(), out + (n - 1), total_dev); }
	FI_CUDA(cudaStreamSynchronize(s));  // scratch buffers die here
}

// ---- radix sort ----------------------------------------------------------------------------------------
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortRounds  = 16;
constexpr int kSortTile    = kSortThreads * kSortRounds;

__global__ void radix_hist(const uint64_t* __restrict__ keys, int64_t n, int shift, uint64_t* __restrict__ hist, int nblocks)
{
	__shared__ unsigned bins[256];
	bins[threadIdx.x] = 0;
	__syncthreads();
	const int64_t base = static_cast<int64_t>(blockIdx.x) * kSortTile;
	for (int r = 0; r < kSortRounds; ++r) {
		const int64_t k = base + r * kSortThreads + threadIdx.x;
		if (k < n) { atomicAdd(&bins[(keys[k] >> shift) & 255u], 1u); }
	}
	__syncthreads();
	hist[static_cast<size_t>(threadIdx.x) * nblocks + blockIdx.x] = bins[threadIdx.x];  // digit-major
}

// Stable scatter: the tile is consumed in rounds of 256 consecutive keys; inside a round the rank of a key
// among equal digits is (earlier warps' counts) + (earlier lanes of the same warp, via match_any).
__global__ void radix_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, int shift,
                              const uint64_t* __restrict__ offsets, int nblocks, uint64_t* __restrict__ keys_out,
                              uint32_t* __restrict__ vals_out)
{
	__shared__ uint64_t       bin_base[256];
	__shared__ unsigned short warp_cnt[kSortThreads / 32][256];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	bin_base[threadIdx.x] = offsets[static_cast<size_t>(threadIdx.x) * nblocks + blockIdx.x];
#pragma unroll
	for (int w = 0; w < kSortThreads / 32; ++w) { warp_cnt[w][threadIdx.x] = 0; }
	__syncthreads();
	const int64_t base = static_cast<int64_t>(blockIdx.x) * kSortTile;
	for (int r = 0; r < kSortRounds; ++r) {
		const int64_t k      = base + r * kSortThreads + threadIdx.x;
		const bool    active = k < n;
		uint64_t      key    = 0;
		unsigned      digit  = 0, rank = 0;
		const unsigned amask = __ballot_sync(0xffffffffu, active);
		if (active) {
			key   = keys[k];
			digit = static_cast<unsigned>((key >> shift) & 255u);
			const unsigned peers = __match_any_sync(amask, digit);
			rank                 = __popc(peers & ((1u << lane) - 1u));
			if (rank == 0) { warp_cnt[wid][digit] = static_cast<unsigned short>(__popc(peers)); }
		}
		__syncthreads();
		if (active) {
			uint64_t dst = bin_base[digit] + rank;
			for (int w = 0; w < wid; ++w) { dst += warp_cnt[w][digit]; }
			keys_out[dst] = key;
			vals_out[dst] = vals[k];
		}
		__syncthreads();
		{
			unsigned tot = 0;
#pragma unroll
			for (int w = 0; w < kSortThreads / 32; ++w) {
				tot += warp_cnt[w][threadIdx.x];
				warp_cnt[w][threadIdx.x] = 0;
			}
			bin_base[threadIdx.x] += tot;
		}
		__syncthreads();
	}
}

}  // namespace

void radix_sort_pairs(DevBuf<uint64_t>& keys, DevBuf<uint32_t>& vals, int64_t n, int bits, cudaStream_t s)
{
	if (n <= 1) { return; }
	const int        nblocks = div_up(n, kSortTile);
	DevBuf<uint64_t> keys2(static_cast<size_t>(n));
	DevBuf<uint32_t> vals2(static_cast<size_t>(n));
	DevBuf<uint64_t> hist(static_cast<size_t>(256) * nblocks);
	uint64_t *       ka = keys.data(), *kb = keys2.data();
	uint32_t *       va = vals.data(), *vb = vals2.data();
	const int        passes = std::max(1, (bits + 7) / 8);
	for (int p = 0; p < passes; ++p) {
		const int shift = 8 * p;
		FI_LAUNCH(radix_hist, nblocks, kSortThreads, 0, s, ka, n, shift, hist.data(), nblocks);
		exclusive_scan_u64(hist.data(), hist.data(), static_cast<int64_t>(256) * nblocks, nullptr, s);
		FI_LAUNCH(radix_scatter, nblocks, kSortThreads, 0, s, ka, va, n, shift, hist.data(), nblocks, kb, vb);
		std::swap(ka, kb);
		std::swap(va, vb);
	}
	if (ka != keys.data()) {
		FI_CUDA(cudaMemcpyAsync(keys.data(), ka, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
		FI_CUDA(cudaMemcpyAsync(vals.data(), va, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
	}
	FI_CUDA(cudaStreamSynchronize(s));
}

}  // namespace fi
