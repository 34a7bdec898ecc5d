// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  A functional emulator of the CUDA execution model for the build container,
// which has nvcc but no GPU: the library's .cu sources are compiled a second time with g++ (this header is
// force-included) into tests/emu/_build/libfi_emu.so, where every kernel launch runs block after block on the CPU, each
// thread of a block as a cooperatively scheduled fiber so that __syncthreads and the warp collectives
// (__shfl_*_sync, __ballot_sync, __match_any_sync, ...) have their real semantics.  "Device" memory is host memory.
//
// What it is for: catching LOGIC errors in kernels (indexing, scan offsets, reduction protocols, barriers that not every
// thread reaches) before a GPU box is available.  What it cannot show: data races (fibers switch only at collectives),
// memory-model bugs, performance, TMA / mbarrier / peer-memory paths (not emulated).  Nothing in the product loads this
// library; only tests/test_emu_*.py do, and those are not parity evidence for the CUDA build — the `-m gpu` tests are.
#pragma once

#define FI_B200_EMU 1
#define __host__
#define __device__
#define __global__
#define __shared__ static
#define __constant__ static
#define __grid_constant__

#include <cuda.h>           // CUtensorMap (opaque 128 bytes: the emulator keeps its own description in them)
#include <cuda_runtime.h>  // vector types, runtime API declarations (implemented in cuda_emu.cpp)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <tuple>

#undef __launch_bounds__
#define __launch_bounds__(...)

namespace cuda_emu {

extern uint3 g_threadIdx, g_blockIdx;
extern dim3  g_blockDim, g_gridDim;

// Runs body() once per thread of a grid x block launch.  When `stream` is being captured into a graph the launch is
// recorded instead and runs at cudaGraphLaunch.
void launch(const char* name, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, std::function<void()> body);
// The block's dynamic shared memory (`extern __shared__`), sized by the launch.
unsigned char* dynamic_smem();
// cp.async.bulk.tensor.3d: copies the box of `map` whose first element is (x, y, z) to dst, zero outside the tensor.
// Returns the bytes delivered.
size_t tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z);
// Inside a spin-wait on memory another thread of the block will write: lets the other threads run.
void spin_yield();
CUresult encode_tiled_entry(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void     block_barrier();
// The warp collectives synchronise the live lanes of `mask` (every one of which must make the same call).
void     warp_barrier(unsigned mask);
// Every calling lane deposits `bits`; returns lane `src`'s deposit (own if src is outside the warp).
uint64_t warp_exchange(unsigned mask, uint64_t bits, int src);
// Every calling lane deposits `bits`; out[l] receives lane l's deposit, the return value is the mask of lanes that took part.
unsigned warp_gather(unsigned mask, uint64_t bits, uint64_t out[32]);
int      lane_id();

template <typename T>
inline uint64_t to_bits(T v)
{
	static_assert(sizeof(T) <= 8, "shuffle of a type wider than 8 bytes");
	uint64_t b = 0;
	std::memcpy(&b, &v, sizeof(T));
	return b;
}
template <typename T>
inline T from_bits(uint64_t b)
{
	T v;
	std::memcpy(&v, &b, sizeof(T));
	return v;
}

}  // namespace cuda_emu

#define threadIdx (::cuda_emu::g_threadIdx)
#define blockIdx (::cuda_emu::g_blockIdx)
#define blockDim (::cuda_emu::g_blockDim)
#define gridDim (::cuda_emu::g_gridDim)
#define warpSize 32

// the library's launch macro (csrc/common.cuh defines its <<< >>> form only when this one is absent)
// (arguments are evaluated at the launch, like a real launch copies them, and travel inside the closure)
#define FI_LAUNCH(kernel, grid, block, smem, stream, ...)                                                                  \
	do {                                                                                                                   \
		::cuda_emu::launch(#kernel, dim3(grid), dim3(block), (smem), (stream), [fi_k_ = kernel, fi_a_ = std::make_tuple(__VA_ARGS__)]() {   \
			std::apply([&](auto&... a) { fi_k_(a...); }, fi_a_);                                                           \
		});                                                                                                                \
		::fi::count_launch();                                                                                              \
	} while (0)

using std::max;
using std::min;

// nvcc's cuda_runtime.h has a function-pointer overload of this one; the host-compiler view of the header does not
template <typename R, typename... A>
inline cudaError_t cudaFuncSetAttribute(R (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }
inline size_t __cvta_generic_to_shared(const void* p) { return reinterpret_cast<size_t>(p); }

inline void __syncthreads() { ::cuda_emu::block_barrier(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { ::cuda_emu::warp_barrier(mask); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }  // peers are other processes of this machine
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	const int lane = ::cuda_emu::lane_id(), base = lane & ~(width - 1);
	return ::cuda_emu::from_bits<T>(::cuda_emu::warp_exchange(mask, ::cuda_emu::to_bits(v), base + (src & (width - 1))));
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int o, int width = 32)
{
	const int lane = ::cuda_emu::lane_id(), src = lane ^ o;
	return ::cuda_emu::from_bits<T>(::cuda_emu::warp_exchange(mask, ::cuda_emu::to_bits(v), (src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane));
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	const int lane = ::cuda_emu::lane_id(), src = lane - static_cast<int>(d);
	return ::cuda_emu::from_bits<T>(::cuda_emu::warp_exchange(mask, ::cuda_emu::to_bits(v), src >= (lane & ~(width - 1)) ? src : lane));
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	const int lane = ::cuda_emu::lane_id(), src = lane + static_cast<int>(d);
	return ::cuda_emu::from_bits<T>(::cuda_emu::warp_exchange(mask, ::cuda_emu::to_bits(v), src < (lane & ~(width - 1)) + width ? src : lane));
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
	uint64_t       all[32];
	const unsigned part = ::cuda_emu::warp_gather(mask, pred ? 1u : 0u, all);
	unsigned       r    = 0;
	for (int l = 0; l < 32; ++l) {
		if ((part >> l & 1u) && (mask >> l & 1u) && all[l]) { r |= 1u << l; }
	}
	return r;
}
inline int __all_sync(unsigned mask, int pred)
{
	uint64_t       all[32];
	const unsigned part = ::cuda_emu::warp_gather(mask, pred ? 1u : 0u, all);
	for (int l = 0; l < 32; ++l) {
		if ((part >> l & 1u) && (mask >> l & 1u) && !all[l]) { return 0; }
	}
	return 1;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v)
{
	uint64_t       all[32];
	const uint64_t mine = ::cuda_emu::to_bits(v);
	const unsigned part = ::cuda_emu::warp_gather(mask, mine, all);
	unsigned       r    = 0;
	for (int l = 0; l < 32; ++l) {
		if ((part >> l & 1u) && (mask >> l & 1u) && all[l] == mine) { r |= 1u << l; }
	}
	return r;
}
inline unsigned __activemask() { return 0xffffffffu; }

template <typename T>
inline T atomicAdd(T* p, T v)
{
	const T old = *p;
	*p          = old + v;
	return old;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v)
{
	const unsigned long long old = *p;
	*p += v;
	return old;
}
template <typename T>
inline T atomicMax(T* p, T v)
{
	const T old = *p;
	*p          = std::max(old, v);
	return old;
}
template <typename T>
inline T atomicMin(T* p, T v)
{
	const T old = *p;
	*p          = std::min(old, v);
	return old;
}
template <typename T>
inline T atomicOr(T* p, T v)
{
	const T old = *p;
	*p          = old | v;
	return old;
}
template <typename T>
inline T atomicExch(T* p, T v)
{
	const T old = *p;
	*p          = v;
	return old;
}
template <typename T>
inline T atomicCAS(T* p, T cmp, T v)
{
	const T old = *p;
	if (old == cmp) { *p = v; }
	return old;
}

template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
template <typename T>
inline T __ldcs(const T* p) { return *p; }
template <typename T>
inline void __stcs(T* p, T v) { *p = v; }
template <typename T>
inline void __stcg(T* p, T v) { *p = v; }

inline int       __popc(unsigned v) { return __builtin_popcount(v); }
inline int       __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int       __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
inline int       __ffs(int v) { return __builtin_ffs(v); }
inline long long clock64() { return 0; }
inline float     __fmul_rn(float a, float b) { return a * b; }
inline float     __fadd_rn(float a, float b) { return a + b; }
inline float     __fsub_rn(float a, float b) { return a - b; }
inline float     __fdiv_rn(float a, float b) { return a / b; }
inline float     __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double    __dmul_rn(double a, double b) { return a * b; }
inline double    __dadd_rn(double a, double b) { return a + b; }
inline double    __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline float     __int_as_float(int v) { return ::cuda_emu::from_bits<float>(static_cast<uint32_t>(v)); }
inline int       __float_as_int(float v) { return static_cast<int>(::cuda_emu::to_bits(v)); }
inline float     rsqrtf(float v) { return 1.0f / std::sqrt(v); }
