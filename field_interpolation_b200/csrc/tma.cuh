// Pieces shared by the TMA-staged stencil kernels (stencil_tma.cu: 3D, stencil_2d.cu: 2D): 16-byte packs, the per-axis
// coefficient tables as a kernel argument, the multigrid epilogue's arguments, and the PTX wrappers for mbarriers and
// cp.async.bulk.tensor loads (with their stand-ins for the CPU functional emulator of tests/emu).
#pragma once

#include <cuda.h>

#include <mutex>

#include "internal.hpp"

namespace fi {
namespace tma {

__host__ __device__ __forceinline__ int row_class(int i, int n)
{
	return n <= 9 ? i : (i < 4 ? i : (i >= n - 4 ? i - n + 9 : 4));
}

template <typename T>
struct PackOf;
template <>
struct PackOf<float>
{
	using type = float4;
};
template <>
struct PackOf<double>
{
	using type = double2;
};

template <typename T, int V>
union PackU
{
	typename PackOf<T>::type v;
	T                        a[V];
};

template <typename T>
struct TmaTables
{
	T band[kMaxDim][9][9];
};

// Epilogue mode (the multigrid smoother, mg.cu): instead of storing q = S p the kernel consumes it in place,
//     res_out = res_in - q;   and, when d_new is given,   d_new = a p + b M^-1 res_out,   e += d_new,
// i.e. one Chebyshev step (or a plain residual update) per pass over the lattice.  The pointwise operands of a
// thread's own pack are prefetched one plane ahead into registers.  res_in / res_out and e are updated pointwise
// and may alias; d_new must not alias the stencil input.
template <typename T>
struct EpiArgs
{
	const T* res_in = nullptr;
	T*       res_out = nullptr;
	const T* minv = nullptr;
	T*       e = nullptr;
	T*       d_new = nullptr;
	T        a = 0, b = 0;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------
#ifdef FI_B200_EMU
// tests/emu (CPU functional emulator of the build container): a TMA load is a box copy with zero fill outside the
// tensor, performed when it is issued; the mbarrier is a word holding {bytes still expected, completed phases}, so a
// consumer that gets ahead of the producer thread still waits, as on the hardware.  Everything else runs as is.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(reinterpret_cast<uintptr_t>(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t) { *bar = 0; }
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { *bar += static_cast<uint64_t>(bytes) << 32; }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	while ((*reinterpret_cast<volatile uint64_t*>(bar) & 1u) == parity) { ::cuda_emu::spin_yield(); }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z)
{
	*bar -= static_cast<uint64_t>(::cuda_emu::tma_load_3d(dst, map, x, y, z)) << 32;
	if ((*bar >> 32) == 0) { *bar += 1; }  // all expected bytes have landed: the phase completes
}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "FI_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra FI_DONE;\n"
	    "bra FI_WAIT;\n"
	    "FI_DONE:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z)
{
	asm volatile(
	    "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
	        smem_u32(dst)),
	    "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
	    : "memory");
}
#endif  // FI_B200_EMU

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeFn encode_fn()
{
	static EncodeFn   fn = nullptr;
	static std::once_flag once;
	std::call_once(once, [] {
		void*                            p = nullptr;
		cudaDriverEntryPointQueryResult  qr;
		if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &qr) == cudaSuccess &&
		    qr == cudaDriverEntryPointSuccess) {
			fn = reinterpret_cast<EncodeFn>(p);
		}
	});
	return fn;
}

}  // namespace tma
}  // namespace fi
