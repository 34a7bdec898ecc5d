// Peer-memory plumbing of the multi-GPU CG iteration: the device-resident CG state, the mailboxes the ranks of one
// NVSwitch domain exchange their partial sums through, and the device functions that publish to / collect from them.
#pragma once

#include "common.cuh"

namespace fi {

// Device-resident scalars of one PCG run; nothing here is read by the host inside the iteration loop except
// at convergence polls.
struct PcgState
{
	double rho[2];  // r.z, ping-pong by iteration parity
	double pq;      // p.(A p)
	double rr;      // r.r (recurrence residual)
	double bb;      // b.b
	double tol2bb;  // tolerance^2 * b.b  (Eigen's stopping rule: |r|^2 <= tol^2 |b|^2)
	double rr0;     // r.r of the initial guess
	int    done;
	int    breakdown;
	long long iters;
	long long max_iters;
	double part[4];  // multi-GPU: this rank's partial sums, all-reduced in place before the finish kernels read them
	double alpha_prev;  // step length of the last even-numbered iteration, whose x update is applied one iteration later (solver.cu)
};

// ---- multi-GPU: scalar all-reduce and halo push over NVLink peer memory, inside the CG kernels -------------
// Every rank owns a mailbox in device memory that its peers map through CUDA IPC.  A kernel publishes this
// rank's partial sums by storing {values, sequence number} into its slot of every peer's mailbox; the kernel
// that needs the sum spins on its local mailbox until all slots carry the expected sequence number and adds
// them in rank order (so every rank forms bit-identical sums).  Sequence numbers only grow: no resets, no ABA.
constexpr int kMaxPeers = 8;  // one NVSwitch domain

struct PeerSlot
{
	double             v[2];
	unsigned long long seq;
	unsigned long long pad;
};

struct Mailbox
{
	PeerSlot slot[2 /* which sum */][2 /* iteration parity */][kMaxPeers];
	int      error;  // set by a kernel that gave up waiting for a peer
	// device timestamps (ns) of the last 512 iterations, for FI_B200_TRACE: [0] p.Ap published (stencil + data term
	// done), [1] update kernel past its wait, [2] iteration finished (all ranks' r.r collected)
	unsigned long long stamp[3][512];
};

struct PeerLink  // passed to kernels by value
{
	int      rank = 0, world = 1;
	Mailbox* local = nullptr;
	Mailbox* peer[kMaxPeers] = {};  // peer[rank] == local
};

template <typename T>
struct HaloPush  // where the update kernel stores its boundary planes of r in the neighbours' copies of r
{
	T*      lo = nullptr;  // rank - 1's upper halo planes (receives my first `count` owned values), or null
	T*      hi = nullptr;  // rank + 1's lower halo planes (receives my last `count` owned values), or null
	int64_t count = 0;     // halo planes * cells per plane
};

// What a kernel that ends a reduction needs to publish this rank's sum itself (instead of a one-warp kernel after it).
struct PeerPublish
{
	PeerLink           link;
	int                which = 0, par = 0;
	unsigned long long base = 0;
	const PcgState*    st = nullptr;
};

#if defined(__CUDACC__) || defined(FI_B200_EMU)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
#ifdef FI_B200_EMU  // tests/emu: no PTX on the CPU functional emulator; a polling thread lets its siblings run (the hardware
	::cuda_emu::spin_yield();  // guarantees forward progress to the other lanes of a warp, sequential fibers do not)
	return *reinterpret_cast<const volatile unsigned long long*>(p);
#else
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
#endif
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
#ifdef FI_B200_EMU
	*reinterpret_cast<volatile unsigned long long*>(p) = v;
#else
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}

// The first warp of a block, all 32 lanes: lane j stores this rank's `count` (<= 2) partial sums and then the sequence
// number into rank j's mailbox — every peer in parallel, one NVLink round trip in all (a single thread doing the
// `world` release stores one after the other costs `world` round trips: 20+ us at 8 ranks, measured).
__device__ __forceinline__ void peer_publish_warp(const PeerLink& L, int which, int par, unsigned long long seq, const double* v, int count)
{
	const int j = threadIdx.x & 31;
	if (j < L.world) {
		PeerSlot* sl = &L.peer[j]->slot[which][par][L.rank];
		for (int k = 0; k < count; ++k) { reinterpret_cast<volatile double*>(sl->v)[k] = v[k]; }
		// cumulative: also orders the halo stores of the other blocks, observed through the ticket, before the flag
		__threadfence_system();
		st_release_sys(&sl->seq, seq);
	}
}

// The first warp of a block, all 32 lanes: lane j waits until rank j's slot carries `seq`; lane 0 then adds the values in
// rank order (every rank forms bit-identical sums).  Gives up after ~10 s (a peer died): flags the mailbox and returns
// false.  The result is valid in lane 0 (and returned to every lane).
__device__ __forceinline__ bool peer_collect_warp(const PeerLink& L, int which, int par, unsigned long long seq, double* out, int count)
{
	const int j  = threadIdx.x & 31;
	double    v0 = 0.0, v1 = 0.0;
	bool      ok = true;
	if (j < L.world) {
		const PeerSlot* sl = &L.local->slot[which][par][j];
		const long long t0 = clock64();
		while (ld_acquire_sys(&sl->seq) != seq) {
			if (clock64() - t0 > 20000000000ll) {
				L.local->error = 1;
				ok             = false;
				break;
			}
		}
		v0 = reinterpret_cast<const volatile double*>(sl->v)[0];
		if (count > 1) { v1 = reinterpret_cast<const volatile double*>(sl->v)[1]; }
	}
	ok = __all_sync(0xffffffffu, ok);
	double t0s = 0.0, t1s = 0.0;
	for (int r = 0; r < L.world; ++r) {
		t0s += __shfl_sync(0xffffffffu, v0, r);
		t1s += __shfl_sync(0xffffffffu, v1, r);
	}
	out[0] = t0s;
	if (count > 1) { out[1] = t1s; }
	return ok;
}

// Sequence numbers of iteration `iters` of the solve with epoch number `base`: 2 * iters + 1 for p.Ap, + 2 for
// (r.Mr, r.r).  Derived from the device-side iteration counter so that the kernels can sit in a CUDA graph.  Once
// the solve is done the counter stops and the leftover iterations of a round re-publish the same numbers, which
// every waiting peer accepts at once (their values are ignored: all ranks are done together).
__device__ __forceinline__ unsigned long long seq_of(unsigned long long base, const PcgState* st, int which)
{
	return base + 2ull * static_cast<unsigned long long>(st->iters) + 1ull + static_cast<unsigned long long>(which);
}

__device__ __forceinline__ unsigned long long global_ns()
{
#ifdef FI_B200_EMU
	return 0;
#else
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
#endif
}

#endif  // __CUDACC__ || FI_B200_EMU

}  // namespace fi
