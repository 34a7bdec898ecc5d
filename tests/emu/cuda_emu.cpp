// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See cuda_emu.hpp.  Fiber scheduler and the slice of the CUDA runtime API the
// library's host code calls, over plain host memory.
#include "cuda_emu.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <memory>
#include <vector>

namespace cuda_emu {

uint3 g_threadIdx, g_blockIdx;
dim3  g_blockDim, g_gridDim;

namespace {

constexpr size_t kStack = 256 * 1024;

// Minimal x86-64 System V context switch (callee-saved registers + stack pointer); ucontext's swapcontext makes a
// sigprocmask system call per switch, which dominates the emulation of collectives.
extern "C" void cuda_emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl cuda_emu_switch
.type cuda_emu_switch,@function
cuda_emu_switch:
	pushq %rbp
	pushq %rbx
	pushq %r12
	pushq %r13
	pushq %r14
	pushq %r15
	movq %rsp, (%rdi)
	movq %rsi, %rsp
	popq %r15
	popq %r14
	popq %r13
	popq %r12
	popq %rbx
	popq %rbp
	ret
.size cuda_emu_switch,.-cuda_emu_switch
)");

enum State { kRunnable, kAtBlockBarrier, kAtWarpBarrier, kDone };

struct Warp
{
	uint64_t slot[32];
	unsigned alive    = 0;   // lanes that have not returned
	unsigned waiting  = 0;   // lanes parked at a warp barrier ...
	unsigned want[32] = {};  // ... and the member mask each of them named
};

struct Fiber
{
	void* sp    = nullptr;
	char* stack = nullptr;
	State state = kRunnable;
	uint3 tid;
	int   lane = 0, warp = 0;
};

void*                        g_sched_sp = nullptr;
Fiber*                       g_cur = nullptr;
const std::function<void()>* g_body = nullptr;
std::vector<Fiber>           g_fibers;
std::vector<Warp>            g_warps;
std::vector<char*>           g_stacks;  // reused across launches
int                          g_alive = 0, g_arrived = 0;

void to_scheduler() { cuda_emu_switch(&g_cur->sp, g_sched_sp); }

void trampoline()
{
	(*g_body)();
	g_cur->state = kDone;
	to_scheduler();
	std::abort();  // a finished fiber is never resumed
}

void release_block_barrier_if_complete()
{
	if (g_arrived > 0 && g_arrived >= g_alive) {
		g_arrived = 0;
		for (Fiber& f : g_fibers) {
			if (f.state == kAtBlockBarrier) { f.state = kRunnable; }
		}
	}
}

// A warp barrier over member mask M completes once every live lane of M is parked with the same mask (lanes that
// returned from the kernel no longer count, as on the hardware).  Disjoint groups of one warp complete independently.
void release_warp_barriers_if_complete(int warp)
{
	Warp& w = g_warps[warp];
	for (int l = 0; l < 32; ++l) {
		if (!(w.waiting >> l & 1u)) { continue; }
		const unsigned group = w.want[l] & w.alive;
		if ((w.waiting & group) != group) { continue; }
		bool same = true;
		for (int k = 0; k < 32; ++k) {
			if ((group >> k & 1u) && (w.want[k] & w.alive) != group) { same = false; }
		}
		if (!same) { continue; }
		w.waiting &= ~group;
		for (int k = 0; k < 32; ++k) {
			if (group >> k & 1u) { g_fibers[warp * 32 + k].state = kRunnable; }
		}
	}
}

void run_block(dim3 block)
{
	const int n = static_cast<int>(block.x * block.y * block.z);
	g_fibers.assign(n, Fiber{});
	while (static_cast<int>(g_stacks.size()) < n) { g_stacks.push_back(static_cast<char*>(std::malloc(kStack))); }
	g_warps.assign((n + 31) / 32, Warp{});
	g_alive   = n;
	g_arrived = 0;
	for (int i = 0; i < n; ++i) {
		Fiber& f = g_fibers[i];
		f.stack  = g_stacks[i];
		f.tid    = uint3{static_cast<unsigned>(i % block.x), static_cast<unsigned>((i / block.x) % block.y), static_cast<unsigned>(i / (block.x * block.y))};
		f.lane   = i & 31;
		f.warp   = i >> 5;
		g_warps[f.warp].alive |= 1u << f.lane;
		// initial frame: six callee-saved registers, then the entry point as the return address; at entry rsp = 16 n - 8
		uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + kStack) & ~uintptr_t{15};
		void**    sp  = reinterpret_cast<void**>(top);
		*--sp = nullptr;
		*--sp = reinterpret_cast<void*>(&trampoline);
		for (int r = 0; r < 6; ++r) { *--sp = nullptr; }
		f.sp = sp;
	}
	// CUDA_EMU_SCHED=reverse|shuffle: the order in which runnable threads of a block are resumed.  Results that change with
	// it point at a race or a missing barrier (the hardware promises no order between warps).
	static const char* sched_env = std::getenv("CUDA_EMU_SCHED");
	static const int   sched     = !sched_env ? 0 : (std::strcmp(sched_env, "reverse") == 0 ? 1 : 2);
	static unsigned    lcg       = 12345u;
	std::vector<int>   order(n);
	for (int i = 0; i < n; ++i) { order[i] = sched == 1 ? n - 1 - i : i; }
	while (g_alive > 0) {
		bool ran = false;
		if (sched == 2) {
			for (int i = n - 1; i > 0; --i) {
				lcg = lcg * 1664525u + 1013904223u;
				std::swap(order[i], order[(lcg >> 8) % (i + 1)]);
			}
		}
		for (int oi = 0; oi < n; ++oi) {
			const int i = order[oi];
			Fiber& f = g_fibers[i];
			if (f.state != kRunnable) { continue; }
			ran         = true;
			g_cur       = &f;
			g_threadIdx = f.tid;
			cuda_emu_switch(&g_sched_sp, f.sp);
			if (f.state == kDone) {
				--g_alive;
				g_warps[f.warp].alive &= ~(1u << f.lane);
				release_warp_barriers_if_complete(f.warp);
				release_block_barrier_if_complete();
			}
		}
		if (!ran) {
			std::fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %d threads alive, %d at the block barrier, the rest at warp barriers\n",
			             g_blockIdx.x, g_blockIdx.y, g_blockIdx.z, g_alive, g_arrived);
			std::abort();
		}
	}
}

struct Graph
{
	std::vector<std::function<void()>> nodes;
};
std::map<cudaStream_t, Graph*> g_capturing;

void run_grid(dim3 grid, dim3 block, const std::function<void()>& body)
{
	g_gridDim  = grid;
	g_blockDim = block;
	g_body     = &body;
	for (unsigned z = 0; z < grid.z; ++z) {
		for (unsigned y = 0; y < grid.y; ++y) {
			for (unsigned x = 0; x < grid.x; ++x) {
				g_blockIdx = uint3{x, y, z};
				run_block(block);
			}
		}
	}
	g_body = nullptr;
}

std::vector<unsigned char> g_dyn_smem;

// The emulator's own layout inside the opaque CUtensorMap (written by encode_tiled below).
struct EmuTensorMap
{
	unsigned long long magic;
	char*              base;
	unsigned           elem, rank;
	unsigned long long dim[3], stride_bytes[3];
	unsigned           box[3];
};
static_assert(sizeof(EmuTensorMap) <= sizeof(CUtensorMap), "emulated tensor map must fit the opaque one");
constexpr unsigned long long kMapMagic = 0x454d5554454e534full;

CUresult encode_tiled(CUtensorMap* out, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)
{
	if (rank != 3) { return CUDA_ERROR_INVALID_VALUE; }
	EmuTensorMap m{};
	m.magic = kMapMagic;
	m.base  = static_cast<char*>(base);
	m.elem  = type == CU_TENSOR_MAP_DATA_TYPE_FLOAT64 ? 8u : 4u;
	m.rank  = rank;
	for (int d = 0; d < 3; ++d) {
		m.dim[d]          = dims[d];
		m.box[d]          = box[d];
		m.stride_bytes[d] = d == 0 ? m.elem : strides[d - 1];
	}
	// the hardware's requirements that the library's geometry checks rely on
	if (reinterpret_cast<uintptr_t>(base) % 16 != 0 || m.stride_bytes[1] % 16 != 0 || m.stride_bytes[2] % 16 != 0 || box[0] * m.elem % 16 != 0 ||
	    box[0] > 256 || box[1] > 256 || box[2] > 256) {
		return CUDA_ERROR_INVALID_VALUE;
	}
	std::memset(out, 0, sizeof(*out));
	std::memcpy(out, &m, sizeof(m));
	return CUDA_SUCCESS;
}

}  // namespace

CUresult encode_tiled_entry(CUtensorMap* out, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                            const cuuint32_t* box, const cuuint32_t* es, CUtensorMapInterleave a, CUtensorMapSwizzle b, CUtensorMapL2promotion c,
                            CUtensorMapFloatOOBfill d)
{
	return encode_tiled(out, type, rank, base, dims, strides, box, es, a, b, c, d);
}

unsigned char* dynamic_smem() { return g_dyn_smem.data(); }

void spin_yield() { to_scheduler(); }

size_t tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z)
{
	EmuTensorMap m;
	std::memcpy(&m, map, sizeof(m));
	if (m.magic != kMapMagic) {
		std::fprintf(stderr, "cuda_emu: tma_load_3d through a tensor map the emulator did not encode\n");
		std::abort();
	}
	char* out = static_cast<char*>(dst);
	for (unsigned bz = 0; bz < m.box[2]; ++bz) {
		for (unsigned by = 0; by < m.box[1]; ++by) {
			for (unsigned bx = 0; bx < m.box[0]; ++bx, out += m.elem) {
				const long long cx = static_cast<long long>(x) + bx, cy = static_cast<long long>(y) + by, cz = static_cast<long long>(z) + bz;
				const bool inside = cx >= 0 && cy >= 0 && cz >= 0 && cx < static_cast<long long>(m.dim[0]) && cy < static_cast<long long>(m.dim[1]) &&
				                    cz < static_cast<long long>(m.dim[2]);
				if (inside) {
					std::memcpy(out, m.base + cx * m.stride_bytes[0] + cy * m.stride_bytes[1] + cz * m.stride_bytes[2], m.elem);
				} else {
					std::memset(out, 0, m.elem);
				}
			}
		}
	}
	return static_cast<size_t>(m.box[0]) * m.box[1] * m.box[2] * m.elem;
}

void launch(const char* name, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, std::function<void()> body)
{
	static const bool trace = std::getenv("CUDA_EMU_TRACE") != nullptr && *std::getenv("CUDA_EMU_TRACE") != '\0';
	auto run = [name, grid, block, smem, body]() {
		if (trace) { std::fprintf(stderr, "cuda_emu: %s <<<(%u,%u,%u),(%u,%u,%u),%zu>>>\n", name, grid.x, grid.y, grid.z, block.x, block.y, block.z, smem); }
		if (g_dyn_smem.size() < smem + 256) { g_dyn_smem.resize(smem + 256); }
		run_grid(grid, block, body);
	};
	auto it = g_capturing.find(stream);
	if (it != g_capturing.end()) {
		it->second->nodes.push_back(run);
		return;
	}
	run();
}

int lane_id() { return g_cur->lane; }

void block_barrier()
{
	++g_arrived;
	g_cur->state = kAtBlockBarrier;
	release_block_barrier_if_complete();
	if (g_cur->state != kRunnable) { to_scheduler(); }
}

void warp_barrier(unsigned mask)
{
	Warp&          w   = g_warps[g_cur->warp];
	const unsigned bit = 1u << g_cur->lane;
	w.want[g_cur->lane] = mask | bit;
	w.waiting |= bit;
	g_cur->state = kAtWarpBarrier;
	release_warp_barriers_if_complete(g_cur->warp);
	if (g_cur->state != kRunnable) { to_scheduler(); }
}

uint64_t warp_exchange(unsigned mask, uint64_t bits, int src)
{
	Warp& w             = g_warps[g_cur->warp];
	w.slot[g_cur->lane] = bits;
	warp_barrier(mask);
	const uint64_t r = (src >= 0 && src < 32) ? w.slot[src] : bits;
	warp_barrier(mask);
	return r;
}

unsigned warp_gather(unsigned mask, uint64_t bits, uint64_t out[32])
{
	Warp& w             = g_warps[g_cur->warp];
	w.slot[g_cur->lane] = bits;
	warp_barrier(mask);
	const unsigned part = (mask | 1u << g_cur->lane) & w.alive;
	for (int l = 0; l < 32; ++l) { out[l] = w.slot[l]; }
	warp_barrier(mask);
	return part;
}

}  // namespace cuda_emu

// ---- the CUDA runtime calls the library makes, over host memory -------------------------------------------------------
using cuda_emu::g_capturing;
using cuda_emu::Graph;

extern "C" {

// Device memory is host memory.  With CUDA_EMU_IPC=1 (the emulated ranks of a multi-process test) every allocation is a
// named shared-memory mapping, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can hand it to another process the
// way CUDA IPC hands device memory to a peer GPU's process.
struct Mapping
{
	std::string name;  // empty: opened from a peer (not ours to unlink)
	size_t      bytes;
};
static std::map<void*, Mapping> g_mappings;
static bool ipc_enabled()
{
	static const bool on = std::getenv("CUDA_EMU_IPC") != nullptr && *std::getenv("CUDA_EMU_IPC") != '\0';
	return on;
}
static void unlink_all_mappings()
{
	for (auto& m : g_mappings) {
		if (!m.second.name.empty()) { shm_unlink(m.second.name.c_str()); }
	}
}

// CUDA_EMU_GUARD=1: every allocation ends (to 16 bytes) at an inaccessible page and starts after one, so a kernel that
// indexes outside its buffer faults at the offending access — the emulator's stand-in for compute-sanitizer memcheck.
static bool guard_enabled()
{
	static const bool on = std::getenv("CUDA_EMU_GUARD") != nullptr && *std::getenv("CUDA_EMU_GUARD") != '\0';
	return on;
}
struct Guarded
{
	void*  map;
	size_t map_bytes;
};
static std::map<void*, Guarded> g_guarded;

cudaError_t cudaMalloc(void** p, size_t n)
{
	if (guard_enabled() && !ipc_enabled()) {
		const size_t page = 4096, body = ((std::max<size_t>(n, 1) + 15) & ~size_t{15});
		const size_t inner = (body + page - 1) & ~(page - 1), total = inner + 2 * page;
		char* m = static_cast<char*>(mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
		if (m == MAP_FAILED) { return cudaErrorMemoryAllocation; }
		mprotect(m, page, PROT_NONE);
		mprotect(m + page + inner, page, PROT_NONE);
		char* user = m + page + inner - body;  // the buffer's end touches the trailing guard page
		std::memset(m + page, 0xCB, inner - body);  // slack before the buffer: poison (reads of it yield garbage, not zeros)
		std::memset(user, 0xFF, body);              // fresh device memory is not zero: NaN for floats, -1 for integers
		g_guarded[user] = Guarded{m, total};
		*p = user;
		return cudaSuccess;
	}
	if (!ipc_enabled()) {
		*p = std::malloc(n ? n : 1);
		return *p ? cudaSuccess : cudaErrorMemoryAllocation;
	}
	static int  counter = 0;
	static bool hooked  = false;
	if (!hooked) {
		std::atexit(unlink_all_mappings);
		hooked = true;
	}
	char name[48];
	std::snprintf(name, sizeof(name), "/fi_emu_%d_%d", static_cast<int>(getpid()), counter++);
	const size_t bytes = (std::max<size_t>(n, 1) + 4095) & ~size_t{4095};
	const int    fd    = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
	if (fd < 0 || ftruncate(fd, static_cast<off_t>(bytes)) != 0) { return cudaErrorMemoryAllocation; }
	void* q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (q == MAP_FAILED) { return cudaErrorMemoryAllocation; }
	if (guard_enabled()) { std::memset(q, 0xFF, bytes); }  // fresh device memory is not zero
	g_mappings[q] = Mapping{name, bytes};
	*p            = q;
	return cudaSuccess;
}
cudaError_t cudaFree(void* p)
{
	auto gi = g_guarded.find(p);
	if (gi != g_guarded.end()) {
		munmap(gi->second.map, gi->second.map_bytes);
		g_guarded.erase(gi);
		return cudaSuccess;
	}
	auto it = g_mappings.find(p);
	if (it == g_mappings.end()) {
		std::free(p);
		return cudaSuccess;
	}
	munmap(p, it->second.bytes);
	if (!it->second.name.empty()) { shm_unlink(it->second.name.c_str()); }
	g_mappings.erase(it);
	return cudaSuccess;
}
struct EmuIpcHandle
{
	char               name[48];
	unsigned long long bytes;
};
static_assert(sizeof(EmuIpcHandle) <= sizeof(cudaIpcMemHandle_t), "emulated IPC handle must fit the opaque one");
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p)
{
	auto it = g_mappings.find(p);
	if (it == g_mappings.end() || it->second.name.empty()) { return cudaErrorNotSupported; }
	EmuIpcHandle e{};
	std::snprintf(e.name, sizeof(e.name), "%s", it->second.name.c_str());
	e.bytes = it->second.bytes;
	std::memset(h, 0, sizeof(*h));
	std::memcpy(h, &e, sizeof(e));
	return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned)
{
	EmuIpcHandle e;
	std::memcpy(&e, &h, sizeof(e));
	const int fd = shm_open(e.name, O_RDWR, 0600);
	if (fd < 0) { return cudaErrorInvalidValue; }
	void* q = mmap(nullptr, e.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (q == MAP_FAILED) { return cudaErrorInvalidValue; }
	g_mappings[q] = Mapping{"", static_cast<size_t>(e.bytes)};
	*p            = q;
	return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p)
{
	auto it = g_mappings.find(p);
	if (it == g_mappings.end()) { return cudaErrorInvalidValue; }
	munmap(p, it->second.bytes);
	g_mappings.erase(it);
	return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }

static void on_stream(cudaStream_t s, std::function<void()> op)
{
	auto it = g_capturing.find(s);
	if (it != g_capturing.end()) {
		it->second->nodes.push_back(std::move(op));
	} else {
		op();
	}
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s)
{
	on_stream(s, [=]() { std::memmove(dst, src, n); });
	return cudaSuccess;
}
cudaError_t cudaMemcpy(void* dst, const void* src, size_t n, cudaMemcpyKind)
{
	std::memmove(dst, src, n);
	return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* dst, int v, size_t n, cudaStream_t s)
{
	on_stream(s, [=]() { std::memset(dst, v, n); });
	return cudaSuccess;
}
cudaError_t cudaMemset(void* dst, int v, size_t n)
{
	std::memset(dst, v, n);
	return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "cuda_emu"; }
cudaError_t cudaGetDevice(int* d)
{
	*d = 0;
	return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n)
{
	*n = 1;
	return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int)
{
	*v = (a == cudaDevAttrMultiProcessorCount) ? 4 : 0;  // a small "GPU": grids sized from the SM count stay cheap to emulate
	return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int)
{
	std::memset(p, 0, sizeof(*p));
	std::snprintf(p->name, sizeof(p->name), "cuda_emu (CPU functional emulator)");
	p->major                       = 10;
	p->multiProcessorCount         = 4;
	p->sharedMemPerBlockOptin      = 227 * 1024;
	p->sharedMemPerBlock           = 48 * 1024;
	p->sharedMemPerMultiprocessor  = 228 * 1024;
	p->maxThreadsPerBlock          = 1024;
	p->maxThreadsPerMultiProcessor = 2048;
	p->totalGlobalMem              = size_t{8} << 30;
	return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaGetDriverEntryPointByVersion(const char* symbol, void** fn, unsigned, unsigned long long, cudaDriverEntryPointQueryResult* qr)
{
	const bool known = std::strcmp(symbol, "cuTensorMapEncodeTiled") == 0;
	*fn = known ? reinterpret_cast<void*>(&cuda_emu::encode_tiled_entry) : nullptr;
	if (qr) { *qr = known ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound; }
	return cudaSuccess;
}
cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total)
{
	*free_b = *total = size_t{8} << 30;
	return cudaSuccess;
}
cudaError_t cudaStreamCreate(cudaStream_t* s)
{
	*s = reinterpret_cast<cudaStream_t>(std::malloc(8));
	return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t s)
{
	std::free(s);
	return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e)
{
	*e = reinterpret_cast<cudaEvent_t>(std::malloc(8));
	return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e)
{
	std::free(e);
	return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t)
{
	*ms = 0.0f;
	return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

// graphs: a capture records the launches and copies of the stream; a launch replays them
cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode)
{
	g_capturing[s] = new Graph();
	return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t s, cudaGraph_t* g)
{
	auto it = g_capturing.find(s);
	*g      = it == g_capturing.end() ? nullptr : reinterpret_cast<cudaGraph_t>(it->second);
	if (it != g_capturing.end()) { g_capturing.erase(it); }
	return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long)
{
	*e = reinterpret_cast<cudaGraphExec_t>(new Graph(*reinterpret_cast<Graph*>(g)));
	return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t)
{
	for (auto& node : reinterpret_cast<Graph*>(e)->nodes) { node(); }
	return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e)
{
	delete reinterpret_cast<Graph*>(e);
	return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t g)
{
	delete reinterpret_cast<Graph*>(g);
	return cudaSuccess;
}

}  // extern "C"
