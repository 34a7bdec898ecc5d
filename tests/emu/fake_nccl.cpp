// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  A stand-in for libnccl.so.2 for the CPU functional emulator (cuda_emu.hpp):
// the handful of NCCL entry points csrc/dist.cu loads with dlopen, implemented between PROCESSES of one machine over a
// POSIX shared-memory segment, so that the z-slab sharded solves (halo exchange, scalar all-reduces, the multigrid
// all-gather) can be exercised with world_size > 1 in the GPU-less build container.  "Device" buffers are host memory
// and every call completes before it returns (the emulated streams are synchronous).  Built to
// tests/emu/_build/fake_nccl/libnccl.so.2; tests put that directory first on LD_LIBRARY_PATH of the rank processes.
#include <fcntl.h>
#include <nccl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr size_t kHeader = 4096;
constexpr size_t kSlot   = size_t{256} << 20;  // per-rank staging area (sparse: untouched pages cost nothing)
constexpr int    kMaxMsg = 256;

struct Header
{
	std::atomic<int> joined, bar_count, bar_sense;
};
struct Slot
{
	int    nmsg;
	struct { int dst; size_t off, bytes; } msg[kMaxMsg];
	size_t used;
	alignas(64) char data[1];
};

struct Comm
{
	int         rank = 0, world = 1, sense = 0;
	std::string name;
	char*       base = nullptr;
	Header*     hdr() { return reinterpret_cast<Header*>(base); }
	Slot*       slot(int r) { return reinterpret_cast<Slot*>(base + kHeader + static_cast<size_t>(r) * kSlot); }
	void        barrier()
	{
		sense ^= 1;
		if (hdr()->bar_count.fetch_add(1) == world - 1) {
			hdr()->bar_count.store(0);
			hdr()->bar_sense.store(sense);
		} else {
			while (hdr()->bar_sense.load() != sense) { sched_yield(); }
		}
	}
};

struct Op
{
	enum Kind { kSend, kRecv, kAllReduce, kAllGather, kBroadcast } kind;
	const void*    send;
	void*          recv;
	size_t         count;
	ncclDataType_t type;
	ncclRedOp_t    red;
	int            peer;  // send/recv peer, broadcast root
	Comm*          comm;
};

thread_local int             t_depth = 0;
thread_local std::vector<Op> t_queue;

size_t type_size(ncclDataType_t t)
{
	switch (t) {
		case ncclInt8: case ncclUint8: return 1;
		case ncclFloat16: return 2;
		case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
		default: return 8;
	}
}

char* stage(Comm* c, const void* src, size_t bytes, int dst)
{
	Slot* s = c->slot(c->rank);
	if (s->nmsg >= kMaxMsg || sizeof(Slot) + s->used + bytes > kSlot) {
		std::fprintf(stderr, "fake_nccl: staging area exhausted (%zu bytes)\n", bytes);
		std::abort();
	}
	s->msg[s->nmsg] = {dst, s->used, bytes};
	++s->nmsg;
	char* at = s->data + s->used;
	std::memcpy(at, src, bytes);
	s->used += (bytes + 63) & ~size_t{63};
	return at;
}
void reset(Comm* c)
{
	c->slot(c->rank)->nmsg = 0;
	c->slot(c->rank)->used = 0;
}

template <typename T>
void reduce_into(T* out, const T* in, size_t n, ncclRedOp_t red, bool first)
{
	for (size_t i = 0; i < n; ++i) {
		if (first) { out[i] = in[i]; }
		else if (red == ncclSum) { out[i] += in[i]; }
		else if (red == ncclMin) { out[i] = in[i] < out[i] ? in[i] : out[i]; }
		else if (red == ncclMax) { out[i] = in[i] > out[i] ? in[i] : out[i]; }
		else { std::abort(); }
	}
}

void run_collective(const Op& op)
{
	Comm*        c     = op.comm;
	const size_t bytes = op.count * type_size(op.type);
	reset(c);
	if (op.kind != Op::kBroadcast || c->rank == op.peer) { stage(c, op.send, bytes, -1); }
	c->barrier();
	if (op.kind == Op::kAllReduce) {
		for (int r = 0; r < c->world; ++r) {
			const char* in = c->slot(r)->data;
			if (op.type == ncclFloat64) { reduce_into(static_cast<double*>(op.recv), reinterpret_cast<const double*>(in), op.count, op.red, r == 0); }
			else if (op.type == ncclFloat32) { reduce_into(static_cast<float*>(op.recv), reinterpret_cast<const float*>(in), op.count, op.red, r == 0); }
			else { std::abort(); }
		}
	} else if (op.kind == Op::kAllGather) {
		for (int r = 0; r < c->world; ++r) { std::memcpy(static_cast<char*>(op.recv) + static_cast<size_t>(r) * bytes, c->slot(r)->data, bytes); }
	} else if (c->rank != op.peer || op.recv != op.send) {
		std::memcpy(op.recv, c->slot(op.peer)->data, bytes);
	}
	c->barrier();
}

// consecutive sends / receives of a group: everything is staged, then everything is delivered
void run_p2p(const Op* ops, size_t n)
{
	Comm* c = ops[0].comm;
	reset(c);
	for (size_t i = 0; i < n; ++i) {
		if (ops[i].kind == Op::kSend) { stage(c, ops[i].send, ops[i].count * type_size(ops[i].type), ops[i].peer); }
	}
	c->barrier();
	std::vector<int> taken(c->world, 0);  // messages from each source already consumed (matched in posting order)
	for (size_t i = 0; i < n; ++i) {
		if (ops[i].kind != Op::kRecv) { continue; }
		const Slot* s     = c->slot(ops[i].peer);
		int         seen  = 0;
		bool        found = false;
		for (int m = 0; m < s->nmsg && !found; ++m) {
			if (s->msg[m].dst != c->rank) { continue; }
			if (seen++ < taken[ops[i].peer]) { continue; }
			const size_t bytes = ops[i].count * type_size(ops[i].type);
			if (bytes != s->msg[m].bytes) {
				std::fprintf(stderr, "fake_nccl: rank %d expects %zu bytes from %d, which sent %zu\n", c->rank, bytes, ops[i].peer, s->msg[m].bytes);
				std::abort();
			}
			std::memcpy(ops[i].recv, s->data + s->msg[m].off, bytes);
			++taken[ops[i].peer];
			found = true;
		}
		if (!found) {
			std::fprintf(stderr, "fake_nccl: rank %d posted a receive from %d that nothing matches\n", c->rank, ops[i].peer);
			std::abort();
		}
	}
	c->barrier();
}

void flush()
{
	std::vector<Op> q;
	q.swap(t_queue);
	size_t i = 0;
	while (i < q.size()) {
		if (q[i].kind == Op::kSend || q[i].kind == Op::kRecv) {
			size_t j = i;
			while (j < q.size() && (q[j].kind == Op::kSend || q[j].kind == Op::kRecv)) { ++j; }
			run_p2p(&q[i], j - i);
			i = j;
		} else {
			run_collective(q[i]);
			++i;
		}
	}
}

ncclResult_t post(Op op)
{
	t_queue.push_back(op);
	if (t_depth == 0) { flush(); }
	return ncclSuccess;
}

}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId* id)
{
	std::memset(id, 0, sizeof(*id));
	std::snprintf(id->internal, sizeof(id->internal), "/fi_fake_nccl_%d_%ld", static_cast<int>(getpid()), static_cast<long>(random()));
	return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t* out, int world, ncclUniqueId id, int rank)
{
	Comm* c  = new Comm();
	c->rank  = rank;
	c->world = world;
	c->name  = std::string(id.internal, strnlen(id.internal, sizeof(id.internal)));
	const size_t bytes = kHeader + static_cast<size_t>(world) * kSlot;
	const int    fd    = shm_open(c->name.c_str(), O_CREAT | O_RDWR, 0600);
	if (fd < 0 || ftruncate(fd, static_cast<off_t>(bytes)) != 0) { return ncclSystemError; }
	c->base = static_cast<char*>(mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
	close(fd);
	if (c->base == MAP_FAILED) { return ncclSystemError; }
	c->hdr()->joined.fetch_add(1);
	while (c->hdr()->joined.load() < world) { sched_yield(); }
	*out = reinterpret_cast<ncclComm_t>(c);
	return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm)
{
	Comm* c = reinterpret_cast<Comm*>(comm);
	if (c->hdr()->joined.fetch_sub(1) == 1) { shm_unlink(c->name.c_str()); }  // the last one out
	munmap(c->base, kHeader + static_cast<size_t>(c->world) * kSlot);
	delete c;
	return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void* s, void* r, size_t n, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c, cudaStream_t)
{
	return post(Op{Op::kAllReduce, s, r, n, t, op, -1, reinterpret_cast<Comm*>(c)});
}
ncclResult_t ncclAllGather(const void* s, void* r, size_t n, ncclDataType_t t, ncclComm_t c, cudaStream_t)
{
	return post(Op{Op::kAllGather, s, r, n, t, ncclSum, -1, reinterpret_cast<Comm*>(c)});
}
ncclResult_t ncclBroadcast(const void* s, void* r, size_t n, ncclDataType_t t, int root, ncclComm_t c, cudaStream_t)
{
	return post(Op{Op::kBroadcast, s, r, n, t, ncclSum, root, reinterpret_cast<Comm*>(c)});
}
ncclResult_t ncclSend(const void* s, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t)
{
	return post(Op{Op::kSend, s, nullptr, n, t, ncclSum, peer, reinterpret_cast<Comm*>(c)});
}
ncclResult_t ncclRecv(void* r, size_t n, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t)
{
	return post(Op{Op::kRecv, nullptr, r, n, t, ncclSum, peer, reinterpret_cast<Comm*>(c)});
}
ncclResult_t ncclGroupStart()
{
	++t_depth;
	return ncclSuccess;
}
ncclResult_t ncclGroupEnd()
{
	if (--t_depth == 0) { flush(); }
	return ncclSuccess;
}
const char* ncclGetErrorString(ncclResult_t) { return "fake_nccl"; }

}  // extern "C"
