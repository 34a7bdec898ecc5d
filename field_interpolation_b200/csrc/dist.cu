// Multi-GPU: one process per GPU, the 3D lattice partitioned into contiguous z slabs.
//
// x is the fastest axis and z the slowest (reference field_interpolation.hpp:104-111), so a z plane is one
// contiguous run of nx*ny values and a halo is a plain contiguous range: no pack kernels.  Each slab stores its
// owned planes plus R halo planes either side (R = highest active model order).  Per PCG iteration:
//   * the fused direction+stencil kernel recomputes the new direction on the halo planes from r, M^-1 and p_old
//     (bit-identical to what the neighbour computes for the same planes), so only r's halo is exchanged:
//     R planes to each neighbour (ncclSend / ncclRecv in one group, in stream order, captured in the CUDA graph);
//   * two scalar all-reduces: p.Ap, and (r.M^-1 r, r.r) together.
// The data term needs no exchange: every slab keeps the cell blocks of all cells that touch a plane it owns
// (points are filtered by floor(z) on the device) and applies only the rows of nodes it owns.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2: the copy PyTorch already mapped when the caller is a
// torchrun rank, else the system one), so single-GPU users of libfi_b200.so do not need it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "solver.hpp"

namespace fi {

namespace {

struct Nccl
{
	void* lib = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                            = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                      = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t)                                                                = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)        = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)   = nullptr;
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)               = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                     = nullptr;
	ncclResult_t (*GroupStart)()                                                                           = nullptr;
	ncclResult_t (*GroupEnd)()                                                                             = nullptr;
	const char* (*GetErrorString)(ncclResult_t)                                                            = nullptr;
	std::string error;
};

Nccl& nccl()
{
	static Nccl           n;
	static std::once_flag once;
	std::call_once(once, [] {
		for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
			n.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (n.lib) { break; }
		}
		if (!n.lib) {
			n.error = std::string("cannot load libnccl: ") + dlerror();
			return;
		}
		auto sym = [&](const char* s) {
			void* p = dlsym(n.lib, s);
			if (!p && n.error.empty()) { n.error = std::string("libnccl lacks ") + s; }
			return p;
		};
		n.GetUniqueId    = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
		n.CommInitRank   = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
		n.CommDestroy    = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
		n.AllReduce      = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
		n.AllGather      = reinterpret_cast<decltype(n.AllGather)>(sym("ncclAllGather"));
		n.Broadcast      = reinterpret_cast<decltype(n.Broadcast)>(sym("ncclBroadcast"));
		n.Send           = reinterpret_cast<decltype(n.Send)>(sym("ncclSend"));
		n.Recv           = reinterpret_cast<decltype(n.Recv)>(sym("ncclRecv"));
		n.GroupStart     = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
		n.GroupEnd       = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
		n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
	});
	FI_REQUIRE(n.error.empty(), FI_ERR_COMM, n.error);
	return n;
}

#define FI_NCCL(expr)                                                                                         \
	do {                                                                                                      \
		ncclResult_t fi_r_ = (expr);                                                                          \
		if (fi_r_ != ncclSuccess) {                                                                           \
			throw ::fi::Error{FI_ERR_COMM, std::string(#expr) + ": " + nccl().GetErrorString(fi_r_)};         \
		}                                                                                                     \
	} while (0)

}  // namespace

void slab_range(int nz, int world, int rank, int* z0, int* z1)
{
	// contiguous, balanced to within one plane
	*z0 = static_cast<int>(static_cast<int64_t>(nz) * rank / world);
	*z1 = static_cast<int>(static_cast<int64_t>(nz) * (rank + 1) / world);
}

}  // namespace fi

struct fi_comm
{
	ncclComm_t comm  = nullptr;
	int        rank  = 0;
	int        world = 1;
	// peer-memory path (CUDA IPC over NVLink): mailboxes for the scalar sums, one shared lattice vector for r
	bool               p2p = false;
	fi::PeerLink       link;
	void*              shared       = nullptr;  // this rank's peer-visible vector
	size_t             shared_bytes = 0;
	void*              shared_lo    = nullptr;  // rank - 1's vector, mapped here
	void*              shared_hi    = nullptr;  // rank + 1's vector, mapped here
	unsigned long long seq          = 0;
	// non-uniform z partition of an nz = cuts_nz lattice (fi_comm_set_slab_cuts); empty: slab_range
	std::vector<int>   cuts;
	int                cuts_nz      = 0;

	void close_shared()
	{
		if (shared_lo) { cudaIpcCloseMemHandle(shared_lo); }
		if (shared_hi) { cudaIpcCloseMemHandle(shared_hi); }
		shared_lo = shared_hi = nullptr;
	}
	~fi_comm()
	{
		close_shared();
		if (shared) { cudaFree(shared); }
		for (int j = 0; j < world && j < fi::kMaxPeers; ++j) {
			if (j != rank && link.peer[j]) { cudaIpcCloseMemHandle(link.peer[j]); }
		}
		if (link.local) { cudaFree(link.local); }
		if (comm) { fi::nccl().CommDestroy(comm); }
	}
};

namespace fi {

namespace {

// Gathers one CUDA IPC handle per rank (collective, through the NCCL communicator).
std::vector<cudaIpcMemHandle_t> gather_handles(fi_comm* c, void* local_ptr, cudaStream_t s)
{
	cudaIpcMemHandle_t mine;
	std::memset(&mine, 0, sizeof(mine));
	if (local_ptr) { FI_CUDA(cudaIpcGetMemHandle(&mine, local_ptr)); }
	DevBuf<unsigned char> send(sizeof(mine)), recv(sizeof(mine) * c->world);
	FI_CUDA(cudaMemcpyAsync(send.data(), &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
	FI_NCCL(nccl().AllGather(send.data(), recv.data(), sizeof(mine), ncclChar, c->comm, s));
	std::vector<cudaIpcMemHandle_t> all(c->world);
	FI_CUDA(cudaMemcpyAsync(all.data(), recv.data(), sizeof(mine) * c->world, cudaMemcpyDeviceToHost, s));
	FI_CUDA(cudaStreamSynchronize(s));
	return all;
}

// All ranks agree (min over ranks) on a yes / no.
bool all_agree(fi_comm* c, bool mine, cudaStream_t s)
{
	DevBuf<double> v(1);
	const double   h = mine ? 1.0 : 0.0;
	FI_CUDA(cudaMemcpyAsync(v.data(), &h, sizeof(h), cudaMemcpyHostToDevice, s));
	FI_NCCL(nccl().AllReduce(v.data(), v.data(), 1, ncclDouble, ncclMin, c->comm, s));
	double out = 0;
	FI_CUDA(cudaMemcpyAsync(&out, v.data(), sizeof(out), cudaMemcpyDeviceToHost, s));
	FI_CUDA(cudaStreamSynchronize(s));
	return out > 0.5;
}

// Mailboxes: allocate, exchange IPC handles, map every peer's.  Leaves c->p2p false when any rank cannot.
void setup_peer_link(fi_comm* c)
{
	const char* env = getenv("FI_B200_P2P");
	bool        ok  = !(env && *env == '0') && c->world <= kMaxPeers && c->world > 1;
	cudaStream_t s = nullptr;
	FI_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	try {
		void* mb = nullptr;
		if (ok && cudaMalloc(&mb, sizeof(Mailbox)) != cudaSuccess) {
			cudaGetLastError();
			ok = false;
			mb = nullptr;
		}
		if (mb) { FI_CUDA(cudaMemset(mb, 0, sizeof(Mailbox))); }
		FI_CUDA(cudaDeviceSynchronize());
		c->link.rank  = c->rank;
		c->link.world = c->world;
		c->link.local = static_cast<Mailbox*>(mb);
		const auto handles = gather_handles(c, mb, s);
		for (int j = 0; ok && j < c->world; ++j) {
			if (j == c->rank) {
				c->link.peer[j] = c->link.local;
				continue;
			}
			void* q = nullptr;
			if (cudaIpcOpenMemHandle(&q, handles[j], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
				cudaGetLastError();
				ok = false;
				break;
			}
			c->link.peer[j] = static_cast<Mailbox*>(q);
		}
		c->p2p = all_agree(c, ok, s);
	} catch (...) {
		cudaStreamDestroy(s);
		throw;
	}
	cudaStreamDestroy(s);
}

// Halo exchange + scalar all-reduce of one slab: NCCL for the one-off steps, peer memory inside the iteration.
struct SlabHooks final : DistHooks
{
	fi_comm* c;
	Geom     g;
	int      halo;
	SlabHooks(fi_comm* comm, const Geom& geom, int h) : c(comm), g(geom), halo(h) {}

	const PeerLink* link() override { return c->p2p ? &c->link : nullptr; }
	unsigned long long next_seq() override { return ++c->seq; }
	int rank() const override { return c->rank; }
	int world() const override { return c->world; }
	std::unique_ptr<DistHooks> for_geom(const Geom& geom, int h) override { return std::make_unique<SlabHooks>(c, geom, h); }

	void allgather_planes(float* full, int64_t plane_cells, const std::vector<std::pair<int, int>>& own, cudaStream_t s) override
	{
		FI_REQUIRE(static_cast<int>(own.size()) == c->world, FI_ERR_INVALID, "allgather_planes: one plane range per rank");
		if (c->world == 1) { return; }
		Nccl&     n     = nccl();
		const int each  = own[0].second - own[0].first;
		bool      equal = true;
		for (int k = 0; k < c->world; ++k) { equal = equal && own[k].first == k * each && own[k].second == (k + 1) * each; }
		if (equal) {  // in place: this rank's planes already sit at full + rank * count
			const size_t count = static_cast<size_t>(each) * static_cast<size_t>(plane_cells);
			FI_NCCL(n.AllGather(full + static_cast<size_t>(c->rank) * count, full, count, ncclFloat, c->comm, s));
		} else {  // ragged: one in-place broadcast per owner, grouped
			FI_NCCL(n.GroupStart());
			for (int k = 0; k < c->world; ++k) {
				float*       at    = full + static_cast<size_t>(own[k].first) * static_cast<size_t>(plane_cells);
				const size_t count = static_cast<size_t>(own[k].second - own[k].first) * static_cast<size_t>(plane_cells);
				FI_NCCL(n.Broadcast(at, at, count, ncclFloat, k, c->comm, s));
			}
			FI_NCCL(n.GroupEnd());
		}
		count_launch();
	}
	int64_t peer_own_cells(int peer) override
	{
		int z0 = 0, z1 = 0;
		comm_slab_range(c, g.size[2], peer, &z0, &z1);
		return static_cast<int64_t>(z1 - z0) * g.stride[2];
	}

	void* shared_vector(size_t bytes, void** lo, void** hi, cudaStream_t s) override
	{
		// grow collectively: every rank runs the same sequence of solves, but slabs differ by a plane, so the decision
		// is taken on the largest request
		DevBuf<double> need(1);
		const double   mine = static_cast<double>(bytes);
		FI_CUDA(cudaMemcpyAsync(need.data(), &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
		FI_NCCL(nccl().AllReduce(need.data(), need.data(), 1, ncclDouble, ncclMax, c->comm, s));
		double most = 0;
		FI_CUDA(cudaMemcpyAsync(&most, need.data(), sizeof(most), cudaMemcpyDeviceToHost, s));
		FI_CUDA(cudaStreamSynchronize(s));
		const size_t want = static_cast<size_t>(most);
		if (want > c->shared_bytes) {
			FI_CUDA(cudaDeviceSynchronize());
			c->close_shared();
			(void)gather_handles(c, nullptr, s);  // barrier: nobody maps the old vectors any more
			if (c->shared) { FI_CUDA(cudaFree(c->shared)); }
			c->shared       = nullptr;
			c->shared_bytes = 0;
			const size_t cap = want + want / 8;
			FI_CUDA(cudaMalloc(&c->shared, cap));
			c->shared_bytes = cap;
			const auto handles = gather_handles(c, c->shared, s);
			if (c->rank > 0) { FI_CUDA(cudaIpcOpenMemHandle(&c->shared_lo, handles[c->rank - 1], cudaIpcMemLazyEnablePeerAccess)); }
			if (c->rank + 1 < c->world) { FI_CUDA(cudaIpcOpenMemHandle(&c->shared_hi, handles[c->rank + 1], cudaIpcMemLazyEnablePeerAccess)); }
		}
		FI_CUDA(cudaMemsetAsync(c->shared, 0, bytes, s));
		*lo = c->shared_lo;
		*hi = c->shared_hi;
		return c->shared;
	}

	void allreduce(double* d_ptr, int count, cudaStream_t s) override
	{
		FI_NCCL(nccl().AllReduce(d_ptr, d_ptr, static_cast<size_t>(count), ncclDouble, ncclSum, c->comm, s));
		count_launch();
	}

	void exchange_halo(void* d_vec, size_t elem, cudaStream_t s) override
	{
		if (c->world == 1 || halo == 0) { return; }
		Nccl&        n     = nccl();
		const size_t plane = static_cast<size_t>(g.stride[2]) * elem;  // bytes per plane
		char*        base  = static_cast<char*>(d_vec);
		const size_t bytes = plane * halo;
		// owned planes are [zown0, zown1); lower halo [zown0 - halo, zown0), upper halo [zown1, zown1 + halo)
		FI_NCCL(n.GroupStart());
		if (c->rank > 0) {
			FI_NCCL(n.Send(base + plane * g.zown0, bytes, ncclChar, c->rank - 1, c->comm, s));
			FI_NCCL(n.Recv(base + plane * (g.zown0 - halo), bytes, ncclChar, c->rank - 1, c->comm, s));
		}
		if (c->rank + 1 < c->world) {
			FI_NCCL(n.Send(base + plane * (g.zown1 - halo), bytes, ncclChar, c->rank + 1, c->comm, s));
			FI_NCCL(n.Recv(base + plane * g.zown1, bytes, ncclChar, c->rank + 1, c->comm, s));
		}
		FI_NCCL(n.GroupEnd());
		count_launch();
	}
};

template <typename T>
void slab_solve_typed(fi_comm* c, const Geom& g, int halo, const ModelAccum& model, const PointStore& pts, const fi_solve_options& o,
                      const float* d_guess_own, float* d_out_own, fi_solve_stats* st, cudaStream_t s)
{
	HostRows  none;
	auto      op = build_operator<T>(g, model, pts, none, s);
	SlabHooks hooks(c, g, halo);
	op->dist     = &hooks;
	op->use_fast = kStencilAuto;
	// the fused kernel evaluates p = M^-1 r + beta p_old on the halo planes too: it needs the neighbours' M^-1 there
	hooks.exchange_halo(op->minv.data(), sizeof(T), s);
	const int64_t off = g.own_offset(), n = g.own_cells();
	DevBuf<T>     x(g.N);
	x.zero(s);
	if (d_guess_own) {
		if (std::is_same<T, float>::value) {
			FI_CUDA(cudaMemcpyAsync(x.data() + off, d_guess_own, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
		} else {
			convert(d_guess_own, reinterpret_cast<double*>(x.data()) + off, n, s);
		}
	}
	op->guess_is_zero = d_guess_own == nullptr;
	const PcgResult r = pcg_solve<T>(*op, nullptr, x.data(), o.tolerance, o.max_iterations, o.check_every, true, s);
	if (std::is_same<T, float>::value) {
		FI_CUDA(cudaMemcpyAsync(d_out_own, x.data() + off, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
	} else {
		convert(reinterpret_cast<const double*>(x.data()) + off, d_out_own, n, s);
	}
	FI_CUDA(cudaStreamSynchronize(s));
	if (st) {
		std::memset(st, 0, sizeof(*st));
		st->iterations        = r.iterations;
		st->relative_residual = r.rel_residual;
		st->true_residual     = r.true_residual;
		st->initial_residual  = r.initial_residual;
		st->setup_ms          = op->setup_ms;
		st->solve_ms          = r.solve_ms;
		st->converged         = r.converged ? 1 : 0;
		st->occupied_cells    = op->data.nocc;
		st->generic_rows      = op->data.nrows;
		st->widened_after     = -1;
	}
}


// Multigrid-preconditioned CG on the slab (mg.cu: SlabMultigrid): the V-cycle always runs in fp32; the outer CG in fp32
// (`wide` false) — continued in fp64 from the fp32 iterate when that stalls at its rounding floor, as on one GPU
// (abi.cu: solve_device) — or in fp64 from the start.
void slab_mg_solve(fi_comm* c, const Geom& g, const SlabMgPlan& plan, const ModelAccum& model, const PointStore& pts, const fi_solve_options& o,
                   bool wide, const float* d_guess_own, float* d_out_own, fi_solve_stats* st, cudaStream_t s)
{
	CudaEvent e0, e1;
	e0.record(s);
	HostRows  none;
	SlabHooks hooks(c, g, plan.halo);
	auto      op32 = build_operator<float>(g, model, pts, none, s);
	op32->dist     = &hooks;
	op32->use_fast = kStencilAuto;
	MgOptions mo = default_mg_options(model, g, true, o.mg_smoothing_steps, o.mg_cheb_ratio);
	mg_options_from_env(mo);
	mo.gamma = 1;  // the sharded levels run V-cycles only
	auto mg = build_slab_multigrid(*op32, model, pts, mo, plan, s);
	std::unique_ptr<Operator<double>> op64;  // the outer CG's operator when it is not the V-cycle's
	auto wide_operator = [&]() -> Operator<double>& {
		if (!op64) {
			op64           = build_operator<double>(g, model, pts, none, s);
			op64->dist     = &hooks;
			op64->use_fast = kStencilAuto;
		}
		return *op64;
	};
	if (wide) { wide_operator(); }
	e1.record(s);
	const int64_t off = g.own_offset(), n = g.own_cells();
	PcgResult     r;
	long long     widened_after = -1;
	if (!wide) {
		DevBuf<float> x(g.N);
		x.zero(s);
		if (d_guess_own) { FI_CUDA(cudaMemcpyAsync(x.data() + off, d_guess_own, n * sizeof(float), cudaMemcpyDeviceToDevice, s)); }
		r = slab_mgpcg_solve<float>(*op32, *mg, nullptr, x.data(), o.tolerance, o.max_iterations, s, 4);
		FI_CUDA(cudaMemcpyAsync(d_out_own, x.data() + off, n * sizeof(float), cudaMemcpyDeviceToDevice, s));
		// every rank sees the same all-reduced sums, so every rank takes the same decision
		const bool budget_left = o.max_iterations <= 0 || r.iterations < o.max_iterations;
		if (!r.converged && !r.zero_rhs && r.stalled && budget_left) {
			widened_after = r.iterations;
			wide          = true;
		}
	}
	if (wide) {
		Operator<double>& op = wide_operator();
		DevBuf<double>    x(g.N);
		x.zero(s);
		const float* from = widened_after >= 0 ? d_out_own : d_guess_own;
		if (from) { convert(from, x.data() + off, n, s); }
		const long long cap = o.max_iterations > 0 ? o.max_iterations - std::max<long long>(widened_after, 0) : 0;
		const PcgResult w   = slab_mgpcg_solve<double>(op, *mg, nullptr, x.data(), o.tolerance, cap, s);
		convert(x.data() + off, d_out_own, n, s);
		if (widened_after >= 0) {
			PcgResult tot        = w;
			tot.iterations       = r.iterations + w.iterations;
			tot.solve_ms         = r.solve_ms + w.solve_ms;
			tot.initial_residual = r.initial_residual;
			r                    = tot;
		} else {
			r = w;
		}
	}
	FI_CUDA(cudaStreamSynchronize(s));
	const double setup = e1.ms_since(e0);
	if (st) {
		std::memset(st, 0, sizeof(*st));
		st->iterations        = r.iterations;
		st->relative_residual = r.rel_residual;
		st->true_residual     = r.true_residual;
		st->initial_residual  = r.initial_residual;
		st->setup_ms          = setup;
		st->solve_ms          = r.solve_ms;
		st->converged         = r.converged ? 1 : 0;
		st->occupied_cells    = op32->data.nocc;
		st->generic_rows      = op32->data.nrows;
		st->widened_after     = widened_after;
	}
}

int64_t slab_mg_gather_cells()
{
	// levels with at most this many cells are replicated on every rank (default: up to 144^3)
	if (const char* e = getenv("FI_B200_MG_GATHER_CELLS")) {
		if (atoll(e) > 0) { return atoll(e); }
	}
	return 3000000;
}

}  // namespace

// add_model_impl's arithmetic (abi.cu) for one Weights value
void accumulate_model(const fi_weights& w, ModelAccum& m);

void slab_sdf_solve(fi_comm* c, const int32_t* sizes, const fi_weights& w, int64_t num_points, const float* positions, const float* normals,
                    const float* point_weights, int loc, const fi_solve_options& o, const float* guess_own, float* solution_own,
                    int sol_loc, fi_solve_stats* st)
{
	TraceScope trace("fi_slab_sdf_solve");
	FI_REQUIRE(c != nullptr && c->comm != nullptr, FI_ERR_INVALID, "communicator is null");
	FI_REQUIRE(o.precision == FI_F32 || o.precision == FI_F64, FI_ERR_UNSUPPORTED, "slab solves run in FI_F32 or FI_F64");
	FI_REQUIRE(w.gradient_smoothness == 0.0f, FI_ERR_UNSUPPORTED, "slab solves need the star-shaped operator (gradient_smoothness = 0)");
	FI_REQUIRE(normals == nullptr || w.gradient_kernel != FI_GRADIENT_LINEAR_INTERPOLATION, FI_ERR_UNSUPPORTED,
	           "slab solves keep the data term in cell blocks: linear-interpolation gradient rows are not supported");
	ModelAccum model;
	accumulate_model(w, model);
	int halo = 0;
	for (int k = 0; k <= 4; ++k) {
		if (model.on[k]) { halo = k; }
	}
	FI_REQUIRE(halo >= 1, FI_ERR_UNSUPPORTED, "slab solves need a smoothness order >= 1");
	const bool multigrid = o.preconditioner == FI_PRECOND_MULTIGRID;
	SlabMgPlan plan;
	if (multigrid) {  // the V-cycle's transfers want at least two halo planes
		plan = plan_slab_multigrid(sizes, c->world, halo, slab_mg_gather_cells(), comm_cuts(c, sizes[2]));
		halo = plan.halo;
	}
	int z0 = 0, z1 = 0;
	comm_slab_range(c, sizes[2], c->rank, &z0, &z1);
	FI_REQUIRE(z1 - z0 >= halo, FI_ERR_INVALID, "slabs thinner than the stencil radius: use fewer ranks");
	const Geom g = make_slab_geom(sizes, z0, z1, halo);

	cudaStream_t s = nullptr;
	FI_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
	try {
		const int      D = 3;
		DevBuf<float>  dpos, dnrm, dpw;
		const float *  ppos = positions, *pnrm = normals, *ppw = point_weights;
		if (loc == FI_HOST && num_points > 0) {
			dpos.resize(num_points * D);
			FI_CUDA(cudaMemcpyAsync(dpos.data(), positions, num_points * D * sizeof(float), cudaMemcpyHostToDevice, s));
			ppos = dpos.data();
			if (normals) {
				dnrm.resize(num_points * D);
				FI_CUDA(cudaMemcpyAsync(dnrm.data(), normals, num_points * D * sizeof(float), cudaMemcpyHostToDevice, s));
				pnrm = dnrm.data();
			}
			if (point_weights) {
				dpw.resize(num_points);
				FI_CUDA(cudaMemcpyAsync(dpw.data(), point_weights, num_points * sizeof(float), cudaMemcpyHostToDevice, s));
				ppw = dpw.data();
			}
		}
		PointStore pts;
		canonicalise_points(g, pts, w.data_pos, w.value_kernel, w.data_gradient, w.gradient_kernel, num_points, ppos, pnrm, ppw, nullptr, s);
		const int64_t n = g.own_cells();
		DevBuf<float> d_guess, d_out;
		const float*  gptr = guess_own;
		float*        optr = solution_own;
		if (sol_loc == FI_HOST) {
			d_out.resize(n);
			optr = d_out.data();
			if (guess_own) {
				d_guess.resize(n);
				FI_CUDA(cudaMemcpyAsync(d_guess.data(), guess_own, n * sizeof(float), cudaMemcpyHostToDevice, s));
				gptr = d_guess.data();
			}
		}
		if (multigrid) {
			slab_mg_solve(c, g, plan, model, pts, o, o.precision != FI_F32, gptr, optr, st, s);
		} else if (o.precision == FI_F32) {
			slab_solve_typed<float>(c, g, halo, model, pts, o, gptr, optr, st, s);
		} else {
			slab_solve_typed<double>(c, g, halo, model, pts, o, gptr, optr, st, s);
		}
		if (sol_loc == FI_HOST) {
			FI_CUDA(cudaMemcpyAsync(solution_own, d_out.data(), n * sizeof(float), cudaMemcpyDeviceToHost, s));
			FI_CUDA(cudaStreamSynchronize(s));
		}
	} catch (...) {
		cudaStreamDestroy(s);
		throw;
	}
	cudaStreamDestroy(s);
}

void comm_unique_id(void* id, int64_t capacity)
{
	FI_REQUIRE(id != nullptr && capacity >= static_cast<int64_t>(sizeof(ncclUniqueId)), FI_ERR_INVALID, "id buffer must hold 128 bytes");
	ncclUniqueId u;
	FI_NCCL(nccl().GetUniqueId(&u));
	std::memcpy(id, &u, sizeof(u));
}

void comm_destroy(fi_comm* c) { delete c; }

const int* comm_cuts(const fi_comm* c, int nz) { return (c && !c->cuts.empty() && c->cuts_nz == nz) ? c->cuts.data() : nullptr; }

void comm_slab_range(const fi_comm* c, int nz, int rank, int* z0, int* z1)
{
	if (const int* cuts = comm_cuts(c, nz)) {
		*z0 = cuts[rank];
		*z1 = cuts[rank + 1];
	} else {
		slab_range(nz, c->world, rank, z0, z1);
	}
}

void comm_set_cuts(fi_comm* c, int nz, const int32_t* cuts)
{
	FI_REQUIRE(c != nullptr, FI_ERR_INVALID, "communicator is null");
	if (cuts == nullptr) {
		c->cuts.clear();
		c->cuts_nz = 0;
		return;
	}
	FI_REQUIRE(nz >= c->world && cuts[0] == 0 && cuts[c->world] == nz, FI_ERR_INVALID, "slab cuts must run from 0 to nz");
	for (int k = 0; k < c->world; ++k) { FI_REQUIRE(cuts[k + 1] > cuts[k], FI_ERR_INVALID, "slab cuts must increase: every rank owns at least one plane"); }
	c->cuts.assign(cuts, cuts + c->world + 1);
	c->cuts_nz = nz;
}

namespace {

// points whose containing cell starts in plane z (the planes that pay for their cell block in the data-term kernels)
__global__ void plane_histogram_kernel(int64_t n, const float* __restrict__ pos, int nz, unsigned long long* __restrict__ hist)
{
	const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
	for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
		const float z = floorf(pos[3 * i + 2]);
		if (z >= 0.0f && z < static_cast<float>(nz)) { atomicAdd(&hist[static_cast<int>(z)], 1ull); }
	}
}

}  // namespace

void balanced_cuts(const int32_t* sizes, int world, int64_t num_points, const float* positions, int loc, double point_weight, int min_planes,
                   int32_t* cuts)
{
	const int nz = sizes[2];
	min_planes   = std::max(1, min_planes);
	FI_REQUIRE(world >= 1 && nz >= world * min_planes, FI_ERR_INVALID, "balanced cuts: fewer planes than ranks x min_planes");
	std::vector<unsigned long long> hist(nz, 0ull);
	if (num_points > 0 && point_weight > 0) {
		cudaStream_t                 s = nullptr;
		DevBuf<unsigned long long>   d_hist(nz);
		DevBuf<float>                staged;
		const float*                 d_pos = positions;
		d_hist.zero(s);
		if (loc == FI_HOST) {
			staged.resize(static_cast<size_t>(num_points) * 3);
			FI_CUDA(cudaMemcpyAsync(staged.data(), positions, static_cast<size_t>(num_points) * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
			d_pos = staged.data();
		}
		const int grid = static_cast<int>(std::min<int64_t>((num_points + 255) / 256, static_cast<int64_t>(sm_count()) * 8));
		FI_LAUNCH(plane_histogram_kernel, grid, 256, 0, s, num_points, d_pos, nz, d_hist.data());
		FI_CUDA(cudaMemcpyAsync(hist.data(), d_hist.data(), static_cast<size_t>(nz) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
		FI_CUDA(cudaStreamSynchronize(s));
	}
	// An iteration on slabs is two phases, each ended by a scalar exchange every rank waits in (solver.cu): A = stencil +
	// data term, B = the update kernel.  With c(z) = nx ny + point_weight * points(z) the work of plane z in phase A (in
	// lattice cells of stencil work) and kUpdateRatio * nx ny its work in phase B, the iteration takes
	//     max_rank A + max_rank B   (+ exchange latencies that do not depend on the cuts),
	// so the cuts minimise that sum: for every cap P on the planes of a slab (B = kUpdateRatio nx ny P) the smallest
	// reachable max A is found by bisection over a greedy left-to-right fill, and the best (P, max A) pair wins.
	// Measured on 8 B200s (512^3, r2e): stencil 0.94 us per 512^2 plane, update 1.15 us, data term 0.106 us per 1000 points
	// -> point_weight ~ 30 lattice cells per point, kUpdateRatio ~ 1.22.
	constexpr double kUpdateRatio = 1.22;
	const double        plane = static_cast<double>(sizes[0]) * sizes[1];
	std::vector<double> cum(nz + 1, 0.0);
	for (int z = 0; z < nz; ++z) { cum[z + 1] = cum[z] + plane + point_weight * static_cast<double>(hist[z]); }
	std::vector<int> trial(world + 1), best(world + 1);
	// greedy fill: every rank takes planes while its phase-A work stays <= target and its slab <= cap planes
	auto fill = [&](double target, int cap) {
		trial[0] = 0;
		for (int k = 0; k < world; ++k) {
			const int z0   = trial[k];
			const int room = nz - z0 - (world - 1 - k) * min_planes;  // leave min_planes for every rank still to come
			int       z1   = std::min(z0 + std::min(cap, room), nz);
			if (z1 - z0 < min_planes) { return false; }
			// largest z1 in (z0, z1] with cum[z1] - cum[z0] <= target
			const int hi = static_cast<int>(std::upper_bound(cum.begin() + z0 + 1, cum.begin() + z1 + 1, cum[z0] + target) - cum.begin()) - 1;
			z1           = std::min(z1, hi);
			if (z1 - z0 < min_planes) { return false; }
			trial[k + 1] = z1;
		}
		return trial[world] == nz;
	};
	double best_cost = -1.0;
	for (int cap = (nz + world - 1) / world; cap <= nz - (world - 1) * min_planes; ++cap) {
		double lo = cum[nz] / world, hi = cum[nz];  // max A lies between the perfect split and everything on one rank
		if (!fill(hi, cap)) { continue; }
		for (int it = 0; it < 60 && hi - lo > 0.25 * plane * 1e-3; ++it) {
			const double mid = 0.5 * (lo + hi);
			if (fill(mid, cap)) { hi = mid; } else { lo = mid; }
		}
		if (!fill(hi, cap)) { continue; }
		double max_a = 0.0;
		int    max_p = 0;
		for (int k = 0; k < world; ++k) {
			max_a = std::max(max_a, cum[trial[k + 1]] - cum[trial[k]]);
			max_p = std::max(max_p, trial[k + 1] - trial[k]);
		}
		const double cost = max_a + kUpdateRatio * plane * max_p;
		if (best_cost < 0.0 || cost < best_cost * (1.0 - 1e-12)) {
			best_cost = cost;
			best      = trial;
		}
	}
	FI_REQUIRE(best_cost >= 0.0, FI_ERR_INVALID, "balanced cuts: no partition satisfies min_planes");
	for (int k = 0; k <= world; ++k) { cuts[k] = best[k]; }
}

fi_comm* comm_create(int rank, int world, const void* id)
{
	FI_REQUIRE(world >= 1 && rank >= 0 && rank < world && id != nullptr, FI_ERR_INVALID, "bad rank / world / id");
	ncclUniqueId u;
	std::memcpy(&u, id, sizeof(u));
	auto c   = std::make_unique<fi_comm>();
	c->rank  = rank;
	c->world = world;
	FI_NCCL(nccl().CommInitRank(&c->comm, world, u, rank));
	setup_peer_link(c.get());
	return c.release();
}

}  // namespace fi
