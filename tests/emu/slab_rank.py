"""TEST INFRASTRUCTURE.  One emulated rank of a z-slab sharded solve (run by tests/test_emu_slab.py, `world` of these in
parallel): loads the CPU functional emulator build of the library (tests/emu/_build/libfi_emu.so) in place of
libfi_b200.so — in THIS process only —, joins a communicator through the fake NCCL on LD_LIBRARY_PATH, runs
fi_slab_sdf_solve on the case named on the command line and saves its owned planes and solve statistics.

    python slab_rank.py <rank> <world> <workdir> <case-json>
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load_emulated_library():
    from field_interpolation_b200 import _lib
    dll = C.CDLL(os.path.join(HERE, "_build", "libfi_emu.so"))
    assert dll.fi_emu_marker() == 1
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(dll, name)
        fn.restype, fn.argtypes = res, args
    _lib._dll = dll
    return _lib


def main():
    rank, world, work = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    case = json.loads(sys.argv[4])
    L = load_emulated_library()
    import field_interpolation_b200 as fi
    from field_interpolation_b200 import dist as fid
    from field_interpolation_b200 import workloads as W

    id_path = os.path.join(work, "nccl_id.bin")
    if rank == 0:
        buf = (C.c_ubyte * 128)()
        L.check(L.lib().fi_comm_unique_id(buf, 128))
        with open(id_path + ".tmp", "wb") as f:
            f.write(bytes(buf))
        os.replace(id_path + ".tmp", id_path)
    t0 = time.time()
    while not os.path.exists(id_path):
        assert time.time() - t0 < 60, "rank 0 never published the communicator id"
        time.sleep(0.01)
    comm = fid.SlabComm(rank, world, open(id_path, "rb").read())

    sizes = case["sizes"]
    cloud = W.sphere_torus_3d(case["points"], seed=case.get("seed", 1))
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    weights = fi.Weights(**case.get("weights", {}))
    runner = fid.SlabRunner.__new__(fid.SlabRunner)
    runner.sizes, runner.weights, runner.rank, runner.world, runner.comm = sizes, weights, rank, world, comm
    cuts = case.get("cuts")  # None: uniform; "balanced": fi_slab_balanced_cuts of the cloud; else world + 1 plane numbers
    if cuts == "balanced":
        cuts = fid.balanced_cuts(sizes, world, pos, case.get("point_weight", 0.0), case.get("min_planes", 4))
    runner.set_cuts(cuts)
    results = {}
    normals = cloud["normals"] if case.get("normals", True) else None
    pw = np.random.default_rng(7).uniform(0.0, 2.0, len(pos)).astype(np.float32) if case.get("point_weights") else None
    nxy = sizes[0] * sizes[1]
    for name, o in case["solves"].items():
        opt = fi.solve_options(fi.FI_F64 if o["precision"] == "f64" else fi.FI_F32, o["max_iterations"], o["tolerance"],
                               preconditioner=fi.FI_PRECOND_MULTIGRID if o.get("multigrid") else fi.FI_PRECOND_JACOBI)
        out = np.zeros(runner.local_cells, np.float32)
        guess = None
        if o.get("guess"):  # this rank's planes of a smooth global guess
            zz, yy, xx = np.meshgrid(np.arange(runner.z0, runner.z1), np.arange(sizes[1]), np.arange(sizes[0]), indexing="ij")
            guess = (0.1 * np.sin(0.3 * xx) + 0.05 * yy - 0.02 * zz).astype(np.float32).ravel()
        st = runner.step(pos, normals, opt, out, guess=guess, point_weights=pw)
        np.save(os.path.join(work, f"{name}_rank{rank}.npy"), out)
        results[name] = st
    results["cuts"] = runner.cuts
    with open(os.path.join(work, f"stats_rank{rank}.json"), "w") as f:
        json.dump(results, f)
    runner.close()


if __name__ == "__main__":
    main()
