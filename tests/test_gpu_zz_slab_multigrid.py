"""GPU (one device): FI_PRECOND_MULTIGRID through the slab path on a 1-rank communicator.  Kept in its own file, sorted
after the other GPU tests: the sharded V-cycle's fused smoother (TMA stencil epilogue on a slab window) was last changed
after round 1's GPU budget was spent — validated on the CPU emulator (tests/test_emu_kernels.py) with 1-3 ranks, first
GPU run at the round-end driver pass."""
import numpy as np
import pytest

from field_interpolation_b200 import workloads as W


@pytest.mark.gpu
@pytest.mark.parametrize("gather", [0, 1000])
def test_one_rank_slab_multigrid_matches_plain_multigrid(gather, monkeypatch):
    """FI_PRECOND_MULTIGRID through the slab path on a 1-rank communicator: the sharded V-cycle (unfused smoother over
    the owned planes, transfers through shifted slab pointers, replicated tail hierarchy) is the same linear operator
    as the single-GPU V-cycle, so the CG takes the same number of iterations (within rounding) to the same field.
    gather = 1000 forces two sharded levels at this size."""
    import field_interpolation_b200 as fi
    from field_interpolation_b200 import dist as fid

    class OneRank:
        @staticmethod
        def get_backend():
            return "gloo"

        @staticmethod
        def broadcast(t, src=0):
            return None

    if gather:
        monkeypatch.setenv("FI_B200_MG_GATHER_CELLS", str(gather))
    sizes = [64, 48, 40]
    assert fid.slab_mg_plan(sizes, 1, 2, gather)["sharded_levels"] == (2 if gather else 1)
    cloud = W.sphere_torus_3d(5000, seed=5)
    pos = W.to_lattice(cloud["unit_pos"], sizes)
    weights = fi.Weights()
    runner = fid.SlabRunner(sizes, weights, 0, 1, OneRank)
    f = fi.sdf_from_points(sizes, weights, pos, cloud["normals"])
    for prec, tol, close in ((fi.FI_F64, 1e-9, 1e-6), (fi.FI_F32, 1e-5, 2e-3)):
        opt = fi.solve_options(prec, 200, tol, preconditioner=fi.FI_PRECOND_MULTIGRID)
        out = np.zeros(runner.local_cells, np.float32)
        st = runner.step(pos, cloud["normals"], opt, out)
        ref, st1 = f.solve(opt)
        assert st["converged"] and st1["converged"]
        assert abs(st["iterations"] - st1["iterations"]) <= 2, (st, st1)
        assert st["true_residual"] <= 30 * max(tol, st1["true_residual"])
        assert np.linalg.norm(out - ref) <= close * np.linalg.norm(ref)
    runner.close()
