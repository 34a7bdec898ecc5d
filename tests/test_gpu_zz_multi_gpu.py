"""GPU, two or more devices: the z-slab sharded solves against the single-GPU solve of the same system — Jacobi-PCG
iterate for iterate over peer memory, multigrid-preconditioned CG through the sharded V-cycle — by running
scripts/slab_check.py --quick under torchrun on 2 ranks.  Skipped on a one-GPU box (the round-end driver's): there the
multi-rank logic is covered by tests/test_emu_kernels.py on the CPU emulator, and the full check runs from
scripts/gpu_round12_multi.sh."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_slab_solves_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(ROOT, "scripts", "slab_check.py"), "--quick"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "SLAB CHECK PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
