"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
ibis_b200/lib/libibis_b200.so, against the golden vectors produced by the real reference
(tests/golden/make_golden.py) -- and, when the compiled reference travelled to the box
(oracle/_ref), against the reference itself on larger seeded inputs."""
import os
import sys

import numpy as np
import pytest

import cases
import parity

pytestmark = pytest.mark.gpu


def _require_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_case_against_golden(name):
    _require_cuda()
    parity.run_case(name)


def test_native_library_is_the_one_loaded():
    _require_cuda()
    from ibis_b200.lib import Library, DEFAULT_LIB
    L = Library.load()
    assert os.path.samefile(L.path, DEFAULT_LIB)
    with open("/proc/self/maps") as f:
        assert "libibis_b200.so" in f.read()


def test_launch_counter_counts_kernels():
    _require_cuda()
    from ibis_b200.lib import Context
    mesh, config, g, prim = cases.build_case("box3d_ausmdv")
    ctx = Context(g, config)
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, prim)
    n0 = ctx.launch_count()
    ctx.compute_dudt(fs, dq)
    assert ctx.launch_count() - n0 == 3       # ghost fill, pass G, pass F
    ctx.close()


@pytest.mark.gpu
def test_decomposed_run_equals_whole_box_on_two_gpus():
    """tools/mgpu_check.py under torchrun: two ranks, peer-memory halo exchange overlapped with the
    interior tiles of pass G', bit-identical to the undecomposed run (skipped on a one-GPU box)."""
    _require_cuda()
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for mode in ("p2p", "nccl"):
        env = dict(os.environ, IBIS_B200_HALO=mode)
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29533",
                              os.path.join(root, "tools", "mgpu_check.py"), "--edge", "32", "--steps", "2"],
                             capture_output=True, text=True, timeout=600, env=env)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert '"ok": true' in out.stdout
