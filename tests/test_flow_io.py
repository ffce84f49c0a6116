"""ibis_b200_flow_read_native / flow_write_native: the reference's native_binary snapshot arrays
(io/native.cpp:200-339: T, p, vx, vy[, vz] over the valid cells) in and out of a FlowStates buffer,
with NativeBinaryInput's update_thermo_from_pT (rho = p / (R T), e = Cv T).  Bit-exact: the closure is
two IEEE operations."""
import numpy as np
import pytest

import cases
from ibis_b200.lib import Context

R, CV = 287.0, 717.5


def _check(name, lib_path=None):
    mesh, config, g, prim = cases.build_case(name)
    n = g.n_cells
    rng = np.random.default_rng(7)
    T = np.ascontiguousarray(prim[:n, 2] * (1 + 0.01 * rng.uniform(-1, 1, n)))
    p = np.ascontiguousarray(prim[:n, 1] * (1 + 0.01 * rng.uniform(-1, 1, n)))
    vx, vy = np.ascontiguousarray(prim[:n, 4]), np.ascontiguousarray(prim[:n, 5])
    vz = np.ascontiguousarray(prim[:n, 6]) if mesh.dim == 3 else None
    ctx = Context(g, config, lib_path=lib_path)
    try:
        fs = ctx.fs_create()
        sentinel = np.full_like(prim, -7.0)
        ctx.fs_upload(fs, sentinel)
        ctx.flow_read_native(fs, T, p, vx, vy, vz)
        got = ctx.fs_download(fs)
        assert np.array_equal(got[n:], sentinel[n:]), "ghost rows must stay untouched"
        want = np.stack([p / (R * T), p, T, CV * T, vx, vy, vz if vz is not None else np.zeros(n)], axis=1)
        assert np.array_equal(got[:n], want)
        back = ctx.flow_write_native(fs)
        for a, b in zip(back, [T, p, vx, vy] + ([vz] if vz is not None else [])):
            assert np.array_equal(a, b)
        # the snapshot feeds the residual exactly like a full FlowStates upload of the same state
        dq1, dq2, fs2 = ctx.cq_create(), ctx.cq_create(), ctx.fs_create()
        full = sentinel.copy()
        full[:n] = want
        ctx.fs_upload(fs2, full)
        ctx.compute_dudt(fs, dq1, True)
        ctx.compute_dudt(fs2, dq2, True)
        assert np.array_equal(ctx.cq_download(dq1)[:n], ctx.cq_download(dq2)[:n])
    finally:
        ctx.close()


@pytest.mark.parametrize("name", ["box3d_ausmdv", "quad2d_ausmdv", "mixed3d_walls"])
def test_native_snapshot_roundtrip_emu(name):
    from emu_util import emu_lib
    _check(name, emu_lib())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["box3d_ausmdv", "quad2d_ausmdv", "mixed3d_walls"])
def test_native_snapshot_roundtrip_gpu(name):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _check(name)
