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


# ---- snapshot FILES through the ring (ibis_b200/native_io.py) against snapshots the reference itself wrote ----
import json
import os

DROPIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dropin")
FILE_CASES = ["uniform_stream", "uniform_stream_3D", "steady_state_uniform_stream", "example_forward_facing_step"]


def _case_context(case, lib_path):
    from ibis_b200 import lib as libmod
    with open(os.path.join(DROPIN, case, "config", "config.json")) as f:
        config = json.load(f)
    mesh = libmod.read_su2(os.path.join(DROPIN, case, "io", "grid", "0000", "block_0000.su2"), lib_path=lib_path)
    g = libmod.build_grid(mesh, None, lib_path=lib_path)
    return Context(g, config, lib_path=lib_path), config


def _check_files(case, tmp_path, lib_path=None):
    """The final snapshot of a reference run (arrays, meta_data.json text, the `flows` list: tests/golden/dropin,
    written by NativeBinaryOutput::write) goes into a FlowStates buffer through the upload half of the ring and comes
    back out through SnapshotWriter: every file must hold the reference's bytes."""
    from ibis_b200 import native_io
    exp = np.load(os.path.join(DROPIN, case, "expected.npz"))
    ctx, config = _case_context(case, lib_path)
    try:
        dim = ctx.dim
        names = native_io.NAMES3[:5 if dim == 3 else 4]
        fs = ctx.fs_create()
        # snapshot 0 is the one the reference's prep wrote: NativeBinaryInput::read + update_thermo_from_pT
        t0 = native_io.read_snapshot(ctx, fs, os.path.join(DROPIN, case, "io", "flow"), 0)
        assert t0 == 0.0
        got = ctx.fs_download(fs)[:ctx.n_cells]
        first, _ = native_io.read_arrays(os.path.join(DROPIN, case, "io", "flow"), 0, ctx.n_cells, dim)
        gm = config["gas_model"]
        assert np.array_equal(got[:, 2], first[0]) and np.array_equal(got[:, 1], first[1])
        assert np.array_equal(got[:, 0], first[1] / (gm["R"] * first[0])) and np.array_equal(got[:, 3], gm["Cv"] * first[0])
        # the reference's final snapshot, in through the ring, out through the writer
        host = ctx.pinned_array((len(names), ctx.n_cells))
        for q, k in enumerate(names):
            host[q] = exp["final_" + k]
        last = int(str(exp["final_dir"]))
        plot_dir = str(tmp_path / "flow")
        w = native_io.SnapshotWriter(ctx, plot_dir)
        for index in range(last + 1):                # one `flows` line per snapshot, as the run left them
            if index == last:
                ctx.snapshot_upload_begin(1, host)
                ctx.snapshot_upload_end(1, fs)
            w.write(fs, index, float(exp["final_time"]) if index == last else 0.0)
        w.close()
        d = os.path.join(plot_dir, f"{last:04d}")
        for k in names:
            with open(os.path.join(d, k), "rb") as f:
                assert f.read() == exp["final_" + k].tobytes(), f"{case}: file {k} differs from the reference's"
        with open(os.path.join(d, "meta_data.json")) as f:
            assert f.read() == str(exp["final_meta_text"])
        with open(os.path.join(plot_dir, "flows")) as f:
            assert f.read() == str(exp["flows_text"])
        assert sorted(os.listdir(d)) == sorted(list(names) + ["meta_data.json"])
    finally:
        ctx.close()


@pytest.mark.parametrize("case", FILE_CASES)
def test_snapshot_files_equal_the_references_emu(case, tmp_path):
    from emu_util import emu_lib
    _check_files(case, tmp_path, emu_lib())


@pytest.mark.gpu
@pytest.mark.parametrize("case", FILE_CASES)
def test_snapshot_files_equal_the_references_gpu(case, tmp_path):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _check_files(case, tmp_path)


def _check_overlap(lib_path=None):
    """Two snapshots in flight while steps are enqueued behind them: each lands with the state of ITS step (the
    ring's slots, events and the writer's double buffer)."""
    import sys
    from ibis_b200 import config as cfg
    from ibis_b200 import native_io
    import tempfile
    mesh, config, g, prim = cases.build_case("box3d_ausmdv")
    tab = cfg.BUTCHER["ssp-rk3"]
    ctx = Context(g, config, lib_path=lib_path)
    try:
        fs, cq = ctx.fs_create(), ctx.cq_create()
        ctx.fs_upload(fs, prim)
        ctx.prim_to_cons(cq, fs)
        want = []
        with tempfile.TemporaryDirectory() as plot_dir:
            w = native_io.SnapshotWriter(ctx, plot_dir)
            for step in range(5):
                ctx.rk_step_async(fs, cq, tab["a"], tab["b"], 0.5, 1.0)
                w.write(fs, step, 0.1 * step)
                # reference copy of the same state, taken synchronously (the stream is in order: this waits for the step)
                want.append(np.stack(ctx.flow_write_native(fs)))
            w.close()
            for step in range(5):
                got, t = native_io.read_arrays(plot_dir, step, ctx.n_cells, 3)
                assert t == 0.1 * step and np.array_equal(got, want[step]), step
        assert not np.array_equal(want[0], want[4])
    finally:
        ctx.close()


def test_snapshots_overlap_steps_emu():
    from emu_util import emu_lib
    _check_overlap(emu_lib())


@pytest.mark.gpu
def test_snapshots_overlap_steps_gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _check_overlap()


def test_snapshot_ring_rejects_misuse_without_poisoning_the_context():
    """bad slot numbers, an `end` without its `begin`, null arrays: EINVAL / EHANDLE from the call itself, and the
    context keeps working afterwards"""
    from emu_util import emu_lib
    from ibis_b200.lib import IbisB200Error
    mesh, config, g, prim = cases.build_case("quad2d_ausmdv")
    ctx = Context(g, config, lib_path=emu_lib())
    try:
        fs = ctx.fs_create()
        ctx.fs_upload(fs, prim)
        host = ctx.pinned_array((4, ctx.n_cells))
        with pytest.raises(IbisB200Error, match="without snapshot_upload_begin"):
            ctx.snapshot_upload_end(0, fs)
        with pytest.raises(IbisB200Error, match="without snapshot_download_begin"):
            ctx.snapshot_download_end(1)
        with pytest.raises(IbisB200Error, match="slot must be 0 or 1"):
            ctx.snapshot_download_begin(2, fs, host)
        with pytest.raises(IbisB200Error, match="not a FlowStates handle"):
            ctx.snapshot_download_begin(0, 9999, host)
        ctx.snapshot_download_begin(0, fs, host)
        with pytest.raises(IbisB200Error, match="without snapshot_download_begin"):
            ctx.snapshot_download_end(1)                   # slot 1 was never used
        ctx.snapshot_download_end(0)
        assert np.array_equal(host[0], prim[:ctx.n_cells, 2]) and np.array_equal(host[3], prim[:ctx.n_cells, 5])
        ctx.snapshot_upload_begin(1, host)
        ctx.snapshot_upload_end(1, fs)
        assert np.array_equal(ctx.fs_download(fs)[:ctx.n_cells, 2], prim[:ctx.n_cells, 2])
    finally:
        ctx.close()
