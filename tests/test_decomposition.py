"""Domain decomposition on the CPU: z-slabs of the box with their two-ring
halo, the send / receive lists and the owned / halo / ghost cell ranges of the kernels.

* one process, three slabs, halo exchange done by hand through the ABI: residual, dt and the
  state after Runge-Kutta stages of every owned cell are BIT-IDENTICAL to the undecomposed box;
* two processes (torch.distributed, gloo, world_size 2): the same with the halo exchange and the
  MIN / SUM reductions going through the process group, the way bench.py drives NCCL on the GPUs.

The kernels run in the host-emulation build (tests/emu); the NCCL exchange itself is covered on
the GPU box by tools/mgpu_check.py."""
import os
import sys

import numpy as np
import pytest

import cases
from emu_util import emu_host_lib, emu_lib
from ibis_b200 import config as cfg
from ibis_b200.boxgrid import BoxGrid
from ibis_b200.lib import Context

SHAPE = (6, 5, 12)


def box_config():
    return cfg.make_config(3, cases.box3_bcs(), flux_calculator="ausmdv", reconstruction_order=2,
                           solver=cfg.runge_kutta(method="ssp-rk3", cfl=0.5))


def centroid_key(c):
    return tuple(np.round(c * 1e6).astype(np.int64))


def global_field(whole):
    prim = cases.perturbed_state(whole, 3)
    index = {centroid_key(c): i for i, c in enumerate(whole.cell_centroid)}
    return prim, index


def local_field(slab, prim, index):
    l2g = np.array([index[centroid_key(c)] for c in slab.cell_centroid])
    return np.ascontiguousarray(prim[l2g]), l2g


def exchange_by_hand(ctxs, fss, slabs):
    """What comm::halo_exchange does with ncclSend/ncclRecv, through fs_download / fs_upload."""
    states = [c.fs_download(f) for c, f in zip(ctxs, fss)]
    for r, s in enumerate(slabs):
        for p, d in enumerate(s.peer_dir):
            o = slabs[r + d]
            q = o.peer_dir.index(-d)
            mine = s.recv_cells[s.recv_offsets[p]:s.recv_offsets[p + 1]]
            theirs = o.send_cells[o.send_offsets[q]:o.send_offsets[q + 1]]
            states[r][mine] = states[r + d][theirs]
    for c, f, st in zip(ctxs, fss, states):
        c.fs_upload(f, st)


@pytest.mark.parametrize("sides", ["0", "1"])
def test_three_slabs_equal_whole_box(sides, monkeypatch):
    monkeypatch.setenv("IBIS_B200_SIDES", sides)
    lib, hlib = emu_lib(), emu_host_lib()
    config = box_config()
    whole = BoxGrid(*SHAPE, host_lib=hlib, core_lib=lib)
    prim, index = global_field(whole)
    wctx = Context(whole, config, lib_path=lib)
    wfs, wdq, wcq = wctx.fs_create(), wctx.cq_create(), wctx.cq_create()
    wctx.fs_upload(wfs, prim)
    wctx.compute_dudt(wfs, wdq, True)
    ref_dudt = wctx.cq_download(wdq)
    ref_dt = wctx.estimate_dt(wfs)
    ref_l2 = wctx.l2_norms(wdq)

    cuts = [0, 4, 8, 12]
    slabs = [BoxGrid(*SHAPE, slab=(cuts[r], cuts[r + 1]), host_lib=hlib, core_lib=lib) for r in range(3)]
    ctxs = [Context(s, config, lib_path=lib) for s in slabs]
    fss, dqs, l2gs = [], [], []
    for s, c in zip(slabs, ctxs):
        loc, l2g = local_field(s, prim, index)
        loc[s.n_owned:s.n_cells] = -1.0e300          # halo copies arrive only through the exchange
        f = c.fs_create()
        c.fs_upload(f, loc)
        fss.append(f)
        dqs.append(c.cq_create())
        l2gs.append(l2g)
    exchange_by_hand(ctxs, fss, slabs)
    dts, l2sq = [], np.zeros(6)
    for s, c, f, dq, l2g in zip(slabs, ctxs, fss, dqs, l2gs):
        c.compute_dudt(f, dq, True)
        d = c.cq_download(dq)
        no = s.n_owned
        assert np.array_equal(d[:no], ref_dudt[l2g[:no]]), "owned residuals differ from the whole box"
        assert not d[no:].any(), "halo / ghost rows of dudt must stay zero"
        dts.append(c.estimate_dt(f))
        l2sq += np.asarray(c.l2_norms(dq)) ** 2
    assert min(dts) == ref_dt
    np.testing.assert_allclose(np.sqrt(l2sq), ref_l2, rtol=1e-13)

    # two explicit Euler stages with an exchange in between: still the whole-box bits
    dt = 0.25 * ref_dt
    wctx.prim_to_cons(wcq, wfs)
    cqs = [c.cq_create() for c in ctxs]
    for c, f, q in zip(ctxs, fss, cqs):
        c.prim_to_cons(q, f)
    for _ in range(2):
        wctx.compute_dudt(wfs, wdq, True)
        wctx.apply_time_derivative(wcq, wcq, wdq, dt)
        wctx.cons_to_prim(wcq, wfs)
        for c, f, dq, q in zip(ctxs, fss, dqs, cqs):
            c.compute_dudt(f, dq, True)
            c.apply_time_derivative(q, q, dq, dt)
            c.cons_to_prim(q, f)
        exchange_by_hand(ctxs, fss, slabs)
    ref_state = wctx.fs_download(wfs)
    for s, c, f, l2g in zip(slabs, ctxs, fss, l2gs):
        st = c.fs_download(f)
        assert np.array_equal(st[:s.n_owned], ref_state[l2g[:s.n_owned]])
    for c in ctxs + [wctx]:
        c.close()


def _worker(rank, world, port, out_queue):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib, hlib = emu_lib(), emu_host_lib()
    config = box_config()
    nz = SHAPE[2]
    cuts = [nz * r // world for r in range(world + 1)]
    whole = BoxGrid(*SHAPE, host_lib=hlib, core_lib=lib)
    prim, index = global_field(whole)
    slab = BoxGrid(*SHAPE, slab=(cuts[rank], cuts[rank + 1]), host_lib=hlib, core_lib=lib)
    loc, l2g = local_field(slab, prim, index)
    loc[slab.n_owned:slab.n_cells] = -1.0e300
    ctx = Context(slab, config, lib_path=lib)
    fs, dq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, loc)

    def exchange():
        st = ctx.fs_download(fs)
        reqs, bufs = [], []
        for p, d in enumerate(slab.peer_dir):
            snd = torch.from_numpy(np.ascontiguousarray(st[slab.send_cells[slab.send_offsets[p]:slab.send_offsets[p + 1]]]))
            rcv = torch.empty((int(slab.recv_offsets[p + 1] - slab.recv_offsets[p]), 7), dtype=torch.float64)
            reqs += [dist.isend(snd, rank + d), dist.irecv(rcv, rank + d)]
            bufs.append((p, rcv, snd))
        for r in reqs:
            r.wait()
        for p, rcv, _ in bufs:
            st[slab.recv_cells[slab.recv_offsets[p]:slab.recv_offsets[p + 1]]] = rcv.numpy()
        ctx.fs_upload(fs, st)

    exchange()
    ctx.compute_dudt(fs, dq, True)
    d = ctx.cq_download(dq)[:slab.n_owned]
    t = torch.tensor([ctx.estimate_dt(fs)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)                    # the dt reduction of estimate_dt
    sq = torch.tensor(np.asarray(ctx.l2_norms(dq)) ** 2)
    dist.all_reduce(sq, op=dist.ReduceOp.SUM)                   # the L2-norm reduction
    out_queue.put((rank, l2g[:slab.n_owned], d, float(t.item()), np.sqrt(sq.numpy())))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    lib, hlib = emu_lib(), emu_host_lib()
    whole = BoxGrid(*SHAPE, host_lib=hlib, core_lib=lib)
    prim, _ = global_field(whole)
    wctx = Context(whole, box_config(), lib_path=lib)
    wfs, wdq = wctx.fs_create(), wctx.cq_create()
    wctx.fs_upload(wfs, prim)
    wctx.compute_dudt(wfs, wdq, True)
    ref = wctx.cq_download(wdq)
    ref_dt, ref_l2 = wctx.estimate_dt(wfs), wctx.l2_norms(wdq)
    wctx.close()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = 0
    for rank, l2g, d, dt, l2 in results:
        assert np.array_equal(d, ref[l2g]), f"rank {rank}: owned residuals differ from the whole box"
        assert dt == ref_dt
        np.testing.assert_allclose(l2, ref_l2, rtol=1e-13)
        seen += len(l2g)
    assert seen == whole.n_cells


if __name__ == "__main__":
    sys.exit(pytest.main([__file__, "-x", "-q"]))


# ---- general unstructured grids: recursive coordinate bisection (ibis_b200/partition.py) ----
def _exchange_subs(ctxs, fss, subs):
    states = [c.fs_download(f) for c, f in zip(ctxs, fss)]
    for r, s in enumerate(subs):
        for p, q in enumerate(s.peer_ranks):
            o = subs[q]
            qp = o.peer_ranks.index(r)
            mine = s.recv_cells[s.recv_offsets[p]:s.recv_offsets[p + 1]]
            theirs = o.send_cells[o.send_offsets[qp]:o.send_offsets[qp + 1]]
            assert len(mine) == len(theirs)
            states[r][mine] = states[q][theirs]
    for c, f, st in zip(ctxs, fss, states):
        c.fs_upload(f, st)


@pytest.mark.parametrize("name,n_parts", [("mixed3d_walls", 3), ("naca_omesh", 4), ("tri2d", 2), ("box3d_ausmdv", 5),
                                          ("quad2d_viscous", 3), ("box3d_profile_cflux", 3),
                                          ("quad2d_bl_profile", 2)])
def test_rcb_partition_equals_whole_grid(name, n_parts):
    from ibis_b200 import partition
    lib = emu_lib()
    mesh, config, g, prim = cases.build_case(name)
    owner, subs = partition.decompose(g, n_parts)
    counts = np.bincount(owner, minlength=n_parts)
    assert counts.max() - counts.min() <= 1 and counts.sum() == g.n_cells

    wctx = Context(g, config, lib_path=lib)
    wfs, wdq = wctx.fs_create(), wctx.cq_create()
    wctx.fs_upload(wfs, prim)
    wctx.compute_dudt(wfs, wdq, True)
    ref = wctx.cq_download(wdq)
    ref_dt = wctx.estimate_dt(wfs)
    ref_o1 = None
    wctx.compute_dudt(wfs, wdq, False)
    ref_o1 = wctx.cq_download(wdq)

    ctxs = [Context(s, config, lib_path=lib) for s in subs]
    fss, dqs = [], []
    for s, c in zip(subs, ctxs):
        loc = np.ascontiguousarray(prim[s.local_to_global])
        loc[s.n_owned:s.n_cells] = -1.0e300                 # halo values arrive only through the exchange
        f = c.fs_create()
        c.fs_upload(f, loc)
        fss.append(f)
        dqs.append(c.cq_create())
    _exchange_subs(ctxs, fss, subs)
    dts, seen = [], 0
    for s, c, f, dq in zip(subs, ctxs, fss, dqs):
        own = s.local_to_global[:s.n_owned]
        c.compute_dudt(f, dq, True)
        d = c.cq_download(dq)
        assert np.array_equal(d[:s.n_owned], ref[own]), f"{name}: owned residuals differ from the whole grid"
        assert not d[s.n_owned:].any()
        dts.append(c.estimate_dt(f))
        c.compute_dudt(f, dq, False)
        assert np.array_equal(c.cq_download(dq)[:s.n_owned], ref_o1[own]), f"{name}: first-order residuals differ"
        seen += s.n_owned
    assert seen == g.n_cells and min(dts) == ref_dt
    for c in ctxs + [wctx]:
        c.close()
