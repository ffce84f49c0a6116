#!/usr/bin/env python
"""tools/mgpu_check.py: decomposed run == undecomposed run, on real GPUs over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/mgpu_check.py [--n 48] [--steps 3]

Every rank owns a z-slab of the n^3 box (two-ring halo, grouped ncclSend/ncclRecv inside
compute_dudt / rk_step, ncclAllReduce MIN for dt); rank 0 also runs the whole box on its GPU.
After `steps` ssp-rk3 steps the owned cells of all ranks must equal the whole-box state BIT FOR
BIT (per-cell arithmetic does not depend on the partition; the dt MIN reduction is exact), and the
L2 residual norms (SUM reduction) must agree to rounding.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", dest="n", type=int, default=48)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--krylov", type=int, default=6, help="GMRES iterations of the decomposed JFNK check")
    ap.add_argument("--case", default=None,
                    help="a case of tests/cases.py (unstructured SU2 mesh) decomposed by recursive coordinate "
                         "bisection (ibis_b200/partition.py) instead of the box in slabs")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    import cases
    from ibis_b200 import config as cfg
    from ibis_b200.boxgrid import BoxGrid
    from ibis_b200.lib import Context

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.n
    tab = cfg.BUTCHER["ssp-rk3"]
    if args.case:
        from ibis_b200 import partition
        mesh, config, whole, gprim = cases.build_case(args.case)
        _, subs = partition.decompose(whole, world)
        slab = subs[rank]
        gid = slab.local_to_global[:slab.n_cells]
        loc = np.ascontiguousarray(gprim[slab.local_to_global])
        peers = slab.peer_ranks
        n_global = whole.n_cells
    else:
        config = bench.box_config()
        whole = BoxGrid(n)
        gprim = cases.perturbed_state(whole, 3)
        slab = bench.box_grid_fast(n, rank, world)
        ijk = np.clip(np.floor(slab.cell_centroid[:slab.n_cells] * n).astype(np.int64), 0, n - 1)
        gid = ijk[:, 0] + n * (ijk[:, 1] + n * ijk[:, 2])
        loc = np.ones((slab.n_total, 7))
        loc[:slab.n_cells] = gprim[gid]
        peers = [rank + d for d in slab.peer_dir]
        n_global = n ** 3
    loc[slab.n_owned:slab.n_cells] = -1.0e300          # halo values must come through the exchange
    ctx = Context(slab, config, device=local)
    mode = ctx.connect(dist, rank, world, peers, slab.send_offsets, slab.send_cells,
                       slab.recv_offsets, slab.recv_cells, p2p=os.environ.get("IBIS_B200_HALO", "p2p") != "nccl")
    fs, cq = ctx.fs_create(), ctx.cq_create()
    ctx.fs_upload(fs, loc)
    ctx.prim_to_cons(cq, fs)
    dt, dts = sys.float_info.max, []
    for _ in range(args.steps):
        dt, _ = ctx.rk_step(fs, cq, tab["a"], tab["b"], 0.5, dt, 1.0)
        dts.append(dt)
    norms = ctx.rk_residual_norms()
    state = ctx.fs_download(fs)[:slab.n_owned]
    gathered = [None] * world
    dist.gather_object((gid[:slab.n_owned], state, dts, list(norms)), gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        wctx = Context(whole, config, device=local)
        wfs, wcq = wctx.fs_create(), wctx.cq_create()
        wctx.fs_upload(wfs, gprim)
        wctx.prim_to_cons(wcq, wfs)
        dt, wdts = sys.float_info.max, []
        for _ in range(args.steps):
            dt, _ = wctx.rk_step(wfs, wcq, tab["a"], tab["b"], 0.5, dt, 1.0)
            wdts.append(dt)
        wstate = wctx.fs_download(wfs)
        wnorms = wctx.rk_residual_norms()
        seen, mism = 0, 0
        for g_ids, st, d, nm in gathered:
            mism += int((st != wstate[g_ids]).any(axis=1).sum())
            seen += len(g_ids)
            ok &= d == wdts
            ok &= bool(np.allclose(nm, wnorms, rtol=1e-12))
        ok &= mism == 0 and seen == n_global
        print(json.dumps({"check": "decomposed == whole grid", "grid": args.case or f"box {n}^3", "ranks": world, "halo": mode, "steps": args.steps,
                          "cells_compared": seen, "cells_differing": mism, "dt_identical": bool(gathered[0][2] == wdts),
                          "l2_norms_decomposed": gathered[0][3], "l2_norms_whole": list(wnorms), "ok": bool(ok)}))
        wctx.close()
    # ---- the JFNK half on the decomposed grid: m GMRES iterations (dual-number matvec + modified Gram-Schmidt) from
    # the same global start vector.  Inner products run over owned cells and are summed over the ranks
    # (abi.cu::dot_collective): every rank must see the Hessenberg columns of the undecomposed run (to rounding: the
    # sums are taken in another order) and hold its part of the same Krylov vectors.
    if not args.case:
        m = args.krylov

        def krylov(c, grid, prim_local, ids):
            nloc = int(grid.n_cells)
            f, q, tmp, dq = c.fs_create(dual=True), c.cq_create(dual=True), c.cq_create(dual=True), c.cq_create(dual=True)
            c.fs_upload(f, prim_local, 0.0 * prim_local)
            c.prim_to_cons(q, f)
            v, w = c.vec_create(nloc * 5), c.vec_create(nloc * 5)
            basis = c.basis_create(nloc * 5, m + 1)
            x0 = np.sin(0.37 * ids[:, None] + np.arange(5)[None, :]).reshape(-1)
            c.vec_upload(v, x0)
            c.vec_scale(v, v, 1.0 / c.vec_norm2(v))
            c.basis_set_col(basis, 0, v)
            H = []
            for j in range(m):
                c.jfnk_set_dual(q, v, tmp)
                c.cons_to_prim(tmp, f)
                c.compute_dudt(f, dq, True)
                c.jfnk_set_vector(dq, v, 1e-4, w)
                H.append(c.arnoldi_mgs(basis, j, w, v).tolist())
            return H, c.vec_download(v, nloc * 5).reshape(nloc, 5)

        loc2 = np.ones((slab.n_total, 7))
        loc2[:slab.n_cells] = gprim[gid]
        Hd, vd = krylov(ctx, slab, loc2, gid)
        gathered = [None] * world
        dist.gather_object((gid[:slab.n_owned], Hd, vd[:slab.n_owned]), gathered if rank == 0 else None, dst=0)
        if rank == 0:
            wctx = Context(whole, config, device=local)
            Hw, vw = krylov(wctx, whole, gprim, np.arange(n_global))
            wctx.close()
            okj, worst_h, worst_v = True, 0.0, 0.0
            for g_ids, Hr, vr in gathered:
                for a, b in zip(Hr, Hw):
                    # relative to the column's largest entry (an entry that cancels to nearly zero has no relative accuracy)
                    worst_h = max(worst_h, float(np.max(np.abs(np.array(a) - np.array(b))) / np.max(np.abs(b))))
                worst_v = max(worst_v, float(np.max(np.abs(vr - vw[g_ids]))))
            okj = worst_h < 1e-9 and worst_v < 1e-9
            ok &= okj
            print(json.dumps({"check": "decomposed GMRES == whole-grid GMRES", "grid": f"box {n}^3", "ranks": world, "krylov_dim": m,
                              "max_diff_hessenberg_rel_to_column": worst_h, "max_abs_diff_last_krylov_vector": worst_v, "ok": bool(okj)}))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
