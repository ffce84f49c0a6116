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
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
