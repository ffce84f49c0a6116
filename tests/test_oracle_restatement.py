"""oracle/restatement.py (plain numpy, written from the reference's sources) against the golden vectors of the
COMPILED reference: bit for bit -- residual, ghost fill, every gradient and limiter, second and first order --
on every inviscid case it covers.  This pins the restatement; tests/test_gpu_parity.py then holds the CUDA path
to the same vectors, and test_product_equals_restatement_on_seeded_input below holds it to the restatement on
inputs no fixture stores."""
import os
import sys

import numpy as np
import pytest

import cases
import parity

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import restatement  # noqa: E402

COVERED = sorted(n for n, (_, bcs_f, _, kw) in cases.CASES.items()
                 if not kw.get("viscous") and kw.get("flux_calculator", "hanel") in restatement.FLUX
                 and all(a["type"] in ("flow_state_copy", "internal_copy", "internal_copy_reflect_normal",
                                       "subsonic_inflow", "subsonic_outflow")
                         for b in bcs_f().values() for a in b["pre_reconstruction"])
                 and not any(b["post_convective_flux"] for b in bcs_f().values()))


def test_coverage_is_what_the_header_says():
    assert len(COVERED) >= 19 and {"box3d_ausmdv", "box3d_hanel", "box3d_rusanov", "box3d_ldfss2", "quad2d_ldfss2",
                                    "quad2d_subsonic_bcs", "quad2d_ausmdv", "mixed3d_walls",
                                    "tet3d", "tri2d", "wedge2d", "box3d_rho_p", "box3d_rho_T", "box3d_unlimited",
                                    "box3d_order1"} <= set(COVERED)


@pytest.mark.parametrize("name", COVERED)
def test_restatement_equals_compiled_reference(name):
    mesh, config, g, prim = cases.build_case(name)
    gold = parity.load_golden(name)
    for allow, sfx in ((True, ""), (False, "_o1")):
        dudt, fs, inter = restatement.compute_dudt(g, config, prim, allow)
        parity.assert_close(dudt, gold["dudt" + sfx], 1.0, f"{name}: dudt{sfx}", True)
        parity.assert_close(fs, gold["prim_out" + sfx], 1.0, f"{name}: ghost fill{sfx}", True)
        if allow:
            for k, v in inter.items():
                if k in gold:
                    parity.assert_close(v, gold[k], 1.0, f"{name}: {k}", True)
            if config["convective_flux"]["reconstruction_order"] == 2:
                assert any(k.startswith("grad_") for k in inter)


@pytest.mark.parametrize("build", [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("seed", [7, 20251018])
def test_product_equals_restatement_on_seeded_input(build, seed):
    """a larger jittered box and a state no fixture holds: the product path through the C ABI (host harness here,
    CUDA under -m gpu) == restatement, bit for bit"""
    lib_path = None
    if build == "emu":
        from emu_util import emu_lib
        lib_path = emu_lib()
    else:
        import torch
        assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ibis_b200 import config as cfg
    from ibis_b200 import grid as gridmod
    from ibis_b200 import su2
    from ibis_b200.lib import Context
    mesh = su2.box_hex(14, 12, 10, jitter=0.25, seed=seed)
    bcs = cases.box3_wall_bcs()
    config = cfg.make_config(3, bcs, flux_calculator="ausmdv")
    g = gridmod.build_grid(mesh, list(bcs.keys()))
    prim = cases.perturbed_state(g, 3, 0.7, seed=seed)
    ctx = Context(g, config, lib_path=lib_path)
    fs, dq = ctx.fs_create(), ctx.cq_create()
    for allow in (True, False):
        ctx.fs_upload(fs, prim)
        ctx.compute_dudt(fs, dq, allow)
        want, ghosts, _ = restatement.compute_dudt(g, config, prim, allow)
        assert np.array_equal(ctx.cq_download(dq)[:g.n_cells], want)
        assert np.array_equal(ctx.fs_download(fs), ghosts)
    ctx.close()
