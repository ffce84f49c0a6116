"""Parity checks shared by the CPU (host-emulation) and GPU test files: run one case of
tests/cases.py through the C ABI of a given build of the library and compare with the
golden vectors produced by the real reference (tests/golden/make_golden.py).

Tolerances.  The library is built without FMA contraction and the kernels follow the
reference's expression trees, so +,-,*,/ and sqrt are the same correctly rounded IEEE
operations on both sides: inviscid results are required to be BIT-IDENTICAL.  The only
non-IEEE function on the path is pow() in Sutherland's law (viscous cases): CUDA's pow is
accurate to 2 ulp where glibc's is < 1 ulp, so viscous residuals are required to match to
1e-12 relative, the north-star tolerance, with the per-cell flux scale as the floor
(|a-b| <= 1e-12 * max(|b|, S_c)).
"""
import os

import numpy as np

import cases
from ibis_b200.lib import Context

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def flux_scale(g, prim, dim):
    """Per-cell scale S_c = sum_f |F|_f A_f / V_c with |F| ~ rho |v| (|v|+a) (+p), per equation."""
    rho, p, T, vx, vy, vz = prim[:, 0], prim[:, 1], prim[:, 2], prim[:, 4], prim[:, 5], prim[:, 6]
    a = np.sqrt(1.4 * 287.0 * np.abs(T))
    v = np.sqrt(vx * vx + vy * vy + vz * vz) + a
    n = g.n_cells
    geom = np.zeros(n)
    np.add.at(geom, np.repeat(np.arange(n), np.diff(g.cell_face_offsets)), g.face_area[g.cell_face_ids])
    geom = geom / g.cell_volume
    mass = (rho * v)[:n] * geom
    mom = (rho * v * v + p)[:n] * geom
    en = (rho * v * (p / np.abs(rho) * 3.5 + 0.5 * v * v))[:n] * geom
    cols = [mass, mom, mom] + ([mom] if dim == 3 else []) + [en]
    return np.stack(cols, axis=1)


def assert_close(a, b, scale, what, exact):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    if exact:
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f"{what}: {int((~same).sum())} of {a.size} values not bit-identical, " \
                           f"max |diff| {np.nanmax(np.abs(a - b)):.3e}"
    else:
        err = np.abs(a - b)
        bound = RTOL * np.maximum(np.abs(b), scale)
        bad = ~((err <= bound) | (np.isnan(a) & np.isnan(b)))
        assert not bad.any(), f"{what}: {int(bad.sum())} values outside 1e-12 relative, " \
                              f"worst {np.nanmax(err / np.maximum(np.abs(b), scale)):.3e}"


def run_case(name, lib_path=None, check_dual=True):
    mesh, config, g, prim = cases.build_case(name)
    gold = load_golden(name)
    np.testing.assert_array_equal(prim, gold["prim_in"])
    exact = not config["viscous_flux"]["enabled"]
    S = flux_scale(g, prim, mesh.dim)
    ctx = Context(g, config, lib_path=lib_path)
    try:
        fs, dq = ctx.fs_create(), ctx.cq_create()
        for allow, sfx in ((True, ""), (False, "_o1")):
            ctx.fs_upload(fs, prim)
            ctx.compute_dudt(fs, dq, allow)
            d = ctx.cq_download(dq)
            assert not d[g.n_cells:].any(), "ghost rows of dudt must stay zero"
            assert_close(d[:g.n_cells], gold["dudt" + sfx], S, f"{name}: dudt{sfx}", exact)
            assert_close(ctx.fs_download(fs), gold["prim_out" + sfx], 1.0, f"{name}: ghost fill{sfx}", True)
            l2 = ctx.l2_norms(dq)
            np.testing.assert_allclose(l2, gold["l2" + sfx], rtol=1e-12, err_msg=f"{name}: L2 norms")
            if allow:
                dt = ctx.estimate_dt(fs)
                if exact:
                    assert dt == float(gold["dt"]), f"{name}: dt {dt} vs {float(gold['dt'])}"
                else:
                    assert abs(dt - float(gold["dt"])) <= 1e-12 * float(gold["dt"])
                assert ctx.count_bad_cells(fs) == int(gold["bad"])
                if config["convective_flux"]["reconstruction_order"] == 2 and exact:
                    th = {"rho_u": "u", "rho_p": "p", "rho_T": "T"}[config["convective_flux"]["thermo_interpolator"]]
                    names = ["rho", th, "vx", "vy"] + (["vz"] if mesh.dim == 3 else [])
                    for var, nm in enumerate(names):
                        gr, lim = ctx.debug_gradients(var)
                        assert_close(gr, gold["grad_" + nm], 1.0, f"{name}: grad {nm}", True)
                        if config["convective_flux"]["limiter"]["type"] == "barth_jespersen":
                            assert_close(lim, gold["lim_" + nm], 1.0, f"{name}: limiter {nm}", True)
        # primitive <-> conserved
        cq = ctx.cq_create()
        ctx.fs_upload(fs, prim)
        ctx.prim_to_cons(cq, fs)
        assert_close(ctx.cq_download(cq), gold["cons"], 1.0, f"{name}: prim_to_cons", True)
        ctx.cons_to_prim(cq, fs)
        assert_close(ctx.fs_download(fs), gold["prim_from_cons"], 1.0, f"{name}: cons_to_prim", True)
        if check_dual:
            fsd, dqd = ctx.fs_create(True), ctx.cq_create(True)
            pdu = cases.dual_seed(prim)
            for allow, sfx in ((True, ""), (False, "_o1")):
                ctx.fs_upload(fsd, prim, pdu)
                ctx.compute_dudt(fsd, dqd, allow)
                dr, dd = ctx.cq_download(dqd, dual=True)
                assert_close(dr[:g.n_cells], gold["dual_dudt_re" + sfx], S, f"{name}: dual re{sfx}", exact)
                # the dual part is a directional derivative: its scale is S_c * |du/u| ~ 1e-3 S_c
                assert_close(dd[:g.n_cells], gold["dual_dudt_du" + sfx], 1e-3 * S, f"{name}: dual du{sfx}",
                             exact and config["convective_flux"]["flux_calculator"]["type"] != "ldfss2")
                # the real part of a dual evaluation IS the real evaluation
                assert_close(dr[:g.n_cells], gold["dudt" + sfx], S, f"{name}: dual re vs real{sfx}", exact)
    finally:
        ctx.close()
