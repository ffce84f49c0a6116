"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref/libibis_ref.so, i.e.
/root/reference compiled by oracle/ref_build/Makefile).  Run in the build container:

    make -C oracle/ref_build -j8 && python tests/golden/make_golden.py

Each fixture holds, for one case of tests/cases.py: the input primitive state, and the
reference's outputs of FiniteVolume<real>::compute_dudt (dudt, ghost-filled state,
gradients, limiters), estimate_dt, count_bad_cells, L2 norms, primitive<->conserved
conversion, and of FiniteVolume<dual>::compute_dudt for a seeded dual part (real and dual
parts of the residual), with allow_reconstruction on and off."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import cases  # noqa: E402
import refharness as rh  # noqa: E402


def main():
    pad = os.path.join(ROOT, "oracle", "_ref", "libpadmalloc.so")
    if os.path.exists(pad) and "libpadmalloc" not in os.environ.get("LD_PRELOAD", ""):
        # oracle/ref_build/padmalloc.c: the reference's CubicSpline constructor overruns two arrays
        os.environ["LD_PRELOAD"] = pad
        os.execv(sys.executable, [sys.executable] + sys.argv)
    tmp = "/tmp/ibis_golden"
    os.makedirs(tmp, exist_ok=True)
    only = sys.argv[1:]                      # optional: regenerate just these cases
    for name in cases.CASES:
        if only and name not in only:
            continue
        mesh, config, g, prim = cases.build_case(name)
        path = os.path.join(tmp, name + ".su2")
        mesh.write(path)
        out = {"prim_in": prim}
        rc = rh.RefCase(path, config)
        ga = rc.grid_arrays()
        for k in ("face_left", "face_right", "face_area", "face_norm", "cell_volume", "cell_centroid",
                  "cell_face_ids", "cell_outsigns", "cell_neighbours"):
            out["grid_" + k] = ga[k]
        for allow, sfx in ((True, ""), (False, "_o1")):
            d, p = rc.compute_dudt(prim, allow_reconstruction=allow)
            out["dudt" + sfx], out["prim_out" + sfx] = d, p
            out["l2" + sfx] = rc.l2_norms()
            if allow:
                it = rc.intermediates()
                for k in ("grad_rho", "grad_u", "grad_p", "grad_T", "grad_vx", "grad_vy", "grad_vz",
                          "lim_rho", "lim_u", "lim_p", "lim_T", "lim_vx", "lim_vy", "lim_vz"):
                    if not np.all(np.isnan(it[k])):
                        out[k] = it[k]
                out["dt"] = np.array(rc.estimate_dt(p))
                out["bad"] = np.array(rc.count_bad_cells(p))
        r = rc.wls_r()
        if r is not None:
            out["wls_r"] = r
        cons = rc.prim_to_cons(prim)
        out["cons"] = cons
        out["prim_from_cons"] = rc.cons_to_prim(cons)
        rc.close()
        # dual residual = one exact Jacobian-vector product direction
        rcd = rh.RefCase(path, config, dual=True)
        pdu = cases.dual_seed(prim)
        for allow, sfx in ((True, ""), (False, "_o1")):
            dr, pr, dd, pd = rcd.compute_dudt(prim, pdu, allow_reconstruction=allow)
            out["dual_dudt_re" + sfx], out["dual_dudt_du" + sfx] = dr, dd
            out["dual_prim_out_du" + sfx] = pd
        rcd.close()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, g.n_cells, "cells", {k: v.shape for k, v in list(out.items())[:3]})


def main_moving(only):
    """moving_<case>.npz: the moving-grid overload of compute_dudt with the rigid_body_translation driver"""
    tmp = "/tmp/ibis_golden"
    os.makedirs(tmp, exist_ok=True)
    for name in cases.MOVING_CASES:
        if only and "moving_" + name not in only:
            continue
        mesh, config, g, prim, vvel = cases.build_moving_case(name)
        path = os.path.join(tmp, "moving_" + name + ".su2")
        mesh.write(path)
        out = {"prim_in": prim, "vertex_vel": vvel}
        rc = rh.RefCase(path, config)
        for allow, sfx in ((True, ""), (False, "_o1")):
            d, p, fv = rc.compute_dudt_moving(prim, allow_reconstruction=allow)
            out["dudt" + sfx], out["prim_out" + sfx], out["face_vel" + sfx] = d, p, fv
        rc.close()
        # the same state on the grid at rest, for contrast (the static overload must not be called on a grid
        # whose motion is enabled: it hands an EMPTY vertex-velocity array to the motion driver)
        d0 = dict(np.load(os.path.join(HERE, name + ".npz")))["dudt"]
        rcd = rh.RefCase(path, config, dual=True)
        pdu = cases.dual_seed(prim)
        for allow, sfx in ((True, ""), (False, "_o1")):
            dr, pr, dd, pd, fv = rcd.compute_dudt_moving(prim, pdu, allow_reconstruction=allow)
            out["dual_dudt_re" + sfx], out["dual_dudt_du" + sfx] = dr, dd
        rcd.close()
        np.savez_compressed(os.path.join(HERE, "moving_" + name + ".npz"), **out)
        print("moving", name, g.n_cells, "cells; max |dudt - dudt_static| / max |dudt| =",
              float(np.abs(out["dudt"] - d0).max() / np.abs(d0).max()))


if __name__ == "__main__":
    if any(a.startswith("moving") for a in sys.argv[1:]):
        pad = os.path.join(ROOT, "oracle", "_ref", "libpadmalloc.so")
        if os.path.exists(pad) and "libpadmalloc" not in os.environ.get("LD_PRELOAD", ""):
            os.environ["LD_PRELOAD"] = pad
            os.execv(sys.executable, [sys.executable] + sys.argv)
        main_moving([a for a in sys.argv[1:] if a != "moving"])
    else:
        main()
