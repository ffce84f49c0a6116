"""Generate tests/golden/dropin/<case>/: the reference's own tests/minimal cases, prepared by the reference's own
`ibis prep` (job.py -> config/config.json, grid, snapshot 0; oracle/refharness.py::run_prep) and solved by the REAL
reference (oracle/_ref/libibis_ref.so, Solver::solve, one OpenMP thread) -- the final snapshot is what
tests/test_gpu_dropin.py holds the drop-in library (the same reference with its FiniteVolume bound to
libibis_b200.so) to.

    make -C oracle/ref_build -j8 && python tests/golden/make_dropin_cases.py
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

import refharness as rh  # noqa: E402

MINIMAL = "/root/reference/tests/minimal"
# case -> (directory with job.py, directory with the grid file the job names)
CASES = {
    "uniform_stream": ("uniform_stream", "uniform_stream"),
    "uniform_stream_3D": ("uniform_stream_3D", "uniform_stream_3D"),
    "steady_state_uniform_stream": ("steady_state_uniform_stream", "uniform_stream"),
}
KEEP = ["config/config.json", "io/grid/0000/block_0000.su2", "io/flow/flows", "io/flow/0000"]

# The reference's EXAMPLES, job.py unchanged, through the reference's own prep.  Their grids are gmsh scripts (gmsh is
# not in this image): the grid file the job names is written by the generators of ibis_b200/su2.py with the same
# boundary tags, at about the example's resolution.  After prep, config.json is edited ONLY as listed under "edit"
# (step caps that bound the run; a CFL number at which the REFERENCE itself survives the impulsive start on these
# structured grids -- at the examples' cfl = 2.0 - 2.5 it reports bad cells within twenty steps; native_binary
# snapshots) -- everything on the residual path is what prep produced.
EXAMPLES = "/root/reference/examples"
STEADY_BLOCK = """config.solver = SteadyState(
    cfl=2.0,
    max_steps=12,
    print_frequency=1,
    plot_frequency=100,
    diagnostics_frequency=1,
    tolerance=1e-10,
    linear_solver=Gmres(tol=1e-2, max_iters=30)
)
"""


def _naca_steady_job(text):
    """examples/naca0012/job.py with SteadyState(...) swapped in for RungeKutta(...) (BASELINE.json configs[2])"""
    a = text.index("config.solver = RungeKutta(")
    b = text.index(")\n", a) + 2
    return text[:a] + STEADY_BLOCK + text[b:]


EXAMPLE_CASES = {
    "example_wedge_transient": dict(job="wedge/2D/transient/job.py", grid=lambda su2: su2.wedge_2d(90, 60),
                                    edit={"solver": {"max_step": 150, "cfl": {"value": 0.8}}}),
    "example_forward_facing_step": dict(job="forward_facing_step/job.py", grid=lambda su2: su2.forward_facing_step_2d(0.025),
                                        edit={"solver": {"max_step": 150, "cfl": {"value": 0.8}}}),
    "example_naca0012_viscous": dict(job="naca0012/job.py", grid=lambda su2: su2.naca0012_omesh(96, 32),
                                     edit={"solver": {"max_step": 60}}),
    # examples/wedge/2D/steady_state has no fixture at this resolution either: from the developed shock of the transient
    # case the reference's GMRES does not converge within its 100 (or 250) iterations at the example's cfl = 10, and
    # on reaching max_iters its Gmres stores a Hessenberg column past its arrays -- the stock build survives that by
    # luck of the heap layout, any other link order does not.  The steady solver is held to the reference on
    # tests/minimal/steady_state_uniform_stream here and on the five steady runs of tests/solve_cases.py.
    # examples/naca0012 with SteadyState(...) swapped in (BASELINE.json configs[2], _naca_steady_job above) is prepared by
    # the reference's prep but has no expected snapshot: the reference itself dies of heap corruption on it
    # ("munmap_chunk(): invalid pointer", viscous or not) before the first step completes.  The product runs it
    # (tests/test_prep_configs.py::test_naca0012_steady_state_runs: residual norms fall, no bad cells).
    "example_naca0012_steady": dict(job="naca0012/job.py", job_filter=_naca_steady_job,
                                    grid=lambda su2: su2.naca0012_omesh(96, 32), edit={}, prep_only=True),
}


def merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict):
            merge(dst[k], v)
        else:
            assert k in dst, k
            dst[k] = v


def prepare_example(name, spec, finals):
    from ibis_b200 import su2
    work = os.path.join("/tmp/ibis_dropin_golden", name)
    if os.path.exists(work):
        shutil.rmtree(work)
    os.makedirs(work)
    with open(os.path.join(EXAMPLES, spec["job"])) as f:
        text = f.read()
    if "job_filter" in spec:
        text = spec["job_filter"](text)
    with open(os.path.join(work, "job.py"), "w") as f:
        f.write(text)
    mesh = spec["grid"](su2)
    mesh.write(os.path.join(work, "grid.su2"))
    config = rh.run_prep(work)
    merge(config, spec["edit"])
    with open(os.path.join(work, "config", "config.json"), "w") as f:
        json.dump(config, f, indent=1)
    if "start_from" in spec:
        for k, v in finals[spec["start_from"]].items():
            np.ascontiguousarray(v).tofile(os.path.join(work, "io/flow/0000", k))
    return work, mesh.dim


def main():
    rl = rh.RefLib.get(1)
    only = sys.argv[1:]
    finals = {}
    todo = [(n, None) for n in CASES] + list(EXAMPLE_CASES.items())
    for name, spec in todo:
        if only and name not in only and not any(EXAMPLE_CASES.get(o, {}).get("start_from") == name for o in only):
            continue
        if spec is None:
            job_dir, grid_dir = CASES[name]
            work = os.path.join("/tmp/ibis_dropin_golden", name)
            if os.path.exists(work):
                shutil.rmtree(work)
            os.makedirs(work)
            for d in (grid_dir, job_dir):
                for f in os.listdir(os.path.join(MINIMAL, d)):
                    shutil.copy(os.path.join(MINIMAL, d, f), os.path.join(work, f))
            config = rh.run_prep(work)
            dim = int(config["grid"]["dimensions"]) if "dimensions" in config["grid"] else (3 if "3D" in name else 2)
        else:
            work, dim = prepare_example(name, spec, finals)
        dst = os.path.join(HERE, "dropin", name)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        for k in KEEP:                                   # the prepared case, before the solve
            src = os.path.join(work, k)
            out = os.path.join(dst, k)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            if os.path.isdir(src):
                shutil.copytree(src, out)
            else:
                shutil.copy(src, out)
        import time
        t0 = time.time()
        if spec is not None and spec.get("prep_only"):
            print(name, "prepared (no reference solve)")
            continue
        rc = rl.solve(work)
        print(f"  reference solve: {time.time() - t0:.1f} s", flush=True)
        flow_dirs = sorted(d for d in os.listdir(os.path.join(work, "io/flow")) if d.isdigit())
        final = {"rc": np.array(rc), "n_snapshots": np.array(len(flow_dirs)), "dim": np.array(dim)}
        for k, v in rh.read_flow(work, int(flow_dirs[-1]), dim).items():
            final["final_" + k] = v
        with open(os.path.join(work, "io/flow", flow_dirs[-1], "meta_data.json")) as f:
            text = f.read()
        final["final_time"] = np.array(json.loads(text)["time"])
        # the bytes NativeBinaryOutput::write put on disk beside the arrays (tests/test_flow_io.py)
        final["final_meta_text"] = np.array(text)
        final["final_dir"] = np.array(flow_dirs[-1])
        with open(os.path.join(work, "io/flow/flows")) as f:
            final["flows_text"] = np.array(f.read())
        finals[name] = {k[6:]: v for k, v in final.items() if k.startswith("final_") and k != "final_time"}
        np.savez_compressed(os.path.join(dst, "expected.npz"), **final)
        print(name, "rc", rc, "snapshots", len(flow_dirs), "cells", len(final["final_T"]), "t", float(final["final_time"]))


if __name__ == "__main__":
    main()
