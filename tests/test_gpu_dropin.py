"""GPU test (-m gpu): the drop-in.  oracle/_ref/libibis_ref_b200.so is the UNMODIFIED reference (its Solver::solve,
RungeKutta, SteadyState / Jfnk / Gmres, JSON configuration, native IO, compiled from /root/reference by
oracle/ref_build/Makefile) with its finite_volume.cpp replaced at link time by ibis_b200/adapter/b200_finite_volume.cpp:
the same class FiniteVolume<T>, every method forwarding to the C ABI of include/ibis_b200.h.

The reference's own tests/minimal cases (uniform_stream, uniform_stream_3D, steady_state_uniform_stream), prepared by
the reference's own `ibis prep` and committed under tests/golden/dropin (tests/golden/make_dropin_cases.py), are run
through that library -- the reference's solver loop driving the CUDA residual -- and the final snapshot is compared
with the one the stock reference (Kokkos-OpenMP) wrote: bit for bit for the explicit runs, 1e-9 for the steady one
(its Krylov dot products are summed in another order on the device).  When the stock library travelled to this box
too, it is run here as well and its snapshot must equal the committed one."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
DROPIN = os.path.join(REF_DIR, "libibis_ref_b200.so")
STOCK = os.path.join(REF_DIR, "libibis_ref.so")
PADMALLOC = os.path.join(REF_DIR, "libpadmalloc.so")
CASES = os.path.join(HERE, "golden", "dropin")

RUNNER = r"""
import ctypes, os, sys
lib = ctypes.CDLL(sys.argv[1], mode=ctypes.RTLD_GLOBAL)
lib.ref_init.argtypes = [ctypes.c_int]
lib.ref_solve.argtypes = [ctypes.c_char_p]
lib.ref_solve.restype = ctypes.c_int
lib.ref_init(1)
rc = lib.ref_solve(os.fsencode(sys.argv[2]))
sys.stdout.flush()
os._exit(0 if rc == 0 else 10 + abs(rc))          # skip Kokkos / CUDA teardown order at interpreter exit
"""


def solve_with(lib_path, case, tmp_path, tag, preload=None):
    work = os.path.join(str(tmp_path), f"{case}_{tag}")
    shutil.copytree(os.path.join(CASES, case), work)
    os.makedirs(os.path.join(work, "log"), exist_ok=True)
    env = dict(os.environ)
    # oracle/ref_build/padmalloc.c pads every allocation by 32 bytes: the reference's Gmres stores one Hessenberg
    # column past its arrays when max_iters is reached (and its CubicSpline one element), which the stock build
    # survives by luck of the heap layout and this one -- more host allocations in between -- does not
    pre = [p for p in (preload, PADMALLOC if os.path.exists(PADMALLOC) else None) if p]
    if pre:
        env["LD_PRELOAD"] = " ".join(pre)
    r = subprocess.run([sys.executable, "-c", RUNNER, lib_path, work], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, f"{tag} solve of {case} failed: rc {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    flow = os.path.join(work, "io", "flow")
    dirs = sorted(d for d in os.listdir(flow) if d.isdigit())
    expected = np.load(os.path.join(CASES, case, "expected.npz"))
    names = ["T", "p", "vx", "vy"] + (["vz"] if int(expected["dim"]) == 3 else [])
    return {k: np.fromfile(os.path.join(flow, dirs[-1], k)) for k in names}, len(dirs), r.stderr


# The reference's EXAMPLES at about their own resolution (job.py unchanged through the reference's prep; what was
# edited in config.json afterwards is listed in tests/golden/make_dropin_cases.py): 150 ssp-rk3 steps of the Mach 3
# wedge (5 400 cells) and of the forward-facing step (4 032 cells), 60 steps of the viscous NACA 0012 (3 072 cells,
# subsonic in/outflow, no-slip wall).
EXAMPLES = {"example_wedge_transient": 0.0, "example_forward_facing_step": 0.0, "example_naca0012_viscous": 1e-11}


def compare(case, got, expected, rtol):
    for k, v in got.items():
        want = expected["final_" + k]
        if rtol == 0.0:
            assert np.array_equal(v, want), f"{case}: {k} differs from the stock reference's snapshot"
        else:
            scale = max(1.0, float(np.abs(want).max()))
            assert np.allclose(v, want, rtol=rtol, atol=rtol * scale), \
                f"{case}: {k} off by {np.max(np.abs(v - want)) / scale:.3e} (relative to the field's scale)"


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(EXAMPLES))
def test_reference_examples_on_the_b200_residual(case, tmp_path):
    import torch
    assert torch.cuda.is_available()
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libibis_ref_b200.so not built (needs /root/reference at build time)")
    expected = np.load(os.path.join(CASES, case, "expected.npz"))
    got, n_snap, log = solve_with(DROPIN, case, tmp_path, "b200")
    assert "[ibis_b200 adapter] FiniteVolume<" in log
    assert n_snap == int(expected["n_snapshots"])
    compare(case, got, expected, EXAMPLES[case])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["uniform_stream", "uniform_stream_3D", "steady_state_uniform_stream"])
def test_reference_solver_loop_on_the_b200_residual(case, tmp_path):
    import torch
    assert torch.cuda.is_available()
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libibis_ref_b200.so not built (needs /root/reference at build time)")
    expected = np.load(os.path.join(CASES, case, "expected.npz"))
    got, n_snap, log = solve_with(DROPIN, case, tmp_path, "b200")
    assert "[ibis_b200 adapter] FiniteVolume<" in log, "the adapter did not bind the residual to libibis_b200.so"
    assert n_snap == int(expected["n_snapshots"])
    for k, v in got.items():
        want = expected["final_" + k]
        if case.startswith("steady"):
            assert np.allclose(v, want, rtol=1e-9, atol=1e-9 * max(1.0, float(np.abs(want).max()))), k
        else:
            assert np.array_equal(v, want), f"{case}: {k} differs from the stock reference's snapshot"
    if os.path.exists(STOCK):
        stock, n_stock, _ = solve_with(STOCK, case, tmp_path, "stock")
        assert n_stock == n_snap
        for k, v in stock.items():
            assert np.array_equal(v, expected["final_" + k]), f"{case}: the stock reference does not reproduce its fixture ({k})"


@pytest.mark.parametrize("case", ["uniform_stream", "uniform_stream_3D", "steady_state_uniform_stream"])
def test_reference_solver_loop_on_the_host_emulation(case, tmp_path):
    """CPU: the same drop-in library with the host-emulation build of the C ABI preloaded (tests/emu: the kernels'
    bodies run as plain loops) -- the reference's solver loops, the adapter and the ABI without a GPU.  Every snapshot,
    the steady one included (serial dot products here), equals the stock reference's bit for bit."""
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libibis_ref_b200.so not built (needs /root/reference at build time)")
    from emu_util import emu_lib
    expected = np.load(os.path.join(CASES, case, "expected.npz"))
    got, n_snap, log = solve_with(DROPIN, case, tmp_path, "emu", preload=emu_lib())
    assert "[ibis_b200 adapter] FiniteVolume<" in log
    assert n_snap == int(expected["n_snapshots"])
    for k, v in got.items():
        assert np.array_equal(v, expected["final_" + k]), f"{case}: {k} differs from the stock reference's snapshot"


def test_dropin_library_resolves_every_symbol():
    """CPU: the drop-in library links completely (no FiniteVolume symbol left to the reference's own translation unit)."""
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libibis_ref_b200.so not built (needs /root/reference at build time)")
    r = subprocess.run(["ldd", "-r", DROPIN], capture_output=True, text=True)
    undefined = [ln for ln in (r.stdout + r.stderr).splitlines() if "undefined symbol" in ln and "FiniteVolume" in ln]
    assert not undefined, undefined
    nm = subprocess.run(["nm", "-DC", DROPIN], capture_output=True, text=True).stdout
    assert "ref_solve" in nm and "FiniteVolume<double>::compute_dudt" in nm and "ibis_b200_compute_dudt" in nm
