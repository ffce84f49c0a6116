#!/bin/bash
# tools/sanitize_emu.sh -- host-side memory / UB check of the C-ABI layer (abi.cu, layout.h, grid_build.cu,
# mesh_io.cu) and of every kernel body: builds the host-emulation harness with AddressSanitizer +
# UndefinedBehaviorSanitizer and runs the CPU parity tests on it.  Last run (end of round 1): no report.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${OUT:-/tmp/ibis_b200_asan}
mkdir -p "$OUT"
SRC=$ROOT/ibis_b200/csrc
/usr/bin/g++ -O1 -g -std=c++17 -fPIC -fopenmp -ffp-contract=off -DIBIS_B200_EMU -fsanitize=address,undefined \
    -fno-omit-frame-pointer -I"$SRC" -I"$ROOT/include" -w -shared -o "$OUT/libibis_b200_emu.so" \
    "$ROOT/tests/emu/emu_kernels.cpp" -x c++ "$SRC/abi.cu" "$SRC/grid_build.cu" "$SRC/mesh_io.cu"
cat > "$OUT/run.py" <<PY
import sys
sys.path[:0] = ["$ROOT", "$ROOT/tests"]
import emu_util
emu_util.emu_lib = lambda: "$OUT/libibis_b200_emu.so"
import pytest
if __name__ == "__main__":
    t = "$ROOT/tests/"
    sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider", "-m", "not gpu", "-k", "not gloo and not host_solver and not host_steady"] +
                         [t + f for f in ("test_emu_parity.py", "test_boundary_actions.py", "test_grid_build.py",
                                          "test_moving_grid.py", "test_reference_kats.py", "test_decomposition.py",
                                          "test_flow_io.py", "test_oracle_restatement.py", "test_properties.py")]))
PY
ASAN=$(/usr/bin/g++ -print-file-name=libasan.so)
UBSAN=$(/usr/bin/g++ -print-file-name=libubsan.so)
LD_PRELOAD=$ASAN:$UBSAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 OMP_NUM_THREADS=2 \
    python "$OUT/run.py" 2>&1 | tee "$OUT/log.txt" | tail -3
echo "UBSan reports: $(grep -c 'runtime error' "$OUT/log.txt" || true)"
