#!/bin/bash
# tools/sweep.sh N "tag[:ENV=VAL,...]" ... : bench.py (no CPU arm) for each built variant library
# (tools/build_variants.sh; tag "default" = ibis_b200/lib) and environment setting; prints one line
# per run: pass G / pass F launch times, evaluation roofline fraction, evals/s.  Run under gpurun.
N=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%:*}"; envs=""; [[ "$spec" == *:* ]] && envs="${spec#*:}"
  lib="ibis_b200/lib/variants/$tag/libibis_b200.so"; [ "$tag" = default ] && lib="ibis_b200/lib/libibis_b200.so"
  out="gpurun_out/sweep_${tag}_$(echo "$envs" | tr ',=' '__').json"
  env IBIS_B200_LIB=$lib $(echo "$envs" | tr ',' ' ') python bench.py --n $N --steps 4 --warmup 3 --no-cpu --no-secondary > "$out" 2> "$out.err" || { echo "$spec FAILED"; tail -3 "$out.err"; continue; }
  python - "$out" "$spec" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print(f"{sys.argv[2]:40s} G {r['pass_G']['avg_launch_ms']:7.3f} ms  F {r.get('pass_F', r)['avg_launch_ms']:7.3f} ms  eval {100 * r['evaluation']['frac']:5.1f} %  "
      f"value {d['value'] / 1e9:6.3f} G/s  step {d['ms_per_step']:7.2f} ms  e2e {d['e2e']['value'] / 1e9:5.2f} G/s")
PY
done
