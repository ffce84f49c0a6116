#!/bin/bash
# tools/sweep.sh N tag...: bench.py --n N for each library variant (and the default library), one line each
N=${1:-160}; shift
for tag in default "$@"; do
  if [ "$tag" = default ]; then unset IBIS_B200_LIB; else export IBIS_B200_LIB=ibis_b200/lib/variants/$tag/libibis_b200.so; fi
  python bench.py --steps 4 --warmup 3 --n $N --no-cpu 2> gpurun_out/sweep_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('%-14s n=$N value %.3e  F %.3f ms (%.3f)  G %.3f ms  eval frac %.3f  step %.2f ms' % ('$tag', d['value'], r['avg_launch_ms'], r['frac'], r['pass_G']['avg_launch_ms'], r['evaluation']['frac'], d['ms_per_step']))" || tail -3 gpurun_out/sweep_$tag.err
done
