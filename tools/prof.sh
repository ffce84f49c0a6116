#!/bin/bash
# tools/prof.sh TAG [N]: ncu launch list of a short bench run + one --set full capture of the
# two residual kernels.  Run under gpurun; outputs land in gpurun_out/.
TAG=${1:-r01}; N=${2:-128}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --n $N --steps 2 > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_residual|k_grad_limit' -s 2 -c 2 \
    -f -o gpurun_out/${TAG}_full python tools/profile_step.py --n $N --steps 2 > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/
