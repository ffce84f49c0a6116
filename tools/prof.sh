#!/bin/bash
# tools/prof.sh TAG N [KERNEL_REGEX]: the two ncu passes of /opt/skills/guides/B200_PROFILING.md over
# tools/profile_step.py (run under gpurun, one GPU).  Writes
#   gpurun_out/TAG_launches_nN.csv   every launch with its device time
#   gpurun_out/TAG_full.ncu-rep      --set full capture of pass G / pass F (one launch each)
#   gpurun_out/TAG_full_raw.csv      its raw page
TAG=${1:-prof}; N=${2:-128}; RX=${3:-'k_residual|k_grad_limit|k_eval'}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches_n${N}.csv python tools/profile_step.py --n $N --steps 2 \
    > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 2 -c 2 -f \
    -o gpurun_out/${TAG}_full python tools/profile_step.py --n $N --steps 2 > gpurun_out/${TAG}_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
