#!/bin/bash
# tools/pf_sweep.sh TAG N: L2 prefetch distance sweep of the pipelined kernels (run under gpurun): bench times and
# DRAM bytes of one launch of each kernel per setting
TAG=${1:-pf}; N=${2:-256}
cd "$(dirname "$0")/.."
for pf in 0 2 4 6 12; do
  IBIS_B200_PREFETCH=$pf timeout 300 python bench.py --n $N --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_pf${pf}.json 2> gpurun_out/${TAG}_pf${pf}.err
  python -c "
import json;d=json.load(open('gpurun_out/${TAG}_pf${pf}.json'));r=d['roofline'];print('pf',$pf,'G',round(r['pass_G']['avg_launch_ms'],3),'F',round(r['pass_F']['avg_launch_ms'],3),'eval frac',round(r['evaluation']['frac'],4))"
  IBIS_B200_PREFETCH=$pf ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_grad_pipe|k_flux_pipe" -s 2 -c 2 --csv python tools/profile_step.py --n 160 --steps 2 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | tr -d '"' | sed "s/^/   pf$pf /"
done
