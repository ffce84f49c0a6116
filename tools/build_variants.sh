#!/bin/bash
# tools/build_variants.sh "tag:-DIBIS_F_MINB=4 ..." ... : builds ibis_b200/lib/variants/<tag>/libibis_b200.so
# for tuning sweeps (tools/sweep.sh).  The default library is unaffected.
cd "$(dirname "$0")/../ibis_b200/csrc"
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  out="../lib/variants/$tag"
  mkdir -p "$out"
  ( make -s -j4 OUT="$out" TUNE="$flags" > "$out/build.log" 2>&1 || { echo "build of $tag failed"; tail -5 "$out/build.log"; } ) &
done
wait
ls ../lib/variants/*/libibis_b200.so
