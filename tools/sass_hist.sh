#!/bin/bash
# tools/sass_hist.sh PATTERN [LIB]: opcode histogram of the SASS of the kernels whose mangled name
# matches PATTERN (static instruction mix: a first look at issue pressure without a GPU).
LIB=${2:-$(dirname "$0")/../ibis_b200/lib/libibis_b200.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk -v pat="$1" '
  /Function : /{p = ($0 ~ pat); if (p) print "== " $3; next}
  p && $1 ~ /^\/\*[0-9a-f]+\*\/$/ { op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;$/, "", op); n[op]++; t++ }
  END{ for (k in n) printf "%6d %s\n", n[k], k | "sort -rn"; close("sort -rn"); print t " instructions" }'
