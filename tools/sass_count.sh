#!/bin/bash
# static SASS instruction mix per kernel of the built library (no GPU needed)
LIB=${1:-ionization_b200/_lib/libionization_b200.so}
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ {name=$3}
/^ +\/\*[0-9a-f]+\*\/ / {n[name]++; if ($0 ~ /DFMA|DMUL|DADD/) d[name]++; if ($0 ~ /SHFL/) s[name]++; if ($0 ~ /BAR\./) b[name]++; if ($0 ~ /MUFU/) m[name]++; if ($0 ~ /LDG|STG/) g[name]++; if ($0 ~ /LDS|STS/) l[name]++; if ($0 ~ /LDL|STL/) sp[name]++}
END {for (k in n) printf "%-64s total %5d fp64 %5d shfl %4d bar %3d mufu %3d ldg/stg %3d lds/sts %3d local %3d\n", k, n[k], d[k], s[k], b[k], m[k], g[k], l[k], sp[k]}' | sort
