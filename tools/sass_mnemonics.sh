#!/bin/bash
# per-kernel counts of the SASS mnemonics that tell Blackwell-native code from legacy code (B200_PROFILING.md table) -> stdout
LIB=${1:-instant-nsr-pl_b200/libnsr_b200.so}
echo "# cuobjdump -sass $LIB: instruction counts per kernel"
echo "# UTCHMMA = tcgen05.mma | UTCBAR = tcgen05.commit | LDTM = tcgen05.ld | UBLKCP = cp.async.bulk (TMA bulk copy) | SYNCS = mbarrier"
echo "# HMMA = mma.sync (legacy tensor path) | REDG = red.global | LDGSTS = cp.async | LDGMC / STGMC / REDGMC = multimem.ld_reduce / st (NVSwitch)"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
/Function :/ {fn=$3}
{ for (i = 1; i <= NF; ++i) if ($i ~ /^(UTCHMMA|UTCBAR|LDTM|UBLKCP|UTMALDG|UTMASTG|SYNCS|HMMA|REDG|LDGMC|STGMC|REDGMC|LDGSTS)/) { split($i, b, "."); c[fn "\t" b[1]]++ } }
END { for (k in c) print k "\t" c[k] }' | sort | while IFS=$'\t' read -r fn mn n; do
  printf "%-12s %6d  %s\n" "$mn" "$n" "$(echo "$fn" | c++filt | sed 's/(anonymous namespace):://g; s/(nsr_.*//; s/(float.*//; s/(void.*//; s/(int.*//; s/(long.*//; s/(__half.*//; s/(unsigned.*//')"
done
