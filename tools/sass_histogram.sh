#!/bin/bash
# SASS opcode histogram of the in-tree libsnk.so (cuobjdump runs without a GPU): bash tools/sass_histogram.sh > profiles/<tag>_sass_opcodes.txt
so=${1:-marl-snake_b200/libsnk.so}
echo "# cuobjdump -sass $so  (sm_100a), opcode counts over all kernels; $(date -u +%F)"
echo "# kernels:"; cuobjdump -sass $so | grep -oE "Function : [A-Za-z0-9_]+" | sed 's/Function : //' | c++filt | sed 's/(.*//' | sort | uniq -c | sed 's/^/#   /'
echo "# proof points: UBLKCP = cp.async.bulk (TMA 1-D bulk copies), SYNCS = mbarrier, MATCH = match.any, REDUX/VOTE = warp votes,"
echo "#   DMUL/DADD without DFMA = float64 reward in the reference's order, no LDL/STL = no local memory"
for op in UBLKCP SYNCS MATCH VOTE REDUX DMUL DADD DFMA LDL STL UTCHMMA UTCQMMA; do
  printf "# %-8s %s\n" $op $(cuobjdump -sass $so | grep -cE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?$op(\.|\s|;)")
done
echo "# per benchmarked instance (cuobjdump -res-usage: registers, stack) and its local-memory instructions:"
for inst in "Li4ELi20ELi11ELi11ELi1ELb0ELi1ELi3E cfg5_bench_line" "Li4ELi20ELi11ELi11ELi1ELb0ELi1ELi0E cfg5_shard" "Li4ELi20ELi20ELi20ELi1ELb1ELi2ELi0E cfg2" "Li4ELi20ELi11ELi11ELi4ELb1ELi0ELi0E cfg3" "Li16ELi64ELi15ELi15ELi1ELb1ELi3ELi0E cfg4"; do
  set -- $inst
  res=$(cuobjdump -res-usage $so 2>/dev/null | grep -A1 "$1" | grep -oE "REG:[0-9]+ STACK:[0-9]+")
  loc=$(cuobjdump -sass -fun "_ZN3snk15snk_tile_kernelI${1}EvNS_7KParamsE" $so 2>/dev/null | grep -cE "\b(LDL|STL)\b")
  printf "#   %-16s %-22s LDL+STL %s\n" $2 "$res" $loc
done
cuobjdump -sass $so | grep -oE "^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P[0-9T] )?[A-Z][A-Z0-9_.]+" | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn
