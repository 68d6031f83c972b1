#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the smoke path -- K1 relabel, K5 cost normalisation, K3 dual GAE,
# K4 persistent cluster kernel (TMA bulk copies, DSMEM st.async + mbarrier exchanges) + dual step, K2 (six multi-CTA kernels,
# ticket / flag exchanges) -- and memcheck over the same path with the many-cluster K4 kernel forced (grid barriers through L2).
# Usage (under gpurun, from the repo root):  bash tools/sanitize.sh r02     -> gpurun_out/sanitize_*_r02.log + .summary
set -u
R=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/sanitize_$R.summary
for tool in memcheck racecheck synccheck; do
    timeout 420 compute-sanitizer --tool $tool --print-limit 30 python __graft_entry__.py smoke > $OUT/sanitize_${tool}_$R.log 2>&1
    echo "== $tool (single-cluster K4): exit $?" >> $OUT/sanitize_$R.summary
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|error" $OUT/sanitize_${tool}_$R.log | tail -6 >> $OUT/sanitize_$R.summary
done
ICRL_PPO_WIDE=1 ICRL_PPO_WIDE_CLUSTERS=6 timeout 420 compute-sanitizer --tool memcheck --print-limit 30 \
    python __graft_entry__.py smoke > $OUT/sanitize_memcheck_wide_$R.log 2>&1
echo "== memcheck (many-cluster K4 forced, 6 clusters): exit $?" >> $OUT/sanitize_$R.summary
grep -E "ERROR SUMMARY|smoke ok|Error|error" $OUT/sanitize_memcheck_wide_$R.log | tail -6 >> $OUT/sanitize_$R.summary
cat $OUT/sanitize_$R.summary
