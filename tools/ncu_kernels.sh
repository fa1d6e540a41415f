#!/bin/bash
# ncu --set full of selected launches of one forward pass:  tools/ncu_kernels.sh <tag> <kernel-regex>:<skip>:<count> ...
TAG=$1; shift
PROF="python tools/profile_ops.py --batch 64 --reps 1"
$PROF > gpurun_out/plain_ncuk.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ncuk.log; exit 1; }
for SPEC in "$@"; do
  IFS=: read KRE S C <<< "$SPEC"
  ncu --set full --clock-control none --import-source on -k regex:$KRE -s $S -c $C -f -o gpurun_out/prof_${TAG}_${KRE}_s$S $PROF > gpurun_out/ncu_${TAG}_${KRE}_s$S.log 2>&1
  echo "ncu $KRE skip=$S count=$C rc=$?"
done
