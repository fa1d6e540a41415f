#!/bin/bash
# ncu --set full of selected conv launches of one forward pass:  tools/ncu_ops.sh <tag> <kernel-regex> <skip1> [skip2 ...]
TAG=$1; KRE=$2; shift 2
PROF="python tools/profile_ops.py --batch 64 --reps 1"
$PROF > gpurun_out/plain_ncuops.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_ncuops.log; exit 1; }
for S in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$KRE -s $S -c 1 -f -o gpurun_out/prof_${TAG}_s$S $PROF > gpurun_out/ncu_${TAG}_s$S.log 2>&1
  echo "ncu skip=$S rc=$?"
done
