#!/bin/bash
# Run on the GPU box (via gpurun): GPU parity tests, a short bench, the ncu launch list of the same bench command and
# one `ncu --set full` capture of the dominant kernel.  Outputs land in gpurun_out/.
#   tools/gpu_check.sh [tag] [kernel-regex] [skip-launches]
set -u
TAG=${1:-r01}
KRE=${2:-conv_umma_kernel}
SKIP=${3:-0}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gputests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
python tools/profile_ops.py --out gpurun_out/ops_$TAG.txt > /dev/null 2> gpurun_out/prof_$TAG.err; echo "profile_ops rc=$?"
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-nms-legs"
$BENCH > gpurun_out/plain_$TAG.log 2>&1
# launches per step from the bench line itself (gpu_launches / steps); the launch list covers two timed steps after the warm-up
LPS=$(python -c "import json,sys; d=json.loads(open('gpurun_out/plain_$TAG.log').read().strip().splitlines()[-1]); print(d['gpu_launches']//d['steps'])")
echo "launches per step: $LPS"
ncu --metrics gpu__time_duration.sum --clock-control none -s $((6 * LPS)) -c $((2 * LPS)) --csv --log-file gpurun_out/launches_$TAG.csv $BENCH > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
PROF="python tools/profile_ops.py --batch 64 --reps 1"
$PROF > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 3 -f -o gpurun_out/prof_$TAG $PROF > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/gputests_$TAG.log
# DRAM traffic of every conv launch of ONE forward pass (the third: two warm-up passes are skipped) -> tools/conv_traffic.py
NCONV=${4:-109}
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"conv_umma_kernel|conv_chain_kernel" \
    -s $((2 * NCONV)) -c $NCONV --csv --log-file gpurun_out/conv_traffic_$TAG.csv $PROF > gpurun_out/ncu_traffic_$TAG.log 2>&1
echo "ncu conv traffic rc=$?"
# tensor-pipe utilisation from the UTCHMMA instruction counter (tools/tensor_util.py)
ncu --metrics sm__inst_executed_pipe_tensor_subpipe_hmma.sum,sm__cycles_elapsed.avg,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum --clock-control none \
    -k regex:"conv_umma_kernel|conv_chain_kernel" -s $((2 * NCONV)) -c $NCONV --csv --log-file gpurun_out/conv_tensor_$TAG.csv $PROF > gpurun_out/ncu_tensor_$TAG.log 2>&1
echo "ncu tensor rc=$?"
python tools/nms_bench.py --out gpurun_out/nms_bench_$TAG.txt > /dev/null 2>&1; echo "nms_bench rc=$?"
python tools/bench_configs.py > gpurun_out/configs_$TAG.txt 2>&1; echo "bench_configs rc=$?"
