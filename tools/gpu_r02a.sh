#!/bin/bash
# round 2, first GPU call: parity tests on the hygiene changes, smoke, a short bench, per-op profile, and the list of ncu
# metrics that could count tcgen05.mma (UTCHMMA) work.
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gputests_$TAG.log
tail -5 gpurun_out/gputests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
python tools/profile_ops.py --out gpurun_out/ops_$TAG.txt > /dev/null 2> gpurun_out/prof_$TAG.err; echo "profile_ops rc=$?"
tail -6 gpurun_out/ops_$TAG.txt
ncu --query-metrics > gpurun_out/ncu_query_metrics.txt 2>&1; echo "query rc=$?"
grep -ciE 'tensor|utc|tmem|pipe_tc' gpurun_out/ncu_query_metrics.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu.txt 2>&1
