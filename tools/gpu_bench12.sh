#!/bin/bash
# bench.py at N=1 and (when the box has 2 GPUs) N=2, plus the reference arm under torchrun (thread-count check)
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'list_api',round(d['e2e']['list_api']['images_per_s_per_gpu']),'frac',round(d['roofline']['frac'],3))
for k,v in d['roofline']['hbm_classes'].items(): print(' ',k,round(v['ms_per_step'],3),'ms',round(v['frac_of_hbm_peak'],3))
print(d['nms']); print(d['cpu_baseline'])
PY
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_2gpu.json 2> gpurun_out/bench_${TAG}_2gpu.err; echo "bench2 rc=$?"
  tail -3 gpurun_out/bench_${TAG}_2gpu.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}_2gpu.json').read().strip().splitlines()[-1])
print('N=2 value',round(d['value']),'e2e',round(d['e2e']['value']),'gather_ok',d.get('gather_ok'),d['config'].get('numa'))
PY
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-400
fi
