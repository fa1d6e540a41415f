#!/bin/bash
# In-box A/B of two settings (boxes differ by +-2 % and sustained runs are power capped, so only runs on the SAME box compare):
#   tools/ab_bench.sh "ENV_A=1" "ENV_B=1" [reps] [steps]        e.g.  tools/ab_bench.sh "" "RY_DW5_FFMA=1" 3 50
# or two builds:  put the alternative library next to the built one as rep-yolo_b200/csrc/alt_lib.so and pass "LIB=alt".
A=${1:-}; B=${2:-}; REPS=${3:-3}; STEPS=${4:-50}
SO=rep-yolo_b200/csrc/librepyolo_b200.so
cp $SO rep-yolo_b200/csrc/main_lib.so
run() {
  local tag=$1 envs=$2
  if [ "$envs" = "LIB=alt" ]; then cp rep-yolo_b200/csrc/alt_lib.so $SO; envs=""; else cp rep-yolo_b200/csrc/main_lib.so $SO; fi
  env $envs python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline 2>/dev/null |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', round(d['value']), 'img/s', round(d['e2e']['value']), 'e2e', d['clocks']['sm_mhz'], 'MHz', d['clocks']['reasons'])"
}
for i in $(seq $REPS); do run "A[$A]" "$A"; run "B[$B]" "$B"; done
cp rep-yolo_b200/csrc/main_lib.so $SO
