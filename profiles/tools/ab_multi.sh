# A/B of several builds on the same box, alternating, with the real benchmark.
# Usage (GPU box): bash profiles/tools/ab_multi.sh <reps> name1 name2 ...   (names of crazyflie_nmpc_b200/variants/libcfnmpc_<name>.so)
reps=$1; shift
for rep in $(seq $reps); do
for v in "$@"; do
  CFNMPC_LIB=$PWD/crazyflie_nmpc_b200/variants/libcfnmpc_$v.so python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['e2e']['value']), d['roofline']['kernels_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
