for rep in 1 2; do
for lib in libcfnmpc_prev.so libcfnmpc.so; do for keep in 0 auto; do
  if [ "$keep" = auto ]; then unset CFNMPC_L2_KEEP; else export CFNMPC_L2_KEEP=$keep; fi
  CFNMPC_LIB=$PWD/crazyflie_nmpc_b200/$lib CFNMPC_MIN_BLOCKS=3 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib keep $keep', d['config']['occupancy']['warps_per_sm'], round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
done; done; done
