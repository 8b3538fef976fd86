# A/B of two builds on the same box: crazyflie_nmpc_b200/libcfnmpc_prev.so (older build, copied there by hand) against
# the current libcfnmpc.so, alternating, with the real benchmark.  Usage (GPU box): bash profiles/tools/ab_libs.sh [MINB]
MB=${1:-3}
for rep in 1 2; do
for lib in libcfnmpc_prev.so libcfnmpc.so; do
  CFNMPC_LIB=$PWD/crazyflie_nmpc_b200/$lib CFNMPC_MIN_BLOCKS=$MB python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', d['config']['occupancy'], round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'])"
done; done
