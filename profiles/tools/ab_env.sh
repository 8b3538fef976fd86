# A/B of environment-selected variants of the same build on the same box, alternating, with the real benchmark.
# Usage (GPU box): bash profiles/tools/ab_env.sh "VAR1=a VAR2=b" "VAR1=c" ...   (an empty string = the defaults)
for rep in 1 2; do
for cfg in "$@"; do
  env $cfg python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$cfg]', d['config']['occupancy'], round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], d.get('gpu_launches'))"
done; done
