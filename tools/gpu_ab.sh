# usage: bash tools/gpu_ab.sh "ENV_A" "ENV_B" ... ; runs each setting twice, interleaved
cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for v in "$@"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-28s' % '$v', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), round(d['roofline']['frac'],4))"
done
done
