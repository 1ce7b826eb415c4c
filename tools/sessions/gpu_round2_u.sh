# interleaved A/B (3 repetitions): layer1 block kernel with shifted taps / last block on the block kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2 3; do
for v in "X=0" "BV_L1_SH=1" "BV_L1_SH=1 BV_L1_LAST=1"; do
env $v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-28s' % '$v', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), round(d['roofline']['frac'],4), d['config'].get('gathered_checksum'))"
done
done
