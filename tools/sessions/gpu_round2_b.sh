set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py -q -x -k "pair" 2>&1 | tail -30 > gpurun_out/r2b_pair_tests.log
tail -15 gpurun_out/r2b_pair_tests.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2b_pytest.log
tail -6 gpurun_out/r2b_pytest.log
for v in "BV_PAIR=0" "BV_PAIR=3" "BV_PAIR=1" "BV_PAIR=2" "BV_PAIR=0" "BV_PAIR=3"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2b_table_${v#BV_PAIR=}.csv 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), round(d['roofline']['frac'],4))"
done
