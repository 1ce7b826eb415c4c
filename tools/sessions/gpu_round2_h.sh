cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_l1_block_gpu.py tests/test_model_gpu.py -q -x 2>&1 | tail -4
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2h_table.csv > gpurun_out/r2h_bench.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['roofline']['frac'],4))"
grep "l1_block" gpurun_out/r2h_table.csv
done
