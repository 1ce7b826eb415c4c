# A/B of two library builds: default (2-stage A ring in l1_block) vs -DBV_L1_STAGES=3 (one staging sub-tile less)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BV_LIB_PATH=$PWD/build/libbiovil_b200_a3.so timeout 600 python -m pytest tests/test_l1_block_gpu.py -q -x 2>&1 | tail -3
for rep in 1 2; do
for v in "X=0" "BV_LIB_PATH=$PWD/build/libbiovil_b200_a3.so"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2x_table.csv > gpurun_out/r2x_bench.json 2>gpurun_out/r2x_bench.err
echo "== $v"; grep -E "l1_block" gpurun_out/r2x_table.csv
python -c "
import json; d=json.load(open('gpurun_out/r2x_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('gathered_checksum'))"
done
done
