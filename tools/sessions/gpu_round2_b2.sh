# l1_block biases in the constant bank (by-value kernel parameter) instead of shared memory: bit-exact tests, then A/B against the
# previous build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_l1_block_gpu.py tests/test_model_gpu.py tests/test_scorer_pins_gpu.py -q -x 2>&1 | tail -3
for rep in 1 2 3; do
for v in "BV_LIB_PATH=$PWD/build/libbiovil_b200_prev.so" "X=0"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2b2_table.csv > gpurun_out/r2b2_bench.json 2>gpurun_out/r2b2_bench.err
echo "== ${v##*/}"; grep -E "l1_block" gpurun_out/r2b2_table.csv | cut -d, -f2 | tr '\n' ' '
python -c "
import json; d=json.load(open('gpurun_out/r2b2_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('gathered_checksum'))"
done
done
