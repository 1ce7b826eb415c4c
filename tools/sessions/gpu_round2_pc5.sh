# ring / staging split of the layer3 pair-chained kernel, re-measured on the final binary: variant 0 (4 stages + 6 staging sub-tiles,
# default), 1 (5 + 5), 2 (3 + 7); then the model-level parity tests with variant 1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for v in "BV_PC_VARIANT=0" "BV_PC_VARIANT=1" "BV_PC_VARIANT=2"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2pc5_table.csv > gpurun_out/r2pc5_bench.json 2>gpurun_out/r2pc5_bench.err
echo "== $v"; grep -E "pair_chain" gpurun_out/r2pc5_table.csv | cut -d, -f2 | tr '\n' ' '
python -c "
import json; d=json.load(open('gpurun_out/r2pc5_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('gathered_checksum'))"
done
done
BV_PC_VARIANT=1 timeout 300 python -m pytest tests/test_model_gpu.py tests/test_scorer_pins_gpu.py tests/test_chain_gpu.py -q -x 2>&1 | tail -2
