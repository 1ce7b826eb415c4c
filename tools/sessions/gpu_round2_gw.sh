# l1_block: one polling epilogue warp + named barrier for the other fifteen (-DBV_L1_GROUP_WAIT=1) against the default build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
BV_LIB_PATH=$PWD/build/libbiovil_b200_gw.so timeout 300 python -m pytest tests/test_l1_block_gpu.py -q -x 2>&1 | tail -2
for rep in 1 2 3; do
for v in "X=0" "BV_LIB_PATH=$PWD/build/libbiovil_b200_gw.so"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2gw_table.csv > gpurun_out/r2gw_bench.json 2>gpurun_out/r2gw_bench.err
echo "== ${v##*/}"; grep -E "l1_block" gpurun_out/r2gw_table.csv | cut -d, -f2 | tr '\n' ' '
python -c "
import json; d=json.load(open('gpurun_out/r2gw_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('gathered_checksum'))"
done
done
