cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_l1_block_gpu.py -q -x 2>&1 | tail -6
for v in "X=0" "BV_L1_SPLIT=1" "BV_L1_SPLIT=1 BV_L1_SH=1 BV_L1_LAST=1"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2q_table.csv > gpurun_out/r2q_bench.json 2>/dev/null
echo "== $v"; grep -E "l1_block|tap3|c7|chain_gemm<128> M=7372800" gpurun_out/r2q_table.csv
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
