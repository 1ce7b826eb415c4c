# pair-chained kernel on by default for layer3 (bit 0); do the other shape classes pay now? (bit 1: layer2 -> layer3, bit 2: layer2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for v in "X=0" "BV_PAIR_CHAIN=3" "BV_PAIR_CHAIN=5"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2pc3_table.csv > gpurun_out/r2pc3_bench.json 2>gpurun_out/r2pc3_bench.err
echo "== $v"; grep -E "pair_chain|chain_gemm|c2>|N=256 K=512" gpurun_out/r2pc3_table.csv | awk -F, '{n[$1]++; s[$1]+=$2} END {for (k in n) printf "%s x%d %.4f | ", substr(k,1,48), n[k], s[k]/n[k]}'; echo
python -c "
import json; d=json.load(open('gpurun_out/r2pc3_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), d['config'].get('gathered_checksum'))"
done
done
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/stress_forward.py 3000 2>&1 | tail -1
