# pair-chained layer3 kernel with biases in the constant bank: bit-exact test, then A/B of BV_PAIR_CHAIN=0/1 (3 interleaved reps)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_chain_gpu.py -q -x 2>&1 | tail -2
for rep in 1 2 3; do
for v in "BV_PAIR_CHAIN=0" "BV_PAIR_CHAIN=1"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2pc2_table.csv > gpurun_out/r2pc2_bench.json 2>gpurun_out/r2pc2_bench.err
echo "== $v"; grep -E "pair_chain|N=1024 K=256 \+res|M=460800 N=256 K=1024" gpurun_out/r2pc2_table.csv | awk -F, '{n[$1]++; s[$1]+=$2} END {for (k in n) printf "%s x%d %.4f | ", substr(k,1,48), n[k], s[k]/n[k]}'; echo
python -c "
import json; d=json.load(open('gpurun_out/r2pc2_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']), d['config'].get('gathered_checksum'))"
done
done
