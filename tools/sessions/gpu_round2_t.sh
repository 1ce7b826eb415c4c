# short-distance L2 prefetch of the identity stream: conv_gemm +res / pair_chain (BV_RES_PREFETCH), chain_gemm (BV_L2_PREFETCH)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "X=0" "BV_RES_PREFETCH=4" "BV_RES_PREFETCH=8" "BV_RES_PREFETCH=16" "BV_L2_PREFETCH=4" "BV_L2_PREFETCH=8" "BV_PAIR_CHAIN=1" "BV_PAIR_CHAIN=1 BV_RES_PREFETCH=8" "BV_PAIR_CHAIN=1 BV_RES_PREFETCH=16" "X=0"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2t_table.csv > gpurun_out/r2t_bench.json 2>gpurun_out/r2t_bench.err
echo "== $v"; grep -E "\+res|pair_chain" gpurun_out/r2t_table.csv | awk -F, '{n[$1]++; s[$1]+=$2} END {for (k in n) printf "%s x%d avg %.4f ms\n", k, n[k], s[k]/n[k]}' | sort
python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('gathered_checksum'))"
done
