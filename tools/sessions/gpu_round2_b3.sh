# conv_gemm / chain_gemm biases by value (constant bank): full GPU suite, then A/B against the previous build (l1_block only)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for rep in 1 2 3; do
for v in "BV_LIB_PATH=$PWD/build/libbiovil_b200_prev.so" "X=0"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2b3_table_$rep.csv > gpurun_out/r2b3_bench.json 2>gpurun_out/r2b3_bench.err
echo "== ${v##*/}"; grep -E "\+res" gpurun_out/r2b3_table_$rep.csv | awk -F, '{n[$1]++; s[$1]+=$2} END {for (k in n) printf "%s x%d %.4f | ", substr(k,1,40), n[k], s[k]/n[k]}'; echo
python -c "
import json; d=json.load(open('gpurun_out/r2b3_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('gathered_checksum'))"
[ "$v" = "X=0" ] || cp gpurun_out/r2b3_table_$rep.csv gpurun_out/r2b3_table_prev_$rep.csv
done
done
