cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 1 2; do
BV_PAIR_CHAIN=1 BV_PC_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2g_table_v$v.csv > gpurun_out/r2g_bench_v$v.json 2>/dev/null
grep "pair_chain" gpurun_out/r2g_table_v$v.csv | head -2
done
