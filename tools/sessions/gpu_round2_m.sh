cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in "X=0" "BV_L1_SH=1 BV_L1_LAST=1"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2m_table.csv > gpurun_out/r2m_bench.json 2>/dev/null
echo "== $v"; head -8 gpurun_out/r2m_table.csv
done
