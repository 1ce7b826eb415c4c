cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in 0 7; do
BV_PAIR_CHAIN=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2f_table_pc$v.csv > gpurun_out/r2f_bench_pc$v.json 2>/dev/null
done
python - <<'PY'
import csv
def load(f): return [(r['name'],float(r['ms'])) for r in csv.DictReader(open(f))]
a=load('gpurun_out/r2f_table_pc0.csv'); b=load('gpurun_out/r2f_table_pc7.csv')
for n,m in a: print(f"{n[:60]:60s} {m:.4f}")
print('---- pair chain on')
for n,m in b: print(f"{n[:60]:60s} {m:.4f}")
print(sum(m for _,m in a), sum(m for _,m in b))
PY
