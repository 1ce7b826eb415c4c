set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/r2a_launch_table.csv > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 600 gpurun_out/r2a_bench.err
cat gpurun_out/r2a_bench.json | head -c 3000
BV_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2a_bench_nopdl.json 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2a_bench_pdl2.json 2>/dev/null
python -c "
import json
for f in ('r2a_bench','r2a_bench_nopdl','r2a_bench_pdl2'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['e2e']['value']))
"
timeout 300 python tests/latency_b1.py 512 200 > gpurun_out/r2a_latency_b1.jsonl 2> gpurun_out/r2a_latency.err; cat gpurun_out/r2a_latency_b1.jsonl; tail -3 gpurun_out/r2a_latency.err
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
