cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2fin_pytest.log; tail -3 gpurun_out/r2fin_pytest.log
timeout 600 python bench.py --steps 30 --warmup 3 --profile-out gpurun_out/r2fin_launch_table_events.csv > gpurun_out/r2fin_bench.json 2> gpurun_out/r2fin_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2fin_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks'], round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['frac_burst'], d['gpu_library_baseline']['value'], d['cpu_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2fin_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/r2fin_bench_reference_arm.json
timeout 300 python tests/latency_b1.py 512 200 > gpurun_out/r2fin_latency_b1.jsonl 2>/dev/null; cat gpurun_out/r2fin_latency_b1.jsonl
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/stress_forward.py 3000 2>&1 | tail -1
N=$(timeout 120 python tools/ncu_step.py 512 2 | awk '{print $NF}')
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'conv_gemm|chain_gemm|pair_chain|conv3x3_tap3|l1_block|stem_|head' -s $N -c $N --csv --log-file gpurun_out/r2fin_ncu_launch_list_dram.csv \
  python tools/ncu_step.py 512 2 > gpurun_out/r2fin_ncu1.log 2>&1
tail -1 gpurun_out/r2fin_ncu1.log
