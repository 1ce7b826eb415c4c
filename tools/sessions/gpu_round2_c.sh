# configs[2] at full size on 1/2/4/8 GPUs of ONE box + the 8-GPU weak-scaling bench + the two-devices-in-one-process test
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_model_gpu.py -q -k two_devices 2>&1 | tail -3
timeout 600 python tools/extract_full_scale.py --out gpurun_out/r2k_extract_224k_1gpu.json 2> gpurun_out/r2k_extract_1.err | cut -c1-400
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
     tools/extract_full_scale.py --out gpurun_out/r2k_extract_224k_${n}gpu.json 2> gpurun_out/r2k_extract_$n.err | cut -c1-400
done
for n in 1 8; do
  if [ $n = 1 ]; then timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2k_bench_1gpu.json 2>/dev/null
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $n --steps 30 --warmup 3 > gpurun_out/r2k_bench_${n}gpu.json 2> gpurun_out/r2k_bench_$n.err; fi
done
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.load(open(f'gpurun_out/r2k_extract_224k_{n}gpu.json'))
        print(n, round(d['images_per_s_excl_gather']), round(d['images_per_s_incl_gather']), round(d['gather_ms'],2), round(d['total_ms_max_over_ranks'],1), d['checksums'])
    except Exception as e: print(n, 'ERR', e)
for n in (1,8):
    try:
        d=json.load(open(f'gpurun_out/r2k_bench_{n}gpu.json')); print('bench',n, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))
    except Exception as e: print('bench', n, 'ERR', e)
PY
tail -3 gpurun_out/r2k_extract_8.err
