# final binary on 2 GPUs: two-devices-in-one-process test, weak-scaling bench at N=2, configs[2] on 2 GPUs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -q -k two_devices 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r2zz_bench_2gpu.json 2> gpurun_out/r2zz_bench_2.err
python -c "
import json; d=json.load(open('gpurun_out/r2zz_bench_2gpu.json')); print('bench 2', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['roofline']['frac'], d['clocks']['sm_mhz'], d['config'].get('collective'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 tools/extract_full_scale.py --out gpurun_out/r2zz_extract_224k_2gpu.json 2> gpurun_out/r2zz_extract_2.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r2zz_extract_224k_2gpu.json')); print('extract 2', round(d['images_per_s_incl_gather']), d['checksums'])"
