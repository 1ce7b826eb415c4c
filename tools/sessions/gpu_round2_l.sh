cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_l1_block_gpu.py -q -x 2>&1 | tail -6
bash tools/gpu_ab.sh "X=0" "BV_L1_SH=1" "BV_L1_LAST=1" "BV_L1_SH=1 BV_L1_LAST=1"
