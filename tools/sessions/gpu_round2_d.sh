cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_l1_block_gpu.py -q -x 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2d_pytest.log; tail -4 gpurun_out/r2d_pytest.log
timeout 400 python tools/stress_forward.py 4000 2>&1 | tail -2
bash tools/gpu_ab.sh "X=0" "BV_NO_L1_DS=1"
