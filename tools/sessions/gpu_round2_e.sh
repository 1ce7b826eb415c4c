cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_chain_gpu.py -q -x -k "pair_chain" 2>&1 | tail -15
bash tools/gpu_ab.sh "BV_PAIR_CHAIN=0" "BV_PAIR_CHAIN=1" "BV_PAIR_CHAIN=3" "BV_PAIR_CHAIN=7"
