# ncu --set full of the pair-chained layer3 kernel (default since the last session)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_chain -s 4 -c 1 -f -o gpurun_out/r2pc4_pair_chain_l3 python tools/ncu_step.py 512 2 > gpurun_out/r2pc4.log 2>&1
ncu -i gpurun_out/r2pc4_pair_chain_l3.ncu-rep --page raw --csv > gpurun_out/r2pc4_pair_chain_l3_raw.csv 2>/dev/null
ls -la gpurun_out/r2pc4_pair_chain_l3.ncu-rep | awk '{print $5, $9}'
