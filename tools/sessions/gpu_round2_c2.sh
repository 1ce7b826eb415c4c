# ncu --set full of the layer2 chain kernel and the layer3 conv3 + identity kernel (final binary), for the LSU-wavefront analysis
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/r2c2_$1 python tools/ncu_step.py 512 2 > gpurun_out/r2c2_$1.log 2>&1
  ncu -i gpurun_out/r2c2_$1.ncu-rep --page raw --csv > gpurun_out/r2c2_$1_raw.csv 2>/dev/null
  ls -la gpurun_out/r2c2_$1.ncu-rep | awk '{print $5, $9}'
}
cap chain_l2 chain_gemm 2
cap conv3_res_l3 conv_gemm_kernel 49
