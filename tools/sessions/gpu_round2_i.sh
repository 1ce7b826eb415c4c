# ncu evidence for the round-2 binary: launch list with DRAM bytes of one forward + --set full captures of three kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 120 python tools/ncu_step.py 512 2 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'conv_gemm|chain_gemm|pair_chain|conv3x3_tap3|l1_block|stem_|head' -s 45 -c 45 --csv --log-file gpurun_out/r2i_ncu_launch_list_dram.csv \
  python tools/ncu_step.py 512 2 > gpurun_out/r2i_ncu1.log 2>&1
tail -2 gpurun_out/r2i_ncu1.log
cap() {  # name, kernel regex, skip
  timeout 600 ncu --set full --clock-control none --import-source off -k regex:$2 -s $3 -c 1 -f -o gpurun_out/r2i_$1 python tools/ncu_step.py 512 2 > gpurun_out/r2i_$1.log 2>&1
  ncu -i gpurun_out/r2i_$1.ncu-rep --page raw --csv > gpurun_out/r2i_$1_raw.csv 2>/dev/null
  ls -la gpurun_out/r2i_$1.ncu-rep | awk '{print $5, $9}'
}
cap pair_deep_l3_3x3 conv_gemm_kernel 49
cap pair_res_l4_conv3 conv_gemm_kernel 68
cap l1_block_ds l1_block 2
cap l1_block_res l1_block 3
