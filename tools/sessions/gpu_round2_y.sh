# ncu evidence for the FINAL round-2 binary: launch list with DRAM bytes of one forward + --set full captures of the kernels
# that changed since r2i (the three layer1 block forms) and of the dominant tensor-bound kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$(timeout 120 python tools/ncu_step.py 512 2 | awk '{print $NF}') || exit 1
echo "launches per forward: $N"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:'conv_gemm|chain_gemm|pair_chain|conv3x3_tap3|l1_block|stem_|head' -s $N -c $N --csv --log-file gpurun_out/r2y_ncu_launch_list_dram.csv \
  python tools/ncu_step.py 512 2 > gpurun_out/r2y_ncu1.log 2>&1
tail -2 gpurun_out/r2y_ncu1.log
cap() {  # name, kernel regex, skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/r2y_$1 python tools/ncu_step.py 512 2 > gpurun_out/r2y_$1.log 2>&1
  ncu -i gpurun_out/r2y_$1.ncu-rep --page raw --csv > gpurun_out/r2y_$1_raw.csv 2>/dev/null
  ls -la gpurun_out/r2y_$1.ncu-rep | awk '{print $5, $9}'
}
cap l1_block_ds l1_block 3
cap l1_block_res l1_block 4
cap l1_block_last l1_block 5
cap pair_deep_l3_3x3 conv_gemm_kernel 48
