# do the suspend-time hints change the number of shared-memory load wavefronts (mbarrier polls) of l1_block?
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:l1_block -s 3 -c 3 --csv --log-file gpurun_out/r2h3_default.csv python tools/ncu_step.py 512 2 > /dev/null 2>&1
BV_LIB_PATH=$PWD/build/libbiovil_b200_hint5000.so timeout 300 ncu --metrics $M --clock-control none -k regex:l1_block -s 3 -c 3 --csv --log-file gpurun_out/r2h3_hint5000.csv python tools/ncu_step.py 512 2 > /dev/null 2>&1
for f in default hint5000; do echo "== $f"; grep -v "^==" gpurun_out/r2h3_$f.csv | awk -F'","' 'NR>1 {print $5 " | " $(NF-2) " | " $NF}' | sed 's/"//g' | cut -c1-160; done
