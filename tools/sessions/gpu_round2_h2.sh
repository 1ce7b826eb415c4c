# A/B of library builds: mbarrier waits without / with a suspend-time hint (500 ns, 5000 ns)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do
for v in "X=0" "BV_LIB_PATH=$PWD/build/libbiovil_b200_hint500.so" "BV_LIB_PATH=$PWD/build/libbiovil_b200_hint5000.so"; do
env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-baseline --profile-out gpurun_out/r2h2_table.csv > gpurun_out/r2h2_bench.json 2>gpurun_out/r2h2_bench.err
echo "== ${v##*/}"; grep -E "l1_block|stem|chain" gpurun_out/r2h2_table.csv | cut -d, -f1,2 | tr '\n' ' '; echo
python -c "
import json; d=json.load(open('gpurun_out/r2h2_bench.json')); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['config'].get('gathered_checksum'))"
done
done
