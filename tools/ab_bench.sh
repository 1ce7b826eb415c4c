#!/bin/bash
# Same-box A/B of two settings of the library's experiment switches (box-to-box spread is +-3 %, so only back-to-back runs
# on ONE box decide a tuning question).  Usage, inside one gpurun call:
#   bash tools/ab_bench.sh "BV_NO_TR=1" "" [repeats] [steps]
# prints images/s, ms per step and the median SM clock of every run, alternating A and B.
A="$1"; B="$2"; REP="${3:-2}"; STEPS="${4:-30}"
run() {
  env $1 timeout 250 python bench.py --steps "$STEPS" --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read())
print('%-28s %7.0f img/s  %.3f ms/step  %.0f MHz' % (sys.argv[1] or '(default)', d['value'], d['ms_per_step'], d['clocks']['sm_mhz']))" "$1"
}
for i in $(seq 1 "$REP"); do run "$A"; run "$B"; done
