#!/bin/bash
# tools/ab_run.sh NAME...  : headline bench (device-timed leg only) once per ab/lib_NAME.so, two rounds, interleaved
cd "$(dirname "$0")/.."
for round in 1 2; do
  for v in "$@"; do
    OCM_SIMCA_LIB=$PWD/ab/lib_$v.so timeout 200 python bench.py --steps 5 --warmup 3 --no-cv --no-e2e --no-cpu 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
