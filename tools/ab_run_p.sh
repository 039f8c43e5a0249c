#!/bin/bash
# tools/ab_run_p.sh P NAME... : like ab_run.sh at another wavelength count
cd "$(dirname "$0")/.."
p=$1; shift
for round in 1 2; do
  for v in "$@"; do
    OCM_SIMCA_LIB=$PWD/ab/lib_$v.so timeout 200 python bench.py --p $p --steps 5 --warmup 3 --no-cv --no-e2e --no-cpu 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'p=$p', round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['roofline'].get('kernel', d.get('kernel_name','')), d['clocks']['reasons'])"
  done
done
