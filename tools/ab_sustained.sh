#!/bin/bash
# tools/ab_sustained.sh NAME... : 40 back-to-back launches of the headline step per ab/lib_NAME.so (power-capped regime)
cd "$(dirname "$0")/.."
for round in 1 2; do
  for v in "$@"; do
    echo -n "$v "; OCM_SIMCA_LIB=$PWD/ab/lib_$v.so timeout 200 python tools/step_times.py --steps 40 2>/dev/null
  done
done
