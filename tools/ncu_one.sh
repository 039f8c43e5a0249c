#!/bin/bash
# One ncu --set full capture of one kernel, summarised on the box: tools/ncu_one.sh <tag> <name> <kernel regex> <kernels_once arg> [skip]
tag=$1; name=$2
python tools/kernels_once.py $4 > gpurun_out/${tag}_${name}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$3" -c 1 -s "${5:-1}" -o gpurun_out/${tag}_${name} -f \
    python tools/kernels_once.py $4 > gpurun_out/${tag}_${name}.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_${name}.ncu-rep gpurun_out/${tag}_${name} > /dev/null 2>&1
rm -f gpurun_out/${tag}_${name}.ncu-rep
