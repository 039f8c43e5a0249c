#!/bin/bash
# Cluster scoring kernel (score_tc2.cu): DRAM bytes and time of one 2 M x 1024 launch as a function of how many stages
# the HBM stream may run ahead of pass 2 (OCM_TC2_LEAD) and of the L2 eviction hints of the two TMA rings (OCM_TC2_POL:
# bits 0-1 pass 1, bits 2-3 pass 2; 0 evict_last / evict_first as shipped, 1 = evict_normal on pass 1, ...).  gpurun, 1 GPU.
export OCM_SCORE_TC2=1
for pol in 0 1; do for lead in 12 14 16 20 100000; do
  OCM_TC2_POL=$pol OCM_TC2_LEAD=$lead ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:score_tc2_kernel -c 1 -s 2 python tools/kernels_once.py score 2>&1 | grep -E "dram__bytes_read|gpu__time" | awk -v p=$pol -v l=$lead '{printf "pol %s lead %s %s %s %s | ", p, l, $1, $2, $3} END {print ""}'
done; done
