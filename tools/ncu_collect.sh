#!/bin/bash
# One ncu --set full capture per hot kernel (run under gpurun, 1 GPU); reports land in gpurun_out/.
# usage: tools/ncu_collect.sh <tag>
tag=${1:-r02}
set -x
python tools/kernels_once.py all > gpurun_out/${tag}_kernels_plain.log 2>&1 || exit 1
cap() {  # name, regex, script arg
  ncu --set full --clock-control none --import-source on -k regex:"$2" -c 1 -s "${4:-1}" -o gpurun_out/${tag}_$1 -f \
      python tools/kernels_once.py $3 > gpurun_out/${tag}_$1.log 2>&1
  # summarise on the box (gpurun brings back at most 64 MiB): keep the full report of the two headline kernels only
  python tools/ncu_summary.py gpurun_out/${tag}_$1.ncu-rep gpurun_out/${tag}_$1 > /dev/null 2>&1
  case "$1" in score_tc_f16|gram_tc) ;; *) rm -f gpurun_out/${tag}_$1.ncu-rep ;; esac
}
cap score_tc_f16 "score_tc_kernel" score 2
cap score_tc_bf16 "score_tc_kernel" score 3
cap score_fixup "score_fixup_kernel" score 1
cap score_tc2_c4 "score_tc2_kernel" score 2
cap gram_tc "gram_tc2?_kernel" gram 2
cap gram_quantize "gram_quantize_kernel" gram 1
cap colstats_minmax "colstats_minmax_kernel" gram 0
cap score_multiA "score_simt_kernel" all 1
cap preprocess "preprocess_kernel" all 1
cap savgol "savgol_.*kernel" all 1
cap latent "latent_score_kernel" all 1
cap jacobi "jacobi_eigh_kernel" all 0
cap symm_mul "symm_mul_kernel" score 3
# launch list of the headline bench (share of each kernel in the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${tag}_launches.log 2>&1
