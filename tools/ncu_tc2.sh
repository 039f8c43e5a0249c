timeout 300 python -m pytest tests -x -q -m gpu -k "tensor_core or nonfinite" 2>&1 | tail -2
for lead in 12 13 14; do echo "lead $lead"; OCM_TC2_LEAD=$lead timeout 200 python tools/ab_score.py --libs ab/lib_tc1.so ab/lib_tc2.so --steps 20 --rounds 2 2>&1 | tail -4; done
