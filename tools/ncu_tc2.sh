timeout 900 python -m pytest tests -x -q -m gpu -k "tensor_core or cluster" 2>&1 | tail -5
echo "p=2048 tcgen05 cluster-of-four"; timeout 200 python tools/ab_score.py --libs ab/lib_tc1_bf16.so --steps 10 --rounds 1 --p 2048 --rows 5000000 2>&1 | tail -1
echo "p=2048 row-team"; OCM_DISABLE_TC=1 timeout 200 python tools/step_times.py --steps 10 --rows 5000000 --p 2048 2>&1 | tail -1
echo "p=2048 dispatcher"; timeout 200 python tools/step_times.py --steps 10 --rows 5000000 --p 2048 2>&1 | tail -1
