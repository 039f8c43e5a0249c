timeout 600 python -m pytest tests -x -q -m gpu -k "tensor_core or snv or nonfinite or bulk_shape" 2>&1 | tail -4
timeout 300 python tools/ab_score.py --libs ab/lib_tc1_f16.so ab/lib_res_f16.so ab/lib_tc1_bf16.so ab/lib_res_bf16.so --steps 20 --rounds 2 2>&1 | tail -9
