timeout 600 python -m pytest tests -x -q -m gpu -k "snv or tensor_core or nonfinite or bulk_shape or host_buffer" 2>&1 | tail -3
timeout 500 python tools/ab_score.py --libs ab/lib_t3d_f16.so ab/lib_flag_f16.so --steps 20 --rounds 4 2>&1 | tail -8
