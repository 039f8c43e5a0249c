timeout 600 python -m pytest tests -x -q -m gpu -k "snv or tensor_core" 2>&1 | tail -12 | cut -c1-200
