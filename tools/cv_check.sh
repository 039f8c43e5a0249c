timeout 300 python tools/cv_profile.py 2>&1 | head -3
timeout 600 python -m pytest tests -x -q -m gpu -k "cv or c2 or golden or smoke" 2>&1 | tail -3
