timeout 300 python tools/cv_profile.py 2>&1 | head -22 | cut -c1-150
timeout 600 python -m pytest tests -x -q -m gpu -k "cv or c2 or bulk_fit or smoke or golden" 2>&1 | tail -3
