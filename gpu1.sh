timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | cut -c1-500
