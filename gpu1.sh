timeout 900 python -m pytest tests -m gpu -q -x -k "full_size" 2>&1 | tail -15 | cut -c1-300
