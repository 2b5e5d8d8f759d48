timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 -k "learning" 2>&1 | tail -8
