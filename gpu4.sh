python -m pytest tests -m gpu -x -q -s 2>&1 | tail -25
