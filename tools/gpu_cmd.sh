python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/time_hex8.py 200 5 fused 2>&1 | tail -3
