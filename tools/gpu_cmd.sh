python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/time_cfg5.py 1200
GL_DISABLE_TPC=1 python tools/time_cfg5.py 1200
