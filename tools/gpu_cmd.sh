python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; tail -2 gpurun_out/bench_r1b.err; cat gpurun_out/bench_r1b.json
for pat in 0 1; do echo pat $pat; GL_HEX8_PATTERN=$pat python tools/time_hex8.py 200 10 2>&1 | tail -1; done
