python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_h8v2.json 2> gpurun_out/bench_h8v2.err; tail -3 gpurun_out/bench_h8v2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_h8v2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
