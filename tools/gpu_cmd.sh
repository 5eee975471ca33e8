python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -2 gpurun_out/bench_r1c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1c.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac']); print(d['e2e'])"
