python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; tail -2 gpurun_out/bench_n1b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1b.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['parity_first_cell_rel'], d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:60], d['checks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; tail -2 gpurun_out/bench_n2b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2b.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['e2e']['workload'][:40])"
