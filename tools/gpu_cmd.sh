python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_full.json 2> gpurun_out/bench_r1_full.err; tail -2 gpurun_out/bench_r1_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; tail -2 gpurun_out/bench_r1_ref.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hex8_bdb -s 3 -c 1 -o gpurun_out/prof_h8_r1 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
