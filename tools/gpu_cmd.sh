python tools/bench_configs.py --scale 0.2 --only 4 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mass_kernel -s 1 -c 1 -o gpurun_out/prof_mass -f python tools/bench_configs.py --scale 0.2 --only 4 > gpurun_out/ncu_mass.log 2>&1
tail -1 gpurun_out/ncu_mass.log
python tools/bench_configs.py --scale 1.0 --only 1 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:small2d_kernel -s 1 -c 1 -o gpurun_out/prof_small2d -f python tools/bench_configs.py --scale 1.0 --only 1 > gpurun_out/ncu_small2d.log 2>&1
tail -1 gpurun_out/ncu_small2d.log
