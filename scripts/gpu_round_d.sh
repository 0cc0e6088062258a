set -x
./build/fp64_pipes_probe > gpurun_out/fp64_pipes_probe.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/quick_time.py 32 148,592 2>&1 | tail -3
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_d.json 2> gpurun_out/bench_r1_d.err; tail -c 600 gpurun_out/bench_r1_d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --chains 1184 > gpurun_out/ncu_d1.log 2>&1
ncu --set full --clock-control none --import-source on -k k_chain -c 1 -o gpurun_out/prof_r1_d python scripts/quick_time_one.py > gpurun_out/ncu_d2.log 2>&1
python scripts/phase_profile.py 32 148 2>&1 | tail -25
