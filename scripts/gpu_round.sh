set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3; python scripts/quick_time_big.py 2>&1 | tail -6 | tee gpurun_out/quick_time_big_r1_m.txt
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r1_m.json 2> gpurun_out/bench_r1_m.err; tail -c 900 gpurun_out/bench_r1_m.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_m_ref.json 2>> gpurun_out/bench_r1_m.err; tail -c 400 gpurun_out/bench_r1_m_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --chains 1184 > gpurun_out/ncu_e1.log 2>&1
ncu --set full --clock-control none --import-source on -k k_chain -c 1 -o gpurun_out/prof_r1_m python scripts/quick_time_one.py > gpurun_out/ncu_e2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
python scripts/stencil_bench.py > gpurun_out/stencil_bench_r1_m.txt 2>&1; tail -3 gpurun_out/stencil_bench_r1_m.txt
