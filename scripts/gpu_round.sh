# one measurement round on the GPU box (tag = $1): tests, bench (both arms), ncu launch list, ncu --set full capture of k_chain, smoke
T=${1:-r2f}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; tail -c 600 gpurun_out/bench_$T.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_ref.json 2>> gpurun_out/bench_$T.err; tail -c 300 gpurun_out/bench_${T}_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --chains 1184 > gpurun_out/ncu_e1.log 2>&1
ncu --set full --clock-control none --import-source on -k k_chain -c 1 -f -o gpurun_out/prof_$T python scripts/quick_time_one.py > gpurun_out/ncu_e2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
python scripts/stencil_bench.py > gpurun_out/stencil_bench_$T.txt 2>&1; tail -3 gpurun_out/stencil_bench_$T.txt
