set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01d_bench_reference_arm.json 2> gpurun_out/r01d_bench_reference_arm.err; tail -c 600 gpurun_out/r01d_bench_reference_arm.json
python bench.py > gpurun_out/r01d_bench_1gpu.json 2> gpurun_out/r01d_bench_1gpu.err; echo rc=$?; tail -c 300 gpurun_out/r01d_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01d_launches_bench_steps3.csv python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -c 200 gpurun_out/ncu_bench.log
