python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r01d_bench_1gpu.json 2> gpurun_out/r01d_bench_1gpu.err; echo rc=$?; tail -c 300 gpurun_out/r01d_bench_1gpu.err
