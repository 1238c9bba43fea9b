# round-2 end state on ONE GPU: tests, smoke, bench line, reference arm, launch list (K4 itself is unchanged since r02f: its ncu capture stands)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/r02i_bench_1gpu.json 2> gpurun_out/r02i_bench_1gpu.err; echo rc=$?; tail -c 300 gpurun_out/r02i_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_reference_arm.json 2> gpurun_out/r02i_bench_reference_arm.err; echo rc=$?
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_short.json 2> gpurun_out/r02i_short.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches_bench_steps3.csv \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02i_ncu_launches.log 2>&1
echo launches rc=$?
