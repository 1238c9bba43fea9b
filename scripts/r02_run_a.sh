# round 2, call A: GPU tests (incl. the new bench-config parity test), then an ncu capture of the current K4 with source
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -m pytest tests/test_bench_config_parity_gpu.py -x -q -s 2>&1 | grep -E "bench-config|passed|failed|Error|assert" | head -40
python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02a_bench_short.json 2> gpurun_out/r02a_bench_short.err &&
ncu --set full --clock-control none --import-source on -k regex:channelize_kernel -s 4 -c 1 -o gpurun_out/r02a_k4 \
    python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02a_ncu.log 2>&1
echo ncu rc=$?
tail -c 600 gpurun_out/r02a_bench_short.json
