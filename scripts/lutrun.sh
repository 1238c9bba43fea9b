timeout 300 python -m pytest tests/test_lut_gpu.py tests/test_forms_gpu.py -m gpu -x -q 2>&1 | tail -4
python scripts/bench_lut.py 2>&1 | grep "full_set\|batch64_dev"
ncu --metrics gpu__time_duration.sum -k regex:"dds_|pack_dram" --clock-control none -c 12 --csv --log-file gpurun_out/r01d_launches_dds.csv python scripts/prof_lut.py 1 > gpurun_out/ncu_lut.log 2>&1; tail -1 gpurun_out/ncu_lut.log
