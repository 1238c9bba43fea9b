timeout 300 python -m pytest tests/test_lut_gpu.py tests/test_forms_gpu.py -m gpu -x -q 2>&1 | tail -4
python scripts/bench_lut.py 2>&1 | grep "full_set\|batch64_dev"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed -k regex:"dds_lut|pack_dram" --clock-control none -c 4 --csv --log-file gpurun_out/r01d_launches_dds.csv python scripts/bench_lut.py > gpurun_out/ncu_lut.log 2>&1; tail -1 gpurun_out/ncu_lut.log
