timeout 600 python -m pytest tests/test_lut_gpu.py tests/test_forms_gpu.py -m gpu -x -q -s 2>&1 | grep -E "dds lut|passed|failed|Error|assert" | tail -8
python scripts/bench_lut.py 2>&1 | grep "full_set"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 24 --csv --log-file gpurun_out/r02_launches_lut_batch64.csv python scripts/prof_lut.py 64 > gpurun_out/ncu_lut.log 2>&1; tail -1 gpurun_out/ncu_lut.log
grep dds_lut gpurun_out/r02_launches_lut_batch64.csv | grep duration | awk -F'","' '{print $NF}'
