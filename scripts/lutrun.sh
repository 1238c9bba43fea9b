timeout 600 python -m pytest tests/test_lut_gpu.py tests/test_forms_gpu.py -m gpu -x -q 2>&1 | tail -2
python scripts/bench_lut.py 2>&1 | grep "full_set\|batch64_dev\|batch1_dev"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 24 --csv --log-file gpurun_out/r02_launches_lut_batch64.csv python scripts/prof_lut.py 64 > gpurun_out/ncu_lut.log 2>&1; tail -1 gpurun_out/ncu_lut.log
grep -E "comb_prep" gpurun_out/r02_launches_lut_batch64.csv | grep duration | awk -F'","' '{print $NF}'
