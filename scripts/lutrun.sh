timeout 300 python -m pytest tests/test_lut_gpu.py -m gpu -x -q 2>&1 | tail -2
for f in 2 4; do echo "rows per CTA $f"; MKID_LUT_R16F=$f python scripts/bench_lut.py 2>&1 | grep "batch1_device\|batch64\|batch512"; done
ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 18 --csv --log-file gpurun_out/r01d_launches_lut.csv python scripts/prof_lut.py 64 > gpurun_out/ncu_lut.log 2>&1; tail -1 gpurun_out/ncu_lut.log
