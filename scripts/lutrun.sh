timeout 300 python -m pytest tests/test_lut_gpu.py tests/test_forms_gpu.py -m gpu -x -q 2>&1 | tail -4
python scripts/bench_lut.py 2>&1 | tail -7
