python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2_t15_tests.log
