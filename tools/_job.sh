python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "adversarial" 2>&1 | tail -30 > gpurun_out/r2_t30_tests.log
