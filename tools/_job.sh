python tools/time_kernels.py > gpurun_out/r2_t36_k.txt 2>&1
python tools/time_c4.py >> gpurun_out/r2_t36_k.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -4 >> gpurun_out/r2_t36_k.txt
