python -m pytest tests/test_gpu_parity.py tests/test_host_logic.py tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_t41.txt
VMV_CAPT_TIMING=1 python tools/time_capt_build.py >> gpurun_out/r2_t41.txt 2>&1
