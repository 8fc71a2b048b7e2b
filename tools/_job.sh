python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "capt or c4 or mvt or pointcloud or attach or fuzz" 2>&1 | tail -12 > gpurun_out/r2_t21_tests.log
python tools/time_c4.py > gpurun_out/r2_t21_c4.txt 2>&1
