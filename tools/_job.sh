python tools/time_c4.py > gpurun_out/r2_t35_c4.txt 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "capt or c4 or mvt or pointcloud or attach or fuzz or adversarial" 2>&1 | tail -5 >> gpurun_out/r2_t35_c4.txt
python tools/time_generic.py >> gpurun_out/r2_t35_c4.txt 2>&1
