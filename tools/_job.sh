python -m pytest tests/test_gpu_parity.py -x -q -k "capt or mvt or c4 or heightfield" 2>&1 | tail -5 > gpurun_out/r2_t13_tests.log
python bench.py --no-c5 --no-edges --steps 10 > gpurun_out/r2_t13_bench.json 2> gpurun_out/r2_t13_bench.err
