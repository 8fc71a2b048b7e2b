python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t7_tests.log
python bench.py --no-c5 --steps 20 > gpurun_out/r2_t7_bench.json 2> gpurun_out/r2_t7_bench.err
python tools/time_generic.py > gpurun_out/r2_t7_generic.txt 2>&1
