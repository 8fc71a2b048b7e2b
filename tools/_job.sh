python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t10_tests.log
python bench.py --no-c5 --no-edges --steps 10 > gpurun_out/r2_t10_bench.json 2> gpurun_out/r2_t10_bench.err
