python bench.py --no-cpu --no-c5 --no-c4 --no-edges > gpurun_out/r2_t37_bench.json 2> gpurun_out/r2_t37_bench.err
python bench.py --no-cpu --no-c5 --no-c4 --no-edges >> gpurun_out/r2_t37_bench.json 2>> gpurun_out/r2_t37_bench.err
