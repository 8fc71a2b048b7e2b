python bench.py > gpurun_out/r2_t48_bench.json 2> gpurun_out/r2_t48_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_t48_ref.json 2>> gpurun_out/r2_t48_bench.err
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu.log 2>&1
