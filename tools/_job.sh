VMV_DBG_E2E=1 python bench.py --no-cpu --no-c5 --no-c4 --no-edges > gpurun_out/r2_t46.json 2> gpurun_out/r2_t46.err
