VMV_CAPT_TIMING=1 python tools/time_capt_build.py 2>&1 | tail -12 > gpurun_out/r2_t52.txt
