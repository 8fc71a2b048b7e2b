VMV_CAPT_TIMING=1 python tools/_dbg.py > gpurun_out/r2_t50.txt 2>&1
