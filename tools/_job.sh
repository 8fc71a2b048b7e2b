for c in 15 16 17 18 19 20; do VMV_CHUNK_LOG2=$c python tools/time_e2e.py; done > gpurun_out/r2_e2e.txt 2>&1
for c in 17 18; do VMV_CHUNK_LOG2=$c python tools/time_e2e.py; done >> gpurun_out/r2_e2e.txt 2>&1
