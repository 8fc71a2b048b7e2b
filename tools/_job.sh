for v in e320 e640 c768 c576; do VMV_LIB=variants/lib_$v.so timeout 200 python tools/time_kernels.py >> gpurun_out/r2_t32_k.txt 2>&1; done
python tools/time_kernels.py >> gpurun_out/r2_t32_k.txt 2>&1
