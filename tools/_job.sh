ncu --set full --clock-control none --import-source on -k regex:k_validate_configs -c 1 -s 2 -o gpurun_out/r2_c4_fetch_v1 python tools/prof_configs.py c4 18 fetch > gpurun_out/r2_ncu2.log 2>&1
