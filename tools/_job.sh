python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_t29_tests.log
python tools/time_kernels.py > gpurun_out/r2_t29_k.txt 2>&1
python tools/time_c4.py >> gpurun_out/r2_t29_k.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_validate_configs_v4 -s 2 -c 1 -o gpurun_out/r2_c4_fetch_v5 -f python tools/prof_configs.py c4 18 fetch > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_validate_configs_v4 -s 2 -c 1 -o gpurun_out/r2_c4_ur5_v5 -f python tools/prof_configs.py c4 18 ur5 >> gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_validate_edges_v4 -s 2 -c 1 -o gpurun_out/r2_edges_v3 -f python tools/prof_configs.py edges 18 >> gpurun_out/ncu.log 2>&1
