python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2_t49_n8.json 2> gpurun_out/r2_t49_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 40 --warmup 5 --no-cpu > gpurun_out/r2_t49_n4.json 2>> gpurun_out/r2_t49_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu > gpurun_out/r2_t49_n2.json 2>> gpurun_out/r2_t49_n8.err
python bench.py --no-cpu --no-c4 > gpurun_out/r2_t49_n1.json 2>> gpurun_out/r2_t49_n8.err
