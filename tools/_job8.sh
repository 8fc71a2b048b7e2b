python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 40 --warmup 5 > gpurun_out/r2_t34_n8.json 2> gpurun_out/r2_t34_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 40 --warmup 5 --gather none --no-cpu --no-edges --no-c5 > gpurun_out/r2_t34_n8_none.json 2>> gpurun_out/r2_t34_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu > gpurun_out/r2_t34_n2.json 2>> gpurun_out/r2_t34_n8.err
python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r2_t34_comm.log
