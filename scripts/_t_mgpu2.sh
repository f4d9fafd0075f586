timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mappo_multi_gpu.py > gpurun_out/r2x_mappo2.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2x_bench_2gpu.json 2> gpurun_out/r2x_bench_2gpu.err
tail -4 gpurun_out/r2x_mappo2.log
