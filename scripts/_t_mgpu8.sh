timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/peer_allreduce_check.py > gpurun_out/r2y_peer8.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2y_bench_8gpu.json 2> gpurun_out/r2y_bench_8gpu.err
tail -3 gpurun_out/r2y_peer8.log
