timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/peer_allreduce_check.py > gpurun_out/r3d_peer2.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mappo_multi_gpu.py > gpurun_out/r3d_mappo2.log 2>&1
timeout 200 python -m pytest tests/test_gpu_peer_allreduce.py -q > gpurun_out/r3d_peer_test.log 2>&1
tail -2 gpurun_out/r3d_peer2.log; tail -2 gpurun_out/r3d_mappo2.log; tail -2 gpurun_out/r3d_peer_test.log
