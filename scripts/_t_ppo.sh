timeout 900 python -m pytest tests/test_gpu_ppo_kernels.py tests/test_gpu_mappo.py -x -q 2>&1 | tail -4 > gpurun_out/r3h_ppo_tests.log
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-steps 3 --e2e-steps 2 --vecenv-steps 0 --small-envs 0 --open-loop-reps 0 --mappo-steps 3 > gpurun_out/r3h_bench_mappo.json 2> gpurun_out/r3h_bench_mappo.err
