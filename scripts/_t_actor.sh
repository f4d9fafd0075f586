timeout 600 python -m pytest tests/test_gpu_actor.py tests/test_gpu_normalization.py tests/test_gpu_mappo.py -x -q 2>&1 | tail -15 > gpurun_out/r2u_actor_tests.log
timeout 60 python scripts/time_actor.py > gpurun_out/r2u_time_actor.log 2>&1
timeout 60 python scripts/trace_actor.py > gpurun_out/r2u_trace_actor.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:actor_tmem -c 1 --launch-skip 3 -o gpurun_out/prof_r2_actor_tmem python scripts/profile_actor.py > gpurun_out/ncu_actor_tmem.log 2>&1
