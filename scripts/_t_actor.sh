timeout 300 python -m pytest tests/test_gpu_actor.py tests/test_gpu_normalization.py -x -q 2>&1 | tail -5 > gpurun_out/r3f_actor_tests.log
timeout 60 python scripts/time_actor.py > gpurun_out/r3f_time_actor.log 2>&1
timeout 60 python scripts/trace_actor.py > gpurun_out/r3f_trace_actor.log 2>&1
