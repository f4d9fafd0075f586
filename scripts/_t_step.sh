timeout 120 python scripts/time_step_shape.py > gpurun_out/r2z_time_cfg5b.log 2>&1
timeout 120 python scripts/time_step_shape.py --envs 16384 >> gpurun_out/r2z_time_cfg5b.log 2>&1
timeout 120 python scripts/time_step_shape.py --envs 262144 --drones 4 --physics dyn_dw >> gpurun_out/r2z_time_cfg5b.log 2>&1
timeout 120 python scripts/time_step_shape.py --envs 32768 --drones 32 --physics dyn_dw >> gpurun_out/r2z_time_cfg5b.log 2>&1
