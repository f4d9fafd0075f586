set -x
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:step_kernel_tile_many -c 1 --launch-skip 1 -o gpurun_out/prof_r2_tile_many python scripts/profile_step.py --envs 8192 --steps 4 --many 64 > gpurun_out/ncu_a.log 2>&1
$NCU -k regex:mlp_fwd2 -c 1 --launch-skip 8 -o gpurun_out/prof_r2_fwd2 python scripts/time_actor.py 262144 --short > gpurun_out/ncu_b.log 2>&1
$NCU -k "regex:mlp_tile|dw_kernel" -c 6 --launch-skip 12 -o gpurun_out/prof_r2_ppo_v2 python scripts/profile_ppo.py --iters 4 > gpurun_out/ncu_c.log 2>&1
$NCU -k regex:step_kernel_tile -c 1 --launch-skip 4 -o gpurun_out/prof_r2_cfg5_tile python scripts/profile_step.py --envs 131072 --drones 16 --physics dyn_dw --steps 8 > gpurun_out/ncu_d.log 2>&1
python bench.py --steps 2 --warmup 3 --mappo-steps 0 --cpu-steps 3 > gpurun_out/r2c_b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --mappo-steps 0 --cpu-steps 3 > gpurun_out/ncu_e.log 2>&1
tail -2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log gpurun_out/ncu_d.log gpurun_out/ncu_e.log
ls -la gpurun_out/prof_r2_*
