timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_mappo_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 3 --e2e-steps 2 --vecenv-steps 0 --small-envs 0 --open-loop-reps 0 --mappo-steps 1 > gpurun_out/ncu_mappo.log 2>&1
tail -2 gpurun_out/ncu_mappo.log | cut -c1-300
wc -l gpurun_out/r2_mappo_launches.csv
