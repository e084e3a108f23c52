#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun): launch lists of the default bench command at C2 / C3 and one
# `--set full` capture per hot kernel.  Outputs land in gpurun_out/ (scratch); profiles/summarize_ncu.py turns them into the
# committed summaries profiles/r2_*.txt.
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs"
L="ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv"
$L --log-file gpurun_out/r2_launches_c2.csv $B > gpurun_out/r2_launches_c2.log 2>&1
$L --log-file gpurun_out/r2_launches_c3.csv $B --envs-per-gpu 65536 > gpurun_out/r2_launches_c3.log 2>&1
F="ncu --set full --import-source on --clock-control none -c 1 -f"
B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-configs"
$F -k regex:ppo_grad_tc_kernel -s 12 -o gpurun_out/r2_grad_tc_c2 $B1 > gpurun_out/r2_ncu_a.log 2>&1
$F -k regex:ppo_grad_tc_kernel -s 48 -o gpurun_out/r2_grad_tc_c3 $B1 --envs-per-gpu 65536 > gpurun_out/r2_ncu_b.log 2>&1
$F -k regex:rollout_tc_kernel -s 3 -o gpurun_out/r2_rollout_tc_c2 $B1 > gpurun_out/r2_ncu_c.log 2>&1
$F -k regex:rollout_tc_kernel -s 3 -o gpurun_out/r2_rollout_tc_c3 $B1 --envs-per-gpu 65536 > gpurun_out/r2_ncu_d.log 2>&1
$F -k regex:gae_kernel -s 3 -o gpurun_out/r2_gae_c3 $B1 --envs-per-gpu 65536 > gpurun_out/r2_ncu_e.log 2>&1
$F -k regex:mlp256_kernel -s 2 -o gpurun_out/r2_mlp256 python profiles/tools/h256_probe.py 524288 > gpurun_out/r2_ncu_f.log 2>&1
$F -k regex:dw2_gemm256 -s 2 -o gpurun_out/r2_dw2_gemm256 python profiles/tools/h256_probe.py 524288 > gpurun_out/r2_ncu_g.log 2>&1
$F -k regex:env_step_kernel -s 5 -o gpurun_out/r2_env_step python profiles/tools/env_step_probe.py > gpurun_out/r2_ncu_h.log 2>&1
$F -k regex:replay_gather_kernel -s 2 -o gpurun_out/r2_replay_gather python profiles/tools/replay_probe.py > gpurun_out/r2_ncu_i.log 2>&1
python profiles/tools/env_step_probe.py > gpurun_out/r2_env_step_probe.log 2>&1
python profiles/tools/replay_probe.py > gpurun_out/r2_replay_probe.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep
