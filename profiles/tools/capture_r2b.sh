#!/bin/bash
# Re-capture after the last kernel changes of round 2 (rollout owner path / layer-1 GEMM, release-acquire grid barrier):
# launch lists of the default bench command at C2 / C3 and --set full captures of the two hot kernels.
set -x
F="ncu --set full --import-source on --clock-control none -c 1 -f"
B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra-configs"
$F -k regex:ppo_grad_tc_kernel -s 12 -o gpurun_out/r2b_grad_tc_c2 $B1 > gpurun_out/r2b_ncu_a.log 2>&1
$F -k regex:ppo_grad_tc_kernel -s 48 -o gpurun_out/r2b_grad_tc_c3 $B1 --envs-per-gpu 65536 > gpurun_out/r2b_ncu_b.log 2>&1
$F -k regex:rollout_tc_kernel -s 3 -o gpurun_out/r2b_rollout_tc_c2 $B1 > gpurun_out/r2b_ncu_c.log 2>&1
$F -k regex:rollout_tc_kernel -s 3 -o gpurun_out/r2b_rollout_tc_c3 $B1 --envs-per-gpu 65536 > gpurun_out/r2b_ncu_d.log 2>&1
L="ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs"
$L --log-file gpurun_out/r2b_launches_c2.csv $B > gpurun_out/r2b_launches_c2.log 2>&1
$L --log-file gpurun_out/r2b_launches_c3.csv $B --envs-per-gpu 65536 > gpurun_out/r2b_launches_c3.log 2>&1
ls -la gpurun_out/r2b_*
