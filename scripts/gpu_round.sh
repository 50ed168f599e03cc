#!/bin/bash
# One GPU session of profile evidence: ncu launch list of the bench command + full captures of the top kernels.
# Outputs under gpurun_out/<tag>/; scripts/make_profiles.py <tag> <round> turns them into profiles/.
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 2500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-extras --no-humanoid --e2e-steps 20 > $OUT/bench_under_ncu.log 2>&1
EXTRA=sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
for T in step big lake lander humanoid; do
  case $T in
    step) K=cartpole_step_kernel; S=80;;
    big) K=cartpole_step_kernel; S=3;;
    rollout) K=cartpole_rollout_kernel; S=1;;
    lake) K=frozenlake; S=8;;
    lander) K=lunarlander_step; S=70;;
    humanoid) K=humanoid_step_warp; S=60;;
  esac
  C=3; case $T in lander|humanoid) C=1;; esac   # gpurun_out/ is capped at 64 MiB: one launch of the two big kernels
  ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o $OUT/ncu_$T \
      python scripts/ncu_targets.py $T > $OUT/ncu_$T.log 2>&1
done
ls -la $OUT
