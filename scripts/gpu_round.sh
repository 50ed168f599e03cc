#!/bin/bash
# One GPU session: tests, bench (both arms), ncu launch list + full captures.  Outputs under gpurun_out/<tag>/.
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
[ -n "$SKIP_TESTS" ] || python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-1500
tail -5 $OUT/bench.err
python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tee $OUT/bench_reference.json | cut -c1-600
# the two physics families (BASELINE configs[3], configs[4] at one GPU), both arms
for E in Humanoid-v5 LunarLander-v3; do
  python bench.py --env $E --steps 40 --warmup 5 2>> $OUT/bench.err | tee $OUT/bench_$E.json | cut -c1-400
  python bench.py --env $E --impl reference --steps 5 --warmup 3 2>/dev/null | tee $OUT/bench_reference_$E.json | cut -c1-400
done
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 2500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-extras --e2e-steps 20 > $OUT/bench_under_ncu.log 2>&1
# full captures of the top kernels
for T in step big rollout lake lander humanoid; do
  case $T in
    step) K=cartpole_step_kernel; S=80;;
    big) K=cartpole_step_kernel; S=3;;
    rollout) K=cartpole_rollout_kernel; S=1;;
    lake) K=frozenlake; S=8;;
    lander) K=lunarlander_step; S=70;;
    humanoid) K=humanoid_step_warp; S=60;;
  esac
  C=3; case $T in lander|humanoid) C=1;; esac   # gpurun_out/ is capped at 64 MiB: one launch of the two big kernels
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o $OUT/ncu_$T \
      python scripts/ncu_targets.py $T > $OUT/ncu_$T.log 2>&1
done
ls -la $OUT
