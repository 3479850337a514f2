#!/bin/bash
# run_watertank_changing.sh of the reference (lines 1-27) on the B200 stack; --num_envs ensemble members per launch.
for s in 0 1 2 3 4; do
  for env in NonLinearWaterTankChangingParamUniformGoalStacking10-SquareDistance-v2 NonLinearWaterTankChangingParamUniformGoalStacking1-SquareDistance-v2; do
    python examples/train.py --fix_K --algo ResidualPPO --robust_test --net_dim 256 --env $env --target_step 2000 \
      --batch_size 256 --repeat_times 10 --break_step 400000 --eval_times1 50 --eval_times2 100 --eval_gap 1 \
      --seed $s --num_envs "${NUM_ENVS:-64}"
  done
done
