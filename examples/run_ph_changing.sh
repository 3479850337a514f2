#!/bin/bash
# run_ph_changing.sh of the reference (lines 1-33) on the B200 stack; --num_envs ensemble members per launch.
for s in 0 1 2 3 4; do
  python examples/train.py --fix_K --algo ResidualIntegratorModularPPO --gamma 0.98 --learning_rate 0.0003 \
    --env PH1DChangingParamUniformGoalIntegrator-SqaureDistance-v35 --net_dim 128 --target_step 1000 --batch_size 128 \
    --repeat_times 8 --lambda_gae_adv 0.99 --ratio_clip 0.2 --break_step 200000 --eval_times1 50 --eval_times2 100 \
    --eval_gap 1 --target_return 0 --test_render_times 50000 --seed $s --num_envs "${NUM_ENVS:-64}"
done
