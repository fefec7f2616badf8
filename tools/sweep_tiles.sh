#!/bin/bash
# K sweep for the sim kernel across envs / dtypes (developer tool)
for cfg in "msd tsit5 f32 1048576 500" "msd tsit5 f64 524288 500" "msd euler f64 524288 500" "tank tsit5 f32 1048576 500" "pendulum tsit5 f32 1048576 500" "pendulum tsit5 f64 524288 500" "cartpole tsit5 f32 1048576 200" "cartpole euler f32 1048576 200" "cartpole tsit5 f64 524288 200" "acrobot tsit5 f32 1048576 200" "acrobot tsit5 f64 524288 200" "pmsm tsit5 f32 262144 500" "pmsm euler f32 262144 500" "pmsm tsit5 f64 131072 500" "pmsm_sat tsit5 f32 131072 500" "pmsm_sat tsit5 f64 65536 500"; do
  for k in 8 16 24 32 48; do
    python tools/profile_one.py sim $cfg $k 2>&1 | tail -1
  done
done
