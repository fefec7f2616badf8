#!/bin/bash
# Device-timed throughput tables for profiles/README.md: vmap_step, vmap_sim_ahead in both layouts (auto tile / auto V),
# the fused-epilogue variants and the launch-latency config.
python tools/quick_bench.py step 2>/dev/null
for lay in tm bm; do
for cfg in "msd tsit5 f32 1048576 2000" "msd euler f32 1048576 2000" "msd tsit5 f64 524288 2000" "msd euler f64 524288 2000" "tank tsit5 f32 1048576 1000" "tank euler f32 1048576 1000" "pendulum tsit5 f32 1048576 1000" "pendulum euler f32 1048576 1000" "pendulum tsit5 f64 524288 1000" "cartpole tsit5 f32 4194304 200" "cartpole euler f32 4194304 200" "cartpole tsit5 f64 2097152 200" "acrobot tsit5 f32 4194304 200" "acrobot euler f32 4194304 200" "acrobot tsit5 f64 2097152 200" "pmsm tsit5 f32 262144 2500" "pmsm euler f32 262144 2500" "pmsm tsit5 f64 131072 2500" "pmsm_sat tsit5 f32 131072 2500" "pmsm_sat euler f32 131072 2500" "pmsm_sat tsit5 f64 65536 2500"; do
  python tools/profile_one.py sim $cfg 0 0 $lay 2>&1 | tail -1
done
done
for cfg in "msd tsit5 f32 1048576 2000" "pendulum tsit5 f32 1048576 1000" "cartpole tsit5 f32 4194304 200" "pmsm tsit5 f32 262144 2500" "pmsm_sat tsit5 f32 131072 2500"; do
  python tools/profile_one.py sim $cfg 0 0 tm epi 2>&1 | tail -1
  python tools/profile_one.py sim $cfg 0 0 bm epi 2>&1 | tail -1
done
python tools/quick_bench.py config1 2>/dev/null
