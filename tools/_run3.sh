python -m pytest tests -m gpu -x -q 2>&1 | tail -4
P="python tools/profile_one.py"
S=/root/repo/exciting-environments_b200/lib_split/libexenv_b200.so
for cfg in "msd tsit5 f32 1048576 2000" "msd euler f32 1048576 2000" "pendulum tsit5 f32 1048576 1000" "pendulum euler f32 1048576 1000" "tank tsit5 f32 1048576 1000" "cartpole tsit5 f32 4194304 200" "acrobot tsit5 f32 4194304 200" "pmsm tsit5 f32 262144 2500" "pmsm_sat tsit5 f32 131072 2500"; do
  echo "main : $($P sim $cfg 0 0 tm 2>&1 | tail -1)"
  echo "split: $(EXENV_B200_LIB=$S $P sim $cfg 0 0 tm 2>&1 | tail -1)"
done
echo "main : $($P sim msd tsit5 f32 1048576 2000 0 0 tm epi 2>&1 | tail -1)"
echo "split: $(EXENV_B200_LIB=$S $P sim msd tsit5 f32 1048576 2000 0 0 tm epi 2>&1 | tail -1)"
echo "f64 v1: $($P sim msd tsit5 f64 524288 2000 0 0 tm 2>&1 | tail -1)"
