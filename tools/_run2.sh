python -m pytest tests -m gpu -x -q 2>&1 | tail -15
P="python tools/profile_one.py"
echo "== thread sweep msd tsit5 f32 2^20"
$P sim msd tsit5 f32 1048576 2000 0 0 tm 2>&1 | tail -1
for t in 64 96 128 160 192 224 256; do echo "THREADS=$t"; EXENV_SIMTM_THREADS=$t $P sim msd tsit5 f32 1048576 2000 0 0 tm 2>&1 | tail -1; done
echo "== exact 3 rounds"
EXENV_SIMTM_THREADS=128 $P sim msd tsit5 f32 909312 2000 0 0 tm 2>&1 | tail -1
echo "== other sizes auto vs 128"
for b in 524288 2097152; do $P sim msd tsit5 f32 $b 1000 0 0 tm 2>&1 | tail -1; EXENV_SIMTM_AUTO_THREADS=0 $P sim msd tsit5 f32 $b 1000 0 0 tm 2>&1 | tail -1; done
echo "== f64"
$P sim msd tsit5 f64 524288 2000 0 0 tm 2>&1 | tail -1
EXENV_SIMTM_AUTO_THREADS=0 $P sim msd tsit5 f64 524288 2000 0 0 tm 2>&1 | tail -1
$P sim msd tsit5 f64 524288 2000 0 1 tm 2>&1 | tail -1
echo "== euler auto vs 128"
$P sim msd euler f32 1048576 2000 0 0 tm 2>&1 | tail -1
EXENV_SIMTM_AUTO_THREADS=0 $P sim msd euler f32 1048576 2000 0 0 tm 2>&1 | tail -1
echo "== other envs (auto, then 128)"
for cfg in "tank tsit5 f32 1048576 1000" "pendulum tsit5 f32 1048576 1000" "cartpole tsit5 f32 4194304 200" "acrobot tsit5 f32 4194304 200" "acrobot tsit5 f64 2097152 200" "cartpole tsit5 f64 2097152 200" "pmsm tsit5 f32 262144 2500"; do
  $P sim $cfg 0 0 tm 2>&1 | tail -1
  EXENV_SIMTM_AUTO_THREADS=0 $P sim $cfg 0 0 tm 2>&1 | tail -1
done
echo "== step tsit5"
for e in pendulum tank cartpole acrobot; do $P step $e tsit5 f32 16777216 2>&1 | tail -1; done
