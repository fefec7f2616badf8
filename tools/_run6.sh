python -m pytest tests -m gpu -q 2>&1 | tail -4
P="python tools/profile_one.py"
$P sim msd tsit5 f32 1048576 2000 0 0 tm 2>&1 | tail -1
$P sim msd euler f32 1048576 2000 0 0 tm 2>&1 | tail -1
$P sim msd tsit5 f64 524288 2000 0 1 tm 2>&1 | tail -1
$P sim msd tsit5 f64 524288 2000 0 2 tm 2>&1 | tail -1
$P sim msd euler f64 524288 2000 0 0 tm 2>&1 | tail -1
$P sim msd euler f64 524288 2000 0 2 tm 2>&1 | tail -1
$P sim msd tsit5 f32 1048576 2000 0 0 tm epi 2>&1 | tail -1
$P sim pendulum tsit5 f32 1048576 1000 0 0 tm 2>&1 | tail -1
$P sim pendulum tsit5 f64 524288 1000 0 0 tm 2>&1 | tail -1
$P sim cartpole tsit5 f32 4194304 200 0 0 tm 2>&1 | tail -1
$P sim acrobot tsit5 f32 4194304 200 0 0 tm 2>&1 | tail -1
$P sim acrobot tsit5 f64 2097152 200 0 0 tm 2>&1 | tail -1
$P sim pmsm tsit5 f32 262144 2500 0 0 tm 2>&1 | tail -1
$P sim pmsm_sat tsit5 f32 131072 2500 0 0 tm 2>&1 | tail -1
$P sim tank tsit5 f32 1048576 1000 0 0 tm 2>&1 | tail -1
