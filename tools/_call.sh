set -x
python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== headline"
python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 4 tm
echo "== packed vs scalar, other envs"
for cfg in "pendulum tsit5 1048576 1000" "pendulum euler 1048576 1000" "tank tsit5 1048576 1000" "tank euler 1048576 1000" "cartpole tsit5 4194304 200" "cartpole euler 4194304 200" "acrobot tsit5 4194304 200" "acrobot euler 4194304 200"; do set -- $cfg
for blk in 2 4; do python tools/profile_one.py sim $1 $2 f32 $3 $4 0 $blk tm; done
EXENV_SIMTM_PACKED=0 python tools/profile_one.py sim $1 $2 f32 $3 $4 0 0 tm
done
echo "== pmsm sat full range / gentle"
python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
EXENV_PROFILE_ACT_SCALE=0.02 python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
bash tools/ncu_box.sh r02g_pendulum_tsit5 simtm_packed 524288000 python tools/profile_one.py sim pendulum tsit5 f32 1048576 500 0 2 tm
bash tools/ncu_box.sh r02g_pmsmsat_epi simtm_kernel 65536000 python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
