set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
bash tools/ncu_box.sh r02a_headline simtm_packed 524288000 python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm
bash tools/ncu_box.sh r02a_pmsmsat_epi simtm_kernel 65536000 python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
bash tools/ncu_box.sh r02_step_pendulum_euler_f32 step_kernel 16777216 python tools/profile_one.py step pendulum euler f32 16777216
bash tools/ncu_box.sh r02a_pendulum_tsit5 simtm_kernel 524288000 python tools/profile_one.py sim pendulum tsit5 f32 1048576 500 0 0 tm
ls -la gpurun_out
