set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 4 tm
python tools/profile_one.py sim msd euler f32 1048576 2000 0 4 tm
python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 4 tm epi
bash tools/ncu_box.sh r02e_headline simtm_packed 524288000 python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra --no-e2e --no-other-layout | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])"
EXENV_PROFILE_ACT_SCALE=0.02 bash tools/ncu_box.sh r02e_pmsmsat_epi_gentle simtm_kernel 65536000 python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
