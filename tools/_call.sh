set -x
python -m pytest tests -m gpu -q 2>&1 | tail -30
bash tools/ncu_box.sh r02c_headline simtm_packed 524288000 python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm
bash tools/ncu_box.sh r02c_pmsmsat_epi simtm_kernel 65536000 python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
python tools/profile_one.py sim msd tsit5 f64 524288 1000 0 0 tm
python tools/profile_one.py sim msd euler f32 1048576 1000 0 0 tm
python tools/profile_one.py sim pendulum tsit5 f32 1048576 500 0 0 tm
python tools/profile_one.py sim pendulum euler f32 1048576 1000 0 0 tm
python tools/profile_one.py sim pmsm tsit5 f32 262144 1000 0 0 tm
python tools/profile_one.py sim acrobot tsit5 f32 1048576 200 0 0 tm
python tools/profile_one.py sim cartpole tsit5 f32 1048576 200 0 0 tm
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -c 300 gpurun_out/bench_r02c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02c.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'])
for c in d['extra']['configs']: print(c.get('name'), c.get('value'), c.get('ms_per_pass'), c.get('roofline',{}).get('bound'), c.get('roofline',{}).get('frac'), c.get('error'))"
