set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/ncu_box.sh r02_simtm_packed_kernel_msd_tsit5_f32 simtm_packed 524288000 python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm
bash tools/ncu_box.sh r02_simtm_kernel_pmsm_sat_tsit5_f32_epi simtm_kernel 65536000 python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
EXENV_PROFILE_ACT_SCALE=0.02 bash tools/ncu_box.sh r02_simtm_kernel_pmsm_sat_tsit5_f32_epi_gentle simtm_kernel 65536000 python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi
bash tools/ncu_box.sh r02_simtm_packed_kernel_pendulum_tsit5_f32 simtm_packed 524288000 python tools/profile_one.py sim pendulum tsit5 f32 1048576 500 0 0 tm
bash tools/ncu_box.sh r02_simtm_packed_kernel_pendulum_euler_f32 simtm_packed 1048576000 python tools/profile_one.py sim pendulum euler f32 1048576 1000 0 0 tm
bash tools/ncu_box.sh r02_simtm_packed_kernel_cartpole_tsit5_f32 simtm_packed 209715200 python tools/profile_one.py sim cartpole tsit5 f32 1048576 200 0 0 tm
bash tools/ncu_box.sh r02_simtm_packed_kernel_acrobot_tsit5_f32 simtm_packed 209715200 python tools/profile_one.py sim acrobot tsit5 f32 1048576 200 0 0 tm
bash tools/ncu_box.sh r02_step_kernel_cartpole_euler_f32 step_kernel 4194304 python tools/profile_one.py step cartpole euler f32 4194304
bash tools/ncu_box.sh r02_simtm_kernel_msd_tsit5_f64 simtm_kernel 262144000 python tools/profile_one.py sim msd tsit5 f64 262144 1000 0 0 tm
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d.get('cpu_baseline',{}).get('value'))
for c in d['extra']['configs']: print(c.get('name'), c.get('value'), c.get('ms_per_pass'), c.get('roofline',{}).get('bound'), c.get('roofline',{}).get('frac'), c.get('error'))"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c 1-400
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/ncu_launches.log 2>&1; grep -c simtm gpurun_out/r02_launches_bench.csv
