set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for cfg in "msd tsit5 f32 1048576 2000" "msd euler f32 1048576 2000" "msd tsit5 f64 524288 2000" "msd euler f64 524288 2000"; do
  python tools/profile_one.py sim $cfg 0 0 tm 2>&1 | tail -1
done
python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 0 tm epi 2>&1 | tail -1
python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 0 bm 2>&1 | tail -1
python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 2 tm 2>&1 | tail -1
EXENV_SIMTM_PACKED=0 python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 4 tm 2>&1 | tail -1
python tools/profile_one.py sim pendulum tsit5 f32 1048576 1000 0 0 tm 2>&1 | tail -1
python tools/profile_one.py sim pmsm tsit5 f32 262144 2500 0 0 tm 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simtm --launch-skip 2 -c 1 -f -o gpurun_out/r01b_packed python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm > gpurun_out/ncu_r01b_packed.log 2>&1
ls -la gpurun_out/*.ncu-rep
