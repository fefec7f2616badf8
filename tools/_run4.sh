python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r1c_n1.json 2> gpurun_out/bench_r1c_n1.err; tail -c 600 gpurun_out/bench_r1c_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1c_ref.json 2> gpurun_out/bench_r1c_ref.err; tail -c 400 gpurun_out/bench_r1c_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_d.log 2>&1
bash tools/final_table.sh > gpurun_out/final_table_r1e.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simtm --launch-skip 2 -c 1 -f -o gpurun_out/r01c_packed python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm > gpurun_out/ncu_r01c_packed.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simtm --launch-skip 2 -c 1 -f -o gpurun_out/r01c_packed_euler python tools/profile_one.py sim msd euler f32 524288 1000 0 0 tm > gpurun_out/ncu_r01c_packed_euler.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:simtm --launch-skip 2 -c 1 -f -o gpurun_out/r01c_acrobot python tools/profile_one.py sim acrobot tsit5 f32 1048576 200 0 0 tm > gpurun_out/ncu_r01c_acrobot.log 2>&1
python tools/ncu_summary.py gpurun_out/r01c_packed.ncu-rep 524288000 > gpurun_out/r01c_simtm_packed_kernel_msd_tsit5_f32.txt 2>&1
python tools/ncu_summary.py gpurun_out/r01c_packed_euler.ncu-rep 524288000 > gpurun_out/r01c_simtm_packed_kernel_msd_euler_f32.txt 2>&1
python tools/ncu_summary.py gpurun_out/r01c_acrobot.ncu-rep 209715200 > gpurun_out/r01c_simtm_kernel_acrobot_tsit5_f32.txt 2>&1
rm -f gpurun_out/r01c_packed_euler.ncu-rep gpurun_out/r01c_acrobot.ncu-rep
ls -la gpurun_out/
