set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm > gpurun_out/p1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:simtm_packed -s 2 -c 1 -f -o gpurun_out/r02_headline_before python tools/profile_one.py sim msd tsit5 f32 524288 1000 0 4 tm > gpurun_out/ncu1.log 2>&1
python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi > gpurun_out/p2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:simtm_kernel -s 2 -c 1 -f -o gpurun_out/r02_pmsmsat_before python tools/profile_one.py sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi > gpurun_out/ncu2.log 2>&1
python tools/profile_one.py step pendulum euler f32 16777216 > gpurun_out/p3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 2 -c 1 -f -o gpurun_out/r02_step_pendulum_euler_f32 python tools/profile_one.py step pendulum euler f32 16777216 > gpurun_out/ncu3.log 2>&1
python tools/profile_one.py sim pendulum tsit5 f32 1048576 500 0 0 tm > gpurun_out/p4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:simtm_kernel -s 2 -c 1 -f -o gpurun_out/r02_pendulum_tsit5_before python tools/profile_one.py sim pendulum tsit5 f32 1048576 500 0 0 tm > gpurun_out/ncu4.log 2>&1
cat gpurun_out/p1.log gpurun_out/p2.log gpurun_out/p3.log gpurun_out/p4.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_before.json 2> gpurun_out/bench_before.err; tail -c 600 gpurun_out/bench_before.json
ls -la gpurun_out
