python -m pytest tests -m gpu -q 2>&1 | tail -25
for s in 3 10 20 40; do
  python bench.py --steps $s --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps',d['steps'],'ms',d['ms_per_step'],d['clocks'],'other',d['other_layout']['ms_per_step'])"
done
python tools/profile_one.py sim msd tsit5 f32 1048576 2000 0 0 tm 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1d_n1.json 2> gpurun_out/bench_r1d_n1.err; tail -c 300 gpurun_out/bench_r1d_n1.json
