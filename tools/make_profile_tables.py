"""Regenerates the measured tables of profiles/README.md in place from the raw files next to it
(r01_throughput_table_raw.txt, r01_bench_n*.json, r01_config1_latency.json).  Text between the tables is kept."""
import json
import re
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"
HBM = 6559.4e9
DIMS = {'msd': (1, 2, 2), 'pendulum': (1, 2, 2), 'tank': (1, 1, 1), 'cartpole': (1, 4, 4), 'acrobot': (1, 4, 4),
        'pmsm': (2, 8, 7), 'pmsm_sat': (2, 8, 7)}
NAMES = {'msd': 'mass_spring_damper', 'pendulum': 'pendulum', 'tank': 'fluid_tank', 'cartpole': 'cart_pole',
         'acrobot': 'acrobot', 'pmsm': 'pmsm (linear)', 'pmsm_sat': 'pmsm saturated (BRUSA LUT)'}


def e(v):
    return f'{v:.3g}'.replace('e+', 'e')


def replace_table(text, header_prefix, rows):
    """Replace the markdown table whose header line starts with `header_prefix`."""
    lines = text.split('\n')
    i = next(k for k, l in enumerate(lines) if l.startswith(header_prefix))
    j = i
    while j < len(lines) and lines[j].startswith('|'):
        j += 1
    return '\n'.join(lines[:i] + rows + lines[j:])


raw = (P / 'r01_throughput_table_raw.txt').read_text().split('\n')
sim, step = {}, {}
for l in raw:
    m = re.match(r'sim (\w+) (\w+) (f\d\d) (tm|bm)( epi)? B=(\d+) T=(\d+).*: ([\d.]+) ms  ([\d.e+]+) env-steps/s', l)
    if m:
        env, sol, dt, lay, epi, B, T, ms, v = m.groups()
        sim[(env, sol, dt, lay, bool(epi))] = (int(B), int(T), float(v))
    if l.startswith('{"kind": "step"'):
        d = json.loads(l)
        step[(d['env'], d['solver'], d['dtype'][-7:], d['B'])] = d
t1 = ['| env | solver | dtype | B × T | time-major env-steps/s | algorithmic GB/s | frac of 6559 GB/s | batch-major env-steps/s | frac | TM / BM |',
      '|---|---|---|---|---|---|---|---|---|---|']
for (env, sol, dt, lay, epi), (B, T, v) in sim.items():
    if lay != 'tm' or epi:
        continue
    A, O, D = DIMS[env]
    by = (4 if dt == 'f32' else 8) * (A + O + D)
    bm = sim[(env, sol, dt, 'bm', False)][2]
    t1.append(f'| {NAMES[env]} | {sol.capitalize()} | {dt} | 2^{B.bit_length()-1} × {T} | {e(v)} | {v*by/1e9:.0f} | '
              f'{v*by/HBM:.2f} | {e(bm)} | {bm*by/HBM:.2f} | {v/bm:.2f}× |')
t2 = ['| env (Tsit5, f32) | B × T | time-major + fused epilogue | batch-major + fused epilogue |', '|---|---|---|---|']
for (env, sol, dt, lay, epi), (B, T, v) in sim.items():
    if lay == 'tm' and epi:
        t2.append(f'| {NAMES[env]} | 2^{B.bit_length()-1} × {T} | {e(v)} | {e(sim[(env, sol, dt, "bm", True)][2])} |')
t3 = ['| env | Euler f32 | Euler f64 | Tsit5 f32 | Tsit5 f64 |', '|---|---|---|---|---|']
for en in ['pendulum', 'msd', 'tank', 'cartpole', 'acrobot', 'pmsm']:
    cells = [f"{e(step[(en, s_, d_, 1 << 24)]['env_steps_per_s'])} ({step[(en, s_, d_, 1 << 24)]['frac_hbm']:.2f})"
             for s_ in ('Euler()', 'Tsit5()') for d_ in ('float32', 'float64')]
    t3.append(f"| {NAMES[en]} | " + ' | '.join(cells) + ' |')
bench = {n: json.loads((P / f'r01_bench_n{n}.json').read_text()) for n in (1, 2, 4) if (P / f'r01_bench_n{n}.json').exists()}
b1 = bench[1]
th = ['| N GPUs | layout | env-steps/s (device resident) | ms / pass | roofline (20 B/env-step ÷ 6559.4 GB/s) | e2e through `exenv_sim_ahead_host` (PCIe) | CPU restatement (oracle, OpenMP) |',
      '|---|---|---|---|---|---|---|']
for n, b in bench.items():
    rel = '' if n == 1 else f" ({b['value']/b1['value']:.2f}× of N=1)"
    cpu = f"{e(b['cpu_baseline']['value'])} on {b['cpu_baseline']['cores']} cores" if 'cpu_baseline' in b else ''
    e2e = f"{e(b['e2e']['value'])} ({b['e2e'].get('batch_per_gpu', 0)} envs per rank)"
    th.append(f"| {n} | time-major (default) | **{e(b['value'])}**{rel} | {b['ms_per_step']:.2f} | "
              f"{b['roofline']['achieved']:.0f} GB/s = **{b['roofline']['frac']:.3f}** per GPU | {e2e} | {cpu} |")
    o = b['other_layout']
    th.append(f"| {n} | batch-major (reference-contiguous) | {e(o['value'])} | {o['ms_per_step']:.2f} | {o['roofline_frac']:.3f} | | |")
c1 = json.loads((P / 'r01_config1_latency.json').read_text().strip().split('\n')[-1])
tc = ['| how the 1000 steps are issued | µs per control step | env-steps/s |', '|---|---|---|',
      f"| Python mirror, one `vmap_step` call per step (eager) | {c1['us_per_step_eager']:.1f} | {e(c1['env_steps_per_s_eager'])} |",
      f"| the same 1000 launches captured once in a CUDA graph and replayed | {c1['us_per_step_graph']:.2f} | {e(c1['env_steps_per_s_graph'])} |",
      f"| one `vmap_sim_ahead` launch over the 1000 actions | {c1['us_per_step_sim_ahead']:.2f} | {e(c1['env_steps_per_s_sim_ahead'])} |"]
readme = (P / 'README.md').read_text()
readme = replace_table(readme, '| N GPUs | layout |', th)
readme = replace_table(readme, '| env | solver | dtype | B × T |', t1)
readme = replace_table(readme, '| env (Tsit5, f32) | B × T |', t2)
readme = replace_table(readme, '| env | Euler f32 | Euler f64 |', t3)
readme = replace_table(readme, '| how the 1000 steps are issued |', tc)
(P / 'README.md').write_text(readme)
print("profiles/README.md tables regenerated")
