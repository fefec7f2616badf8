"""Rebuilds profiles/kernel_facts.json from the committed ncu summaries (tools/ncu_box.sh writes them).

bench.py reads the file for the numerators of the issue / FMA rooflines of the instruction-bound kernels: warp
instructions and packed-fp32 instructions per env-step.  One row per summary; B x T is the workload of the profiled
launch (the tools/profile_one.py command in `source`)."""
import json
import re
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"

# key: (summary file, envs, steps per launch, profile_one.py arguments)
ROWS = {
    "simtm_msd_tsit5_f32": ("r02_simtm_packed_kernel_msd_tsit5_f32.txt", 524288, 1000, "sim msd tsit5 f32 524288 1000 0 4 tm"),
    "simtm_pmsm_sat_tsit5_f32_epi": ("r02_simtm_kernel_pmsm_sat_coef_tsit5_f32_epi_gentle.txt", 131072, 500,
                                     "EXENV_PROFILE_ACT_SCALE=0.02 sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi"),
    "simtm_pmsm_sat_tsit5_f32_epi_full_range": ("r02_simtm_kernel_pmsm_sat_coef_tsit5_f32_epi.txt", 131072, 500,
                                                "sim pmsm_sat tsit5 f32 131072 500 0 0 tm epi"),
    "simtm_pendulum_euler_f32": ("r02_simtm_packed_kernel_pendulum_euler_f32.txt", 1048576, 1000,
                                 "sim pendulum euler f32 1048576 1000 0 0 tm"),
    "simtm_pendulum_tsit5_f32": ("r02_simtm_packed_kernel_pendulum_tsit5_f32.txt", 1048576, 500,
                                 "sim pendulum tsit5 f32 1048576 500 0 0 tm"),
    "simtm_cart_pole_tsit5_f32": ("r02_simtm_packed_kernel_cartpole_tsit5_f32.txt", 1048576, 200,
                                  "sim cartpole tsit5 f32 1048576 200 0 0 tm"),
    "simtm_acrobot_tsit5_f32": ("r02_simtm_packed_kernel_acrobot_tsit5_f32.txt", 1048576, 200,
                                "sim acrobot tsit5 f32 1048576 200 0 0 tm"),
    "simtm_msd_tsit5_f64": ("r02_simtm_kernel_msd_tsit5_f64.txt", 262144, 1000, "sim msd tsit5 f64 262144 1000 0 0 tm"),
    "step_pendulum_euler_f32": ("r02_step_kernel_pendulum_euler_f32.txt", 16777216, 1, "step pendulum euler f32 16777216"),
    "step_cart_pole_euler_f32": ("r02_step_kernel_cartpole_euler_f32.txt", 4194304, 1, "step cartpole euler f32 4194304"),
    "step_acrobot_euler_f32": ("r02_step_kernel_acrobot_euler_f32.txt", 4194304, 1, "step acrobot euler f32 4194304"),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ms": 1.0, "s": 1e3}


def metric(text, name):
    m = re.search(rf"^\s*{re.escape(name)}\s+([\d.]+)\s*(\S*)\s*$", text, re.M)
    if not m:
        return None
    return float(m.group(1)) * UNIT.get(m.group(2), 1.0)


def main():
    facts = {}
    for key, (fname, envs, steps, cmd) in ROWS.items():
        f = P / fname
        if not f.exists():
            continue
        t = f.read_text()
        mix = re.search(r"instruction mix per env-step: (.*)", t)
        ops = dict((a, float(b)) for a, b in re.findall(r"(\w+) ([\d.]+)", mix.group(1))) if mix else {}
        packed = sum(ops.get(o, 0.0) for o in ("FFMA2", "FMUL2", "FADD2"))
        instr = re.search(r"warp instructions per warp-env-step: ([\d.]+)", t)
        rd, wr = metric(t, "dram__bytes_read.sum"), metric(t, "dram__bytes_write.sum")
        pct = lambda n: (lambda v: None if v is None else v / 100.0)(metric(t, n))
        facts[key] = {
            "warp_instructions_per_env_step": float(instr.group(1)) if instr else None,
            "packed_fp32_instructions_per_env_step": round(packed, 1) if packed else None,
            "issue_slots_busy": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_busy": pct("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "dram_pct_of_peak": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": int(metric(t, "launch__registers_per_thread") or 0),
            "time_ms": metric(t, "gpu__time_duration.sum"),
            "dram_bytes_per_env_step": None if rd is None or wr is None else (rd + wr) / (envs * steps),
            "source": "ncu --set full --clock-control none, " + re.sub(r"^((?:\w+=\S+ )*)", r"\1tools/profile_one.py ", cmd)
                      + f" (profiles/{fname})",
        }
    (P / "kernel_facts.json").write_text(json.dumps(facts, indent=1))
    for k, v in facts.items():
        print(f"{k:44s} {v['warp_instructions_per_env_step']:8.1f} instr  {v['time_ms']:9.4f} ms  {v['registers']:4d} regs")


if __name__ == "__main__":
    main()
