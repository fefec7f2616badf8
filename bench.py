#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched dynamics hot path on N B200s (driver contract in the task statement).

Workload (BASELINE.json configs[1]): MASS_SPRING_DAMPER, batch 2^20 environments PER GPU (weak scaling: the batch
axis shards with no collective), diffrax-Tsit5 `vmap_sim_ahead` over 2000 steps (obs_stepsize = action_stepsize =
tau = 1e-4), float32, synthetic uniform(-1, 1) states/actions.  One bench "step" = one sim_ahead pass over the batch
= B * 2000 env-steps, one launch of the persistent sm_100a kernel.

  value      env-steps/s with inputs and outputs resident in HBM (device-timed with CUDA events, max over ranks), in the
             layout named by --layout: `time_major` (default; the structure-of-arrays batch layout of the north star:
             same logical arrays as the reference, batch axis innermost in memory) or `batch_major` (the reference's
             C-contiguous arrays).  The other layout is timed in the same run and reported under "other_layout".
  e2e        the same call through the C ABI's host-buffer entry point (exenv_sim_ahead_host): pinned host inputs in
             the reference's layout are copied H2D, outputs D2H, inside the timed region
  roofline   algorithmic HBM bytes (S*(A+O+D) = 20 B per env-step) / measured launch time vs the measured copy peak
  cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/, OpenMP over all host cores) --
             JAX/diffrax are not installable in this image, so the reference's own jit(vmap) cannot be timed.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ENV, SOLVER, T_STEPS, TAU = "mass_spring_damper", "Tsit5", 2000, 1e-4
B_PER_GPU = 1 << 20
S, A, O, D = 4, 1, 2, 2
BYTES_PER_ENV_STEP = S * (A + O + D)  # actions in, observations + D trajectory leaves out


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=B_PER_GPU, help="environments per GPU (default: the BASELINE config)")
    p.add_argument("--layout", default="time_major", choices=["time_major", "batch_major"])
    p.add_argument("--no-other-layout", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="skip the extra.configs block (the other BASELINE workloads)")
    p.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi", "none"],
                   help="how SM clocks / throttle reasons are sampled during the timed region: an in-process NVML thread "
                        "(default; falls back to nvidia-smi when pynvml is missing), an `nvidia-smi -lms` child, or not at all")
    return p.parse_args()


def config_dict(args, batch):
    return {"workload": f"{ENV} batch={batch}/GPU Tsit5 vmap_sim_ahead {T_STEPS} steps tau={TAU} float32 "
                        f"(BASELINE.json configs[1])",
            "env": ENV, "solver": SOLVER, "batch_per_gpu": batch, "n_steps": T_STEPS, "tau": TAU,
            "layout": getattr(args, "layout", None),
            "l2_policy": "inputs (8.4 GB actions) and outputs (33.6 GB) are far larger than the 126 MB L2",
            "parallelism": f"batch-sharded x{args.gpus}, no collective"}


# ---- CPU arm: the restated reference path on the host cores ------------------------------------------------------------
def cpu_sample(batch, n_threads):
    """One pass of the oracle over `batch` environments; returns seconds."""
    import numpy as np

    from oracle import oracle as Or

    os.environ["OMP_NUM_THREADS"] = str(n_threads)
    orc = Or.Oracle(ENV, solver=Or.TSIT5, dtype=np.float32, tau=TAU)
    rng = np.random.default_rng(1234)
    state = [rng.uniform(-5, 5, batch).astype(np.float32) for _ in range(2)]
    actions = rng.uniform(-1, 1, (batch, T_STEPS, 1)).astype(np.float32)
    t0 = time.perf_counter()
    orc.sim_ahead(state, actions, TAU, TAU)
    return time.perf_counter() - t0


def cpu_baseline(target_seconds=12.0):
    cores = os.cpu_count() or 1
    probe_b = 1 << 10
    t = cpu_sample(probe_b, cores)
    rate = probe_b * T_STEPS / t
    batch = int(min(1 << 20, max(probe_b, 2 ** int(__import__("math").log2(max(1.0, rate * target_seconds / T_STEPS))))))
    t = cpu_sample(batch, cores)
    return {"value": batch * T_STEPS / t, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"oracle/exenv_oracle.c (C restatement of the reference path, OpenMP) on {batch} envs x "
                      f"{T_STEPS} Tsit5 steps, float32, {t:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    probe = cpu_sample(1 << 10, cores)
    rate = (1 << 10) * T_STEPS / probe
    batch = 1 << max(10, min(17, int(__import__("math").log2(max(1.0, rate * 4.0 / T_STEPS)))))
    for _ in range(args.warmup):
        cpu_sample(min(batch, 1 << 12), cores)
    times = [cpu_sample(batch, cores) for _ in range(args.steps)]
    tot = sum(times)
    value = batch * T_STEPS * args.steps / tot
    line = {"impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, args.batch),
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"each step = oracle C restatement (OpenMP, {cores} threads) over {batch} envs "
                                       f"x {T_STEPS} Tsit5 steps; the reference's JAX jit(vmap) path is not "
                                       f"installable here (no jax/diffrax wheels, no network)"},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(dev_index):
    """Run this rank (and first-touch its pinned host buffers) on the CPU cores next to its GPU: the host-buffer leg is
    PCIe bound and a far-socket buffer halves it.  Best effort: only cores this process is already allowed to use."""
    try:
        import pynvml
        import torch

        pr = torch.cuda.get_device_properties(dev_index)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        near = {64 * i + b for i, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(near & allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
        return {"near_gpu": len(near), "allowed": len(allowed), "bound_to": len(cpus) if cpus else len(allowed)}
    except Exception as exc:  # no NVML, no permission, ...: keep the inherited affinity
        return {"error": str(exc)[:80]}


# ---- clocks ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 50 ms from before the warm-up (its start-up can take longer than a short timed region);
    stop(t0, t1) keeps the samples whose own timestamps fall inside the timed region [t0, t1] (wall clock)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.06)  # one more polling period: the last sample of the region is on its way
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in open(self.f.name):
            r = line.strip().split(", ")
            if len(r) < 8 or not r[1].replace(".", "").isdigit():
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                ts = None
            rows.append((ts, r[1:]))
        os.unlink(self.f.name)
        inside = [r for ts, r in rows if ts is not None and t0 is not None and t0 - 0.02 <= ts <= t1 + 0.02]
        window = "timed region"
        if not inside:  # a region shorter than the polling period: the samples closest to it
            after = [r for ts, r in rows if ts is None or t0 is None or ts >= t0 - 0.5]
            inside, window = (after or [r for _, r in rows]), "nearest to the timed region"
        sm = sorted(float(r[0]) for r in inside)
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = max(float(r[1]) for r in inside)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any(r[3 + i] == "Active" for r in inside)]
            out["samples"] = len(sm)
            out["sm_mhz_min"] = sm[0]
            pw = [float(r[2]) for r in inside if r[2].replace(".", "").isdigit()]
            out["power_w_max"] = max(pw) if pw else None
            out["window"] = window
        return out


class NvmlClockSampler:
    """The same samples through NVML from a thread of this process (no child process polling the driver): SM clock,
    power and the clocks-event reasons every 20 ms with their wall-clock time stamps."""

    def __init__(self, gpu_index):
        import threading

        import pynvml

        self.nv = pynvml
        pynvml.nvmlInit()
        import torch

        pr = torch.cuda.get_device_properties(gpu_index)
        self.h = pynvml.nvmlDeviceGetHandleByPciBusId(
            f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        self.rows, self.stop_flag = [], False
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        try:
            self.power_limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
        except Exception:
            self.power_limit_w = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                try:
                    mem = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM))
                except Exception:
                    mem = 0.0
                self.rows.append((time.time(), float(mhz), pw, int(rs), mem))
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self, t0=None, t1=None):
        nv = self.nv
        time.sleep(0.03)
        self.stop_flag = True
        self.t.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "how": "NVML thread, 20 ms"}
        if self.power_limit_w:
            out["power_limit_w"] = self.power_limit_w
        inside = [r for r in self.rows if t0 is not None and t0 - 0.01 <= r[0] <= t1 + 0.01]
        window = "timed region"
        if not inside:
            inside, window = [r for r in self.rows if t0 is None or r[0] >= t0 - 0.5] or self.rows, "nearest to the timed region"
        if inside:
            sm = sorted(r[1] for r in inside)
            out.update(sm_mhz=sm[len(sm) // 2], sm_mhz_min=sm[0], samples=len(sm), window=window,
                       power_w_max=max(r[2] for r in inside))
            mem = sorted(r[4] for r in inside if len(r) > 4 and r[4] > 0)
            if mem:
                out.update(mem_mhz=mem[len(mem) // 2], mem_mhz_min=mem[0])
            bits = 0
            for r in inside:
                bits |= r[3]
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            out["reasons"] = [n for n, b in names.items() if bits & b]
        return out


def make_clock_sampler(kind, gpu_index):
    if kind == "none":
        return None
    if kind == "nvml":
        try:
            return NvmlClockSampler(gpu_index)
        except Exception:
            pass
    return ClockSampler(gpu_index)


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


KERNELS = {"time_major": ("simtm", "exenv::simtm_packed_kernel<MassSpringDamper<float>, Tsit5, V=4> (FFMA2)"),
           "batch_major": ("sim", "exenv::sim_kernel<MassSpringDamper<float>, Tsit5, float>")}


def ncu_facts(layout):
    """Facts about the dominant kernel from the committed `ncu --set full` summaries (profiles/sim_kernel_traffic.json):
    dram bytes per env-step (read + write), issue-slot and FMA-pipe utilisation."""
    p = ROOT / "profiles" / "sim_kernel_traffic.json"
    try:
        return json.load(open(p))[layout]
    except Exception:
        return {}


# ---- the other BASELINE.json workloads, timed in the same run (extra.configs) ---------------------------------------------
def fma_peaks(dev_index):
    """Measured FFMA / FFMA2 / DFMA and issue-slot peaks of this GPU (tools/microbench/fma_peak.cu, built by
    __graft_entry__.build(); a 0.1 s run).  None when the binary is missing."""
    exe = ROOT / "tools" / "microbench" / "_build" / "fma_peak"
    try:
        out = subprocess.run([str(exe), str(dev_index)], capture_output=True, text=True, timeout=60).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception as exc:
        return {"error": str(exc)[:120]}


def kernel_facts():
    """Per-kernel constants read off the committed ncu summaries (profiles/kernel_facts.json): warp instructions and
    FMA-pipe instructions per env-step -- the numerators of the issue / FMA rooflines of the instruction-bound kernels."""
    try:
        return json.load(open(ROOT / "profiles" / "kernel_facts.json"))
    except Exception:
        return {}


def run_extras(ex, torch, dist, dev, rank, world, peaks):
    """Each entry: one BASELINE workload other than the headline, device-resident, CUDA events, max over ranks."""
    import math

    hbm_peak = measured_peak_gbs()[0]
    facts = kernel_facts()
    f32 = torch.float32

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ex.launch_count()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n, (ex.launch_count() - l0) // n

    def entry(name, workload, env_steps_per_pass, sec, launches, bytes_per_env_step, kernel, fact_key=None, dtype="f32"):
        rate = env_steps_per_pass / sec
        hbm = bytes_per_env_step * rate / 1e9
        roof = {"bound": "hbm", "achieved": hbm, "peak": hbm_peak, "unit": "GB/s", "frac": hbm / hbm_peak,
                "algorithmic_bytes_per_env_step": bytes_per_env_step}
        f = facts.get(fact_key or "", {})
        issue_peak = (peaks or {}).get("issue_ginst")
        if f.get("warp_instructions_per_env_step") and issue_peak:
            # warp instructions per second = (instructions per env-step) x (env-steps/s) / 32 lanes
            ginst = f["warp_instructions_per_env_step"] * rate / 32 / 1e9
            issue = {"bound": "issue", "achieved": ginst, "peak": issue_peak, "unit": "G warp-instructions/s",
                     "frac": ginst / issue_peak, "warp_instructions_per_env_step": f["warp_instructions_per_env_step"],
                     "ncu_source": f.get("source")}
            if issue["frac"] > roof["frac"]:  # quote the roofline that binds, keep the other one beside it
                roof, issue = issue, roof
            roof["other"] = {k: issue[k] for k in ("bound", "achieved", "peak", "unit", "frac")}
        return {"name": name, "workload": workload, "value": world * rate, "unit": "env-steps/s", "dtype": dtype,
                "ms_per_pass": sec * 1e3, "launches_per_pass": int(launches), "n_gpus": world, "roofline": roof,
                "kernel": kernel}

    out = []

    def step_case(name, cls, solver, B, n_steps, kernel, fact_key=None, **kw):
        env = cls(batch_size=B, solver=solver, dtype=f32, device=dev, **kw)
        _, state = env.vmap_reset(rng=torch.Generator(device=dev).manual_seed(7 + rank))
        # one action tensor per control step (n_steps distinct ones: 200 x 2^22 x 4 B = 3.4 GB), states updated in place
        acts = torch.rand((n_steps, B, env.action_dim), device=dev) * 2 - 1
        box = {"s": state}

        def one_pass():
            st = box["s"]
            for t in range(n_steps):
                _, st = env.vmap_step(st, acts[t], inplace=True)
            box["s"] = st

        # The n_steps launches are captured ONCE into a CUDA graph and replayed: at 2^22 environments a step kernel
        # runs ~45 us, about what one eager vmap_step call costs in Python / ctypes / allocator time, so an eager loop
        # would time the host (profiles/README.md: 36 us per eager call, 3.9 us per replayed launch).
        one_pass()
        torch.cuda.synchronize(dev)
        graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                one_pass()
        torch.cuda.synchronize(dev)
        sec, _ = timed(graph.replay, 3, warm=2)
        D, A, O = env.physical_state_dim, env.action_dim, env.obs_dim
        tsit = isinstance(solver, ex.Tsit5)
        bpe = 4 * (2 * D + A + O + (env.N_ODE if tsit else 0))
        out.append(entry(name, f"{name}: batch={B}/GPU, {n_steps} vmap_step calls (one launch each, captured once into a "
                               f"CUDA graph and replayed), float32", B * n_steps, sec, n_steps, bpe, kernel, fact_key))
        del acts, state, box, graph

    def sim_case(name, cls, solver, B, T, kernel, fact_key=None, epilogue=False, chunks=1, dtype=f32, **kw):
        env = cls(batch_size=B, solver=solver, dtype=dtype, device=dev, **kw)
        env.max_steps = None if T // chunks > 4096 else 4096
        _, state = env.vmap_reset(rng=torch.Generator(device=dev).manual_seed(11 + rank))
        D, A, O = env.physical_state_dim, env.action_dim, env.obs_dim
        Tc = T // chunks
        scale = 0.02 if cls is ex.PMSM else 1.0  # voltages that keep the currents inside the flux maps
        acts = [((torch.rand((Tc, B, A), device=dev, dtype=dtype) * 2 - 1) * scale).permute(1, 0, 2) for _ in range(chunks)]
        outb = dict(obs=torch.empty((Tc + 1, B, O), device=dev, dtype=dtype),
                    traj=[torch.empty((Tc + 1, B), device=dev, dtype=dtype) for _ in range(D)])

        def one_pass():
            st = state
            for c in range(chunks):  # a chunk continues from the previous one's last_state (diffrax: max_steps = 4096)
                r = env._sim_impl(env.env_properties, B, st, acts[c], env.tau, env.tau, out=outb, layout="time_major",
                                  epilogue=epilogue)
                st = r[2]

        sec, launches = timed(one_pass, 3, warm=2)
        S = 4 if dtype == f32 else 8
        bpe = S * (A + O + D) + ((S + env.n_trunc + 1) if epilogue else 0)
        what = "vmap_sim_ahead" + (" + fused reward/truncated/terminated" if epilogue else "")
        ch = f" as {chunks} x {Tc}-step chunks" if chunks > 1 else ""
        out.append(entry(name, f"{name}: batch={B}/GPU, {what}, {T} steps{ch}, time-major, "
                               f"{'float32' if dtype == f32 else 'float64'}", B * T, sec, launches, bpe, kernel, fact_key,
                         dtype="f32" if dtype == f32 else "f64"))
        del acts, outb, state
        torch.cuda.empty_cache()

    def guarded(fn, *a, **kw):
        try:
            fn(*a, **kw)
        except Exception as exc:  # one failing extra must not take the headline down
            out.append({"name": a[0], "error": str(exc)[:200]})
        torch.cuda.empty_cache()

    E, T5 = ex.Euler(), ex.Tsit5()
    # north-star target 1: pendulum, explicit Euler (>= 1e11 env-steps/s on 8 GPUs)
    guarded(step_case, "pendulum Euler vmap_step", ex.Pendulum, E, 1 << 24, 50, "step_kernel<Pendulum<float>, Euler, VEC=4>",
            "step_pendulum_euler_f32")
    guarded(sim_case, "pendulum Euler vmap_sim_ahead", ex.Pendulum, E, 1 << 20, 1000,
            "simtm_packed_kernel<Pendulum<float>, Euler, V=4>", "simtm_pendulum_euler_f32")
    # north-star target 2 = BASELINE configs[3]: PMSM with saturated flux maps, Tsit5, fused epilogue (>= 1e10 on 8 GPUs);
    # 2^17 environments per GPU is the shard of the 2^20 batch on 8 GPUs, 5000 steps as two 2500-step chunks
    guarded(sim_case, "pmsm saturated (BRUSA LUT) Tsit5 vmap_sim_ahead + epilogue", ex.PMSM, T5, 1 << 17, 5000,
            "simtm_kernel<Pmsm<float, saturated>, Tsit5, EPI>", "simtm_pmsm_sat_tsit5_f32_epi", epilogue=True, chunks=2,
            saturated=True, motor_variant=ex.MotorVariant.BRUSA)
    # BASELINE configs[2]: cart-pole and acrobot, B = 2^22, 200-step horizon
    for nm, cls in (("cart_pole", ex.CartPole), ("acrobot", ex.Acrobot)):
        guarded(step_case, f"{nm} Euler vmap_step", cls, E, 1 << 22, 200, f"step_kernel<{cls.__name__}<float>, Euler, VEC=4>",
                f"step_{nm}_euler_f32")
        guarded(sim_case, f"{nm} Tsit5 vmap_sim_ahead", cls, T5, 1 << 22, 200, f"simtm_packed_kernel<{cls.__name__}<float>, Tsit5, V=2>",
                f"simtm_{nm}_tsit5_f32")
    # BASELINE configs[1], the fp64 half
    guarded(sim_case, "mass_spring_damper Tsit5 vmap_sim_ahead fp64", ex.MassSpringDamper, T5, 1 << 20, 2000,
            "simtm_kernel<MassSpringDamper<double>, Tsit5>", "simtm_msd_tsit5_f64", dtype=torch.float64)
    return out


def run_b200(args):
    import torch

    import exciting_environments_b200 as ex

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout (NCCL_DEBUG=VERSION from the environment or from /etc/nccl.conf); stdout
        # carries exactly one JSON line, so the communicator is created with fd 1 pointing at stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    B = args.batch
    env = ex.MassSpringDamper(batch_size=B, solver=ex.Tsit5(), tau=TAU, dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1234 + 1000 * 2 + rank)
    _, state = env.vmap_reset(rng=gen)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def buffers(layout):
        """Synthetic actions and preallocated outputs in the physical layout of `layout`."""
        g = torch.Generator(device=dev).manual_seed(99 + rank)
        if layout == "time_major":
            acts = (torch.rand((T_STEPS, B, A), device=dev, generator=g) * 2 - 1).permute(1, 0, 2)  # logical (B, T, A)
            out = dict(obs=torch.empty((T_STEPS + 1, B, O), device=dev),
                       traj=[torch.empty((T_STEPS + 1, B), device=dev) for _ in range(D)])
        else:
            acts = torch.rand((B, T_STEPS, A), device=dev, generator=g) * 2 - 1
            out = dict(obs=torch.empty((B, T_STEPS + 1, O), device=dev),
                       traj=[torch.empty((B, T_STEPS + 1), device=dev) for _ in range(D)])
        return acts, out

    def timed(layout, n_steps, n_warm, sample_clocks):
        acts, out = buffers(layout)

        def step():
            return env._sim_impl(env.env_properties, B, state, acts, TAU, TAU, out=out, layout=layout)

        sampler = make_clock_sampler(args.clock_sampler, local_rank) if (sample_clocks and rank == 0) else None
        for _ in range(n_warm):
            step()
        barrier()
        launches0 = ex.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        for _ in range(n_steps):
            step()
        e1.record()
        barrier()
        t_wall1 = time.time()
        launches = ex.launch_count() - launches0
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        elapsed = e0.elapsed_time(e1) * 1e-3
        t = torch.tensor([elapsed], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return dict(elapsed=elapsed, elapsed_max=float(t.item()), launches=launches, clocks=clocks, acts=acts, out=out)

    main = timed(args.layout, args.steps, args.warmup, True)
    elapsed, elapsed_max, launches, clocks = main["elapsed"], main["elapsed_max"], main["launches"], main["clocks"]
    value = world * B * T_STEPS * args.steps / elapsed_max
    other = None
    if not args.no_other_layout:
        main.pop("acts"), main.pop("out")  # free 42 GB before the second set of buffers
        torch.cuda.empty_cache()
        o_layout = "batch_major" if args.layout == "time_major" else "time_major"
        n_o = max(1, min(args.steps, 5))
        r = timed(o_layout, n_o, 3, False)
        other = {"layout": o_layout, "value": world * B * T_STEPS * n_o / r["elapsed_max"], "unit": "env-steps/s",
                 "ms_per_step": r["elapsed_max"] / n_o * 1e3, "steps": n_o,
                 "roofline_frac": BYTES_PER_ENV_STEP * B * T_STEPS / (r["elapsed"] / n_o) / 1e9 / measured_peak_gbs()[0],
                 "kernel": KERNELS[o_layout][1]}
        del r
    main.pop("acts", None), main.pop("out", None)
    torch.cuda.empty_cache()

    # ---- end to end through the C ABI with host buffers (rank-local, then max over ranks)
    e2e = None
    if not args.no_e2e:
        L = ex._lib
        NS = T_STEPS + 1
        # Pinned host buffers hold 40 KB per environment (42 GB per rank at the full batch).  All ranks of the node
        # share one host memory: the e2e batch per rank is the largest power of two that keeps the node's pinned
        # total under min(96 GB, 40 % of the available host memory) -- the full batch at N = 1 and 2, a shard of it
        # at N = 4 and 8 (PCIe bound either way: the rate does not depend on the batch).
        per_env_host = 4 * (T_STEPS * A + NS * (O + D) + 2 * D)
        budget = 96e9
        try:
            import psutil

            budget = min(budget, 0.4 * psutil.virtual_memory().available)
        except Exception:
            pass
        Be = B
        while Be > 1024 and world * Be * per_env_host > budget:
            Be //= 2
        B_full, B = B, Be  # the e2e section below works on the first `Be` environments of this rank
        try:
            h_state = [getattr(state.physical_state, n)[:B].cpu().pin_memory() for n in env.PHYS_FIELDS]
            g = torch.Generator().manual_seed(7 + rank)
            h_act = (torch.rand((B, T_STEPS, A), generator=g) * 2 - 1).pin_memory()
            h_obs = torch.empty((B, NS, O), dtype=torch.float32).pin_memory()
            h_traj = [torch.empty((B, NS), dtype=torch.float32).pin_memory() for _ in range(D)]
            h_last = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(D)]
            a = L.ExenvSimArgs()
            a.batch, a.n_actions, a.obs_stepsize, a.action_stepsize, a.max_steps = B, T_STEPS, TAU, TAU, 4096
            for i in range(D):
                a.state_in[i], a.traj[i], a.last[i] = h_state[i].data_ptr(), h_traj[i].data_ptr(), h_last[i].data_ptr()
            a.actions, a.obs = h_act.data_ptr(), h_obs.data_ptr()
            plan = env._plan().handle
            n_e2e = max(1, min(args.steps, 3))
            L.check(L.lib().exenv_sim_ahead_host(plan, C.byref(a), 0))  # warm-up (allocations, first touch)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                L.check(L.lib().exenv_sim_ahead_host(plan, C.byref(a), 0))
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], device=dev, dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            h2d = h_act.numel() * 4 + sum(x.numel() for x in h_state) * 4
            d2h = h_obs.numel() * 4 + sum(x.numel() for x in h_traj) * 4 + sum(x.numel() for x in h_last) * 4
            # check a slice of the host-buffer result against a device-resident call on the same 64 environments
            env64 = ex.MassSpringDamper(batch_size=64, solver=ex.Tsit5(), tau=TAU, dtype=torch.float32, device=dev)
            _, st64 = env64.vmap_reset()
            st64.physical_state = env64.PhysicalState(*[x[:64].to(dev) for x in h_state])
            chk = env64.vmap_sim_ahead(st64, h_act[:64].to(dev), TAU, TAU)[0].cpu()
            ok = bool(torch.equal(h_obs[:64], chk))
            sec = float(tt.item()) / n_e2e
            e2e = {"value": world * B * T_STEPS * n_e2e / float(tt.item()), "unit": "env-steps/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
                   "batch_per_gpu": B, "matches_device_path": ok, "cpu_affinity": affinity,
                   "h2d_gbs_per_gpu": h2d / sec / 1e9, "d2h_gbs_per_gpu": d2h / sec / 1e9,
                   "bound": "the host link: every pass moves h2d + d2h bytes per GPU between pinned host memory and "
                            "HBM (copies of both directions overlap the kernel on 3 streams); the kernel itself needs "
                            "well under 1 % of this time and there is no collective -- at N > 1 the ranks share the "
                            "host's memory / PCIe root bandwidth, which is what limits the e2e scaling curve",
                   "how": "exenv_sim_ahead_host: pinned host buffers, chunked H2D / kernel / D2H on 3 streams kept in "
                          "the plan, wall clock around the blocking C call (max over ranks)"}
            del h_obs, h_traj, h_act
        except Exception as exc:  # e.g. not enough host memory to pin the buffers
            e2e = {"value": None, "unit": "env-steps/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                   "error": str(exc)[:200]}
        B = B_full

    extra = None
    if not args.no_extra:
        peaks = fma_peaks(local_rank) if rank == 0 else None
        if dist is not None:
            box = [peaks]
            dist.broadcast_object_list(box, src=0)
            peaks = box[0]
        extra = {"fma_peak": peaks, "configs": run_extras(ex, torch, dist, dev, rank, world, peaks)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        facts = ncu_facts(args.layout)
        launch_s = elapsed / args.steps  # one launch per step on this rank
        achieved = BYTES_PER_ENV_STEP * B * T_STEPS / launch_s / 1e9
        line = {"metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_max / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config_dict(args, B), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak,
                             "traffic": (facts["dram_bytes_per_env_step"] * B * T_STEPS) if facts else None,
                             "peak_source": peak_src, "kernel": KERNELS[args.layout][1],
                             "algorithmic_bytes_per_env_step": BYTES_PER_ENV_STEP,
                             "ncu": {k: facts.get(k) for k in ("issue_slots_busy", "fma_pipe_busy", "dram_pct_of_peak",
                                                                "warp_instructions_per_env_step", "source")},
                             "note": "the Tsit5 trajectory kernel sits between the fp32 FMA pipe (7 right-hand sides + dense "
                                     "output per env-step: 83 fp32 operations as 41.5 packed instructions, pipe ~62 % busy) "
                                     "and HBM (~63 % of peak under ncu); 20 back-to-back passes run under the board's "
                                     "power management (SM clock dips to ~1.75-1.85 GHz), a 3-pass burst is ~7 % faster; "
                                     "see profiles/README.md"},
                "other_layout": other, "extra": extra}
        if extra and extra.get("fma_peak") and extra["fma_peak"].get("ffma2_tflops"):
            # the headline kernel against the fp32 FMA pipe as well: packed (FFMA2 / FMUL2 / FADD2) warp instructions per
            # second vs the measured FFMA2 instruction rate (one FFMA2 warp instruction = 128 flop)
            fp = facts.get("packed_fp32_instructions_per_env_step")
            if fp:
                ginst = fp * B * T_STEPS / launch_s / 32 / 1e9
                peak_ginst = extra["fma_peak"]["ffma2_tflops"] * 1e12 / 128 / 1e9
                line["roofline"]["fma"] = {"achieved": ginst, "peak": peak_ginst, "unit": "G packed warp-instructions/s",
                                           "frac": ginst / peak_ginst, "packed_fp32_instructions_per_env_step": fp}
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_baseline()
            except Exception as exc:
                line["cpu_baseline"] = {"value": None, "error": str(exc)[:200]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
