#!/usr/bin/env python
"""Headline benchmark: Jitterbug control steps/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # CUDA path, one rank per GPU (torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the CPU path (fp64 oracle port) on the host cores

One "step" = one control step of every env of the batch: 50 physics sub-steps + reward + observation
(jitterbug_dmc/jitterbug.xml:18, jitterbug_dmc/jitterbug.py:56-57).  Workload: move_from_origin, uniform random
actions in [-1, 1], E envs per GPU (weak scaling), synthetic data (Philox reset distribution + torch-generated
actions).  E defaults to 1,048,576 so the per-step state traffic (~400 MB) is larger than the 126 MB L2.

Prints ONE JSON line (see the task contract): value = device-resident throughput, e2e = through the host-buffer
C-ABI call (numpy in, numpy out, copies timed), roofline (HBM line, required schema) + fp32 (the real bound of this
compute-bound kernel), cpu_baseline (oracle port on the host cores, bounded sample), clocks, gpu_launches.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "jitterbug_control_steps_per_sec"
UNIT = "control_steps/s"
TASK = "move_from_origin"
# algorithmic bytes per env per control step: read qpos16+rev1+qvel15+warm15+target4+task1+
# action1+count/flags(3), write qpos16+rev1+qvel15+warm15 + obs15 + reward1 + step_type1 + count1, 4 B each
BYTES_PER_ENV_STEP = (16 + 1 + 15 + 15 + 4 + 1 + 1 + 3 + 16 + 1 + 15 + 15 + 15 + 1 + 1 + 1) * 4


def _flop_model():
    try:
        with open(os.path.join(ROOT, "profiles", "flop_model.json")) as f:
            return json.load(f)["fit"]
    except Exception:
        return {"per_substep": 7376.0, "per_newton_iteration": 1948.0, "per_contact_substep": 1272.0}


def _traffic_per_env():
    """DRAM bytes per env per launch from the committed ncu --set full capture (profiles/r01_v8_ncu_summary.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_v8_ncu_summary.json")) as f:
            return float(json.load(f)["dram_traffic_bytes_per_env"])
    except Exception:
        return None


def _measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def cpu_rollout_rate(n_env: int, n_steps: int, threads: int, seed: int = 0):
    """Control steps/s of the fp64 oracle port on `threads` host threads (one env per worker at a time)."""
    from oracle import oracle as O
    from jitterbug_b200 import mjcf
    cm = mjcf.compile_xml()
    L = O.lib()
    rng = np.random.default_rng(seed)
    models = (O.JboModel * n_env)()
    datas = (O.JboData * n_env)()
    proto = O.make_model(cm)
    for e in range(n_env):
        C.memmove(C.byref(models[e]), C.byref(proto), C.sizeof(O.JboModel))
        L.jbo_reset_data(C.byref(models[e]), C.byref(datas[e]))
        a = rng.uniform(0, 2 * np.pi)
        ax = np.array([rng.uniform(-.025, .025), rng.uniform(-.025, .025), 1.0])
        ax /= np.linalg.norm(ax)
        q = [np.cos(a / 2), *(np.sin(a / 2) * ax)]
        for k in range(4):
            datas[e].qpos[3 + k] = q[k]
        L.jbo_forward(C.byref(models[e]), C.byref(datas[e]))
    actions = rng.uniform(-1, 1, size=(n_steps, n_env))
    rewards = np.zeros((n_steps, n_env))
    dp = C.POINTER(C.c_double)
    t0 = time.perf_counter()
    L.jbo_rollout(models, datas, n_env, 0, actions.ctypes.data_as(dp), n_steps, 50, threads,
                  rewards.ctypes.data_as(dp), None)
    dt = time.perf_counter() - t0
    return n_env * n_steps / dt, dt, float(rewards.mean())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_env, n_steps = 16 * cores, 100        # one "step" of this arm = a bounded sample: 16*cores envs x 100 control steps
    for _ in range(max(args.warmup, 0)):
        cpu_rollout_rate(n_env, 2, cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt, _ = cpu_rollout_rate(n_env, n_steps, cores)
        rates.append(r); times.append(dt)
    value = n_env * n_steps * len(times) / sum(times)
    sample = f"{n_env} envs x {n_steps} control steps per step, fp64 C oracle port (oracle/jb_oracle.c), {cores} pthreads"
    emit_json({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{TASK}, uniform random actions, CPU oracle port (MuJoCo/dm_control not installable "
                               f"here: parity unpinned)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })
    return 0


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from jitterbug_b200 import _native
    from jitterbug_b200.batched_env import BatchedJitterbugEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E = args.envs
    extra = {}
    if args.workload == "augmented":
        # BASELINE.json config 4: per-env randomised model parameters (legs + mass, the reference training script's
        # defaults, benchmark.py:129-136).  --models 0 (default): every env gets its own model, drawn and compiled on
        # the device (Philox keyed by global env id).  --models K: a population of K models compiled on the host and
        # assigned by global env id.  Either way each env's model is staged in shared memory by the kernel.
        from jitterbug_b200 import augmented_jitterbug
        if args.models > 0:
            pop = augmented_jitterbug.make_models(args.models, seed=args.seed)
            extra = dict(models=pop, model_index=(np.arange(E, dtype=np.int64) + rank * E) % args.models)
        else:
            t_c = time.perf_counter()
            deltas = augmented_jitterbug.sample_deltas(E, seed=args.seed, device=dev, env_id_offset=rank * E)
            extra = dict(packed_models=augmented_jitterbug.compile_models_device(deltas))
            torch.cuda.synchronize(dev)
            compile_s = time.perf_counter() - t_c
            del deltas
    task = TASK
    if args.workload == "mixed":
        # BASELINE.json config 3: all five tasks in one batch (task = global env id mod 5, observations padded to 19),
        # actions from the device port of the reference's heuristic policies (one extra kernel launch per step)
        from jitterbug_b200.batched_env import TASKS
        task = [TASKS[(rank * E + e) % 5] for e in range(E)]
    env = BatchedJitterbugEnv(task, num_envs=E, device=dev, seed=args.seed, env_id_offset=rank * E, **extra)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_act = 4
    acts = [torch.rand(E, device=dev, generator=g) * 2 - 1 for _ in range(n_act)]
    env.reset()
    stream = torch.cuda.current_stream(dev)
    mixed = args.workload == "mixed"

    def action(i):
        return env.heuristic_actions() if mixed else acts[i % n_act]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # untimed burn-in to the stationary regime (motor spun up, contacts active: BASELINE.md section 4 asks for
    # ~200 control steps before sampling), then the W warm-up steps of the contract
    for i in range(args.burn_in):
        env.step(action(i))
    for i in range(args.warmup):
        env.step(action(i))
    barrier()
    c0 = env.counters()
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        env.step(action(i))
        ev[i + 1].record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    c1 = env.counters()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * E * args.steps / (total_ms_max * 1e-3)

    # ---- e2e: host buffers through the C-ABI host entry point (H2D actions, D2H obs/reward/step_type each step)
    h_pin = [a.cpu().pin_memory() for a in acts]            # inputs in pinned host memory (kept alive by h_pin)
    h_acts = [t.numpy() for t in h_pin]
    e2e_steps = max(2, min(args.steps, 8))
    env.step_host(h_acts[0])
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        env.step_host(h_acts[i % n_act])
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * E * e2e_steps / float(t.item())

    # ---- optional end-of-rollout statistics gather (the only collective; NOT in the step path)
    stats = torch.stack([env._reward.double().sum(), (env._reward.double() ** 2).sum(),
                         torch.tensor(float(E), device=dev, dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    mean_reward = float(stats[0] / stats[2])

    if rank == 0:
        sub = c1["substeps"] - c0["substeps"]
        newton = (c1["newton_iterations"] - c0["newton_iterations"]) / max(sub, 1)
        ncon = (c1["contacts"] - c0["contacts"]) / max(sub, 1)
        fm = _flop_model()
        flops_per_step = 50 * (fm["per_substep"] + fm["per_newton_iteration"] * newton + fm["per_contact_substep"] * ncon)
        ms_kernel = float(np.mean(kernel_ms))
        hbm_peak, which = _measured_peaks()
        achieved_gbs = BYTES_PER_ENV_STEP * E / (ms_kernel * 1e-3) / 1e9
        fp32_peak = C.c_double()
        _native.check(_native.load().jb_measure_fp32_peak(local, C.byref(fp32_peak)))
        achieved_tf = flops_per_step * E / (ms_kernel * 1e-3) / 1e12
        # CPU baseline: bounded sample of the same workload on the host cores (oracle port)
        cores = os.cpu_count() or 1
        cpu = None
        if world >= 1 and not args.no_cpu_baseline:
            n_env, n_steps = 16 * cores, 1000          # one full 1000-step episode per env, ~5-20 s of CPU work
            rate, dt, _ = cpu_rollout_rate(n_env, n_steps, cores)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n_env} envs x {n_steps} control steps ({dt:.1f} s), fp64 C oracle port, {cores} pthreads"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("all five tasks (env id mod 5), device heuristic policies" if mixed else
                                    f"{TASK}, uniform random actions") + f", {E} envs per GPU x {world} GPU, 50 sub-steps + "
                                   f"reward + obs per step, episodes of 1000 steps with in-kernel auto-reset"
                                   + ((f", per-env augmented models (population {args.models}, host-compiled)"
                                       if args.models > 0 else
                                       f", one augmented model per env compiled on the device ({compile_s:.2f} s)")
                                      if args.workload == "augmented" else ""),
                       "envs_per_gpu": E, "task": "mixed" if mixed else TASK, "l2_policy": "inputs larger than L2 "
                       f"({BYTES_PER_ENV_STEP * E / 1e6:.0f} MB state traffic per step vs 126 MB L2)"
                       if BYTES_PER_ENV_STEP * E > 126e6 else "inputs smaller than L2 (compute-bound kernel)",
                       "parallelism": f"env-sharded x{world}, no collective in the step path",
                       "burn_in_steps": args.burn_in,
                       "newton_iters_per_substep": newton, "contacts_per_substep": ncon,
                       "mean_reward": mean_reward},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * E,
                    "d2h_bytes_per_step": 4 * E * (env.obs_dim + 2), "steps": e2e_steps},
            "gpu_launches": args.steps * (2 if mixed else 1),
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak,
                         "traffic": (_traffic_per_env() * E if _traffic_per_env() else None), "peak_source": which,
                         "note": "kernel is FP32-issue bound by design (50 sub-steps fused on chip); see fp32"},
            "fp32": {"achieved": achieved_tf, "peak": fp32_peak.value, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp32_peak.value if fp32_peak.value else None,
                     "flops_per_control_step": flops_per_step, "peak_source": "FFMA-chain microbenchmark, this run"},
            "kernel_ms": ms_kernel,
            "clocks": clocks,
        }
        if cpu:
            out["cpu_baseline"] = cpu
        emit_json(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = 1


def emit_json(obj):
    sys.stdout.flush()
    os.write(_JSON_FD, (json.dumps(obj) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--workload", default="shared", choices=["shared", "augmented", "mixed"],
                    help="shared: one model (headline); augmented: per-env randomised models (config 4); "
                         "mixed: five tasks in one batch driven by the device heuristic policies (config 3)")
    ap.add_argument("--models", type=int, default=0,
                    help="augmented: 0 = one device-compiled model per env; K > 0 = host-compiled population of K")
    ap.add_argument("--burn-in", type=int, default=150, help="untimed control steps before warm-up (steady state)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything else that lands on file descriptor 1 (NCCL's version banner, any
    # library chatter) is sent to stderr, and the line itself goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
