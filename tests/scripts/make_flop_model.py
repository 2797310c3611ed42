"""Instrumented algorithmic FLOP model of the QUAD sub-step (tests/hostsim/quadflops.cpp) -> profiles/flop_model.json.

    python tests/scripts/make_flop_model.py [n_envs] [control_steps]

Random-policy move_from_origin trajectories (reset distribution of the task); every control step is run through the
counting instantiation of csrc/jb_quad.cuh (which also advances the state), and a linear model
flops = a * sub-steps + b * Newton trips + c * contact-sub-steps is fitted for bench.py."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import hostsim  # noqa: E402
from jitterbug_b200 import mjcf, model as jbm  # noqa: E402

HS = os.path.join(ROOT, "tests", "hostsim")
so = os.path.join(HS, "libquadflops.so")
subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-ffp-contract=off", "-o", so,
                       os.path.join(HS, "quadflops.cpp")])
L = C.CDLL(so)
n_env = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_ctl = int(sys.argv[2]) if len(sys.argv) > 2 else 130
cm = mjcf.compile_xml()
P = hostsim.pack(jbm.fill_struct(jbm.flatten_model(cm), double=True), True)
rng = np.random.RandomState(0)
rows = []
dp = C.POINTER(C.c_double)
for e in range(n_env):
    qpos = cm.qpos0.copy()
    a = rng.uniform(0, 2 * np.pi)
    ax = np.array([rng.uniform(-.025, .025), rng.uniform(-.025, .025), 1.0]); ax /= np.linalg.norm(ax)
    qpos[3:7] = [np.cos(a / 2), *(np.sin(a / 2) * ax)]
    qvel, warm = np.zeros(15), np.zeros(15)
    for k in range(n_ctl):
        alg, exe, it, nc = C.c_longlong(), C.c_longlong(), C.c_int(), C.c_int()
        L.hq_count_flops(P.ctypes.data_as(dp), qpos.ctypes.data_as(dp), qvel.ctypes.data_as(dp), warm.ctypes.data_as(dp),
                         C.c_double(rng.uniform(-1, 1)), 50, C.byref(alg), C.byref(exe), C.byref(it), C.byref(nc))
        rows.append((alg.value, exe.value, it.value, nc.value))
rows = np.array(rows, dtype=np.float64)
A = np.stack([np.full(len(rows), 50.0), rows[:, 2], rows[:, 3]], axis=1)
fit, *_ = np.linalg.lstsq(A, rows[:, 0], rcond=None)
fit_exe, *_ = np.linalg.lstsq(A, rows[:, 1], rcond=None)
pred = A @ fit
out = {
    "description": "Algorithmic FLOPs of the fused step kernel's algorithm (the QUAD sub-step, csrc/jb_quad.cuh), counted by "
                   "instantiating it with a counting scalar in the fiber emulation (tests/hostsim/quadflops.cpp): "
                   "+,-,*,/,compare,sqrt,abs = 1 flop, FMA = 2, sincos = 2; work the four lanes replicate is counted once "
                   "(per-phase rule in that file), 'executed' counts it four times. Random-policy move_from_origin "
                   f"trajectories, {n_env} envs x {n_ctl} control steps.",
    "flops_per_control_step_mean": float(rows[:, 0].mean()),
    "executed_flops_per_control_step_mean": float(rows[:, 1].mean()),
    "newton_iters_per_substep": float(rows[:, 2].mean() / 50),
    "contacts_per_substep": float(rows[:, 3].mean() / 50),
    "fit": {"per_substep": float(fit[0]), "per_newton_iteration": float(fit[1]), "per_contact_substep": float(fit[2])},
    "fit_executed": {"per_substep": float(fit_exe[0]), "per_newton_iteration": float(fit_exe[1]),
                     "per_contact_substep": float(fit_exe[2])},
    "fit_relative_rms_error": float(np.sqrt(np.mean((pred - rows[:, 0]) ** 2)) / rows[:, 0].mean()),
}
json.dump(out, open(os.path.join(ROOT, "profiles", "flop_model.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
