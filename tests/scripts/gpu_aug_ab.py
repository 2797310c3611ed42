"""Tuning helper (GPU): separates the cost of the per-env-model MODE from the cost of augmented model DATA."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from jitterbug_b200 import BatchedJitterbugEnv, augmented_jitterbug as aug, mjcf  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 262144


def run(name, **kw):
    env = BatchedJitterbugEnv("move_from_origin", num_envs=E, seed=0, **kw)
    g = torch.Generator(device="cuda").manual_seed(1)
    acts = [torch.rand(E, device="cuda", generator=g) * 2 - 1 for _ in range(4)]
    env.reset()
    for i in range(150):
        env.step(acts[i % 4])
    torch.cuda.synchronize()
    c0 = list(env.counters().values())
    t0 = time.perf_counter()
    for i in range(10):
        env.step(acts[i % 4])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c1 = list(env.counters().values())
    sub = c1[0] - c0[0]
    print(f"{name:34s} {E * 10 / dt / 1e6:7.3f} M steps/s  newton/sub {(c1[1] - c0[1]) / sub:.3f}  ncon/sub {(c1[2] - c0[2]) / sub:.3f}"
          f"  overflow {c1[3] - c0[3]}  bad {c1[4] - c0[4]}", flush=True)
    env.close()


base = aug.make_aug_base()
zero = aug.compile_models_device(torch.zeros(1, 31, dtype=torch.float64, device="cuda"), base)
run("shared, stock model")
run("per-env mode, identical zero-delta", packed_models=zero.repeat(E, 1).contiguous())
d1 = aug.sample_deltas(1, seed=5)
one = aug.compile_models_device(d1, base)
run("per-env mode, identical augmented", packed_models=one.repeat(E, 1).contiguous())
xml_vals = list(d1.cpu().numpy()[0, 3:30])


class R:
    def __init__(s, v): s.v = list(v)
    def normal(s, a, b): return s.v.pop(0)


run("shared mode, that augmented model", model=mjcf.compile_xml(aug.augment_Jitterbug(modify_legs=True, modify_mass=True, random=R(xml_vals))))
run("per-env mode, all different", packed_models=aug.compile_models_device(aug.sample_deltas(E, seed=0), base))
legs_only = aug.sample_deltas(E, seed=0, modify_mass=False)
run("per-env, legs only", packed_models=aug.compile_models_device(legs_only, base))
mass_only = aug.sample_deltas(E, seed=0, modify_legs=False)
run("per-env, mass only", packed_models=aug.compile_models_device(mass_only, base))
