"""GPU parity tests proper: the CUDA path (through the C-ABI) against the fp64 CPU oracle on identical inputs.

Protocol: the GPU free-runs N envs under random actions to reach diverse contact
states; then, from each env's *identical* pre-step state (float32 state read back, cast to float64), the oracle and
the GPU both advance and are compared.  Tolerances are stated per test.  The oracle is MuJoCo-faithful but unpinned
(see oracle/jb_oracle.h).
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from jitterbug_b200 import mjcf  # noqa: E402
from jitterbug_b200.batched_env import BatchedJitterbugEnv, TASKS, OBS_DIM  # noqa: E402


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _oracle():
    from oracle import oracle as O
    return O


def _free_run(env, steps, seed):
    g = torch.Generator(device=env.device).manual_seed(seed)
    env.reset()
    for _ in range(steps):
        a = torch.rand(env.num_envs, device=env.device, generator=g) * 2 - 1
        env.step(a)
    return g


@pytest.mark.parametrize("N,warm", [(512, 120), (4096, 200)])
def test_single_substep_state_parity(compiled_model, N, warm):
    """One physics sub-step from identical (qpos, qvel, warmstart, ctrl): qpos within 1e-6, qvel within 1e-5 relative
    (fp32 kernel vs fp64 oracle; the north_star tolerance is 1e-5), active contact sets identical.  (4096, 200) is
    BASELINE.json's config 2: 4096 move_from_origin envs, 200 warm-up steps so that contacts are active."""
    _need_gpu()
    O = _oracle()
    env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=1)
    g = _free_run(env, warm, 7)
    qpos, qvel, warm = [t.cpu().numpy().astype(np.float64) for t in env.get_state()]
    # same model, one sub-step per control step
    env1 = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=1, control_timestep=0.0002, time_limit=float("inf"))
    env1.reset()
    env1.set_state(qpos, qvel, warm, step_count=np.ones(N))
    a = (torch.rand(N, device=env.device, generator=g) * 2 - 1)
    ncon, geom, dist = [t.cpu().numpy() for t in env1.contacts()]
    env1.step(a)
    q1, v1, w1 = [t.cpu().numpy().astype(np.float64) for t in env1.get_state()]
    a = a.cpu().numpy()
    sim = O.OracleSim(compiled_model)
    nbad_con = 0
    eq, ev, ev_tight = [], [], []
    skipped = 0
    for e in range(N):
        sim.set_state(qpos[e], qvel[e], warm[e], ctrl=a[e])
        sim.step1()
        g2, d2, ex, _ = sim.contacts()
        act = [(int(gg), float(dd)) for gg, dd, xx in zip(g2, d2, ex) if not xx]
        mine = [(int(geom[e, i]), float(dist[e, i])) for i in range(ncon[e])]
        # knife-edge contacts (|dist| < 1e-6 m) may legitimately flip in fp32: skip those envs (SURVEY hard part 3)
        edge = any(abs(dd) < 1e-6 for _, dd in act) or any(abs(dd) < 1e-6 for _, dd in mine) or \
            any(0 <= dd < 1e-6 for dd in d2)
        if edge:
            skipped += 1
            continue
        if [gg for gg, _ in act] != [gg for gg, _ in mine]:
            nbad_con += 1
            continue
        sim.step2()
        dq = np.abs(q1[e] - sim.qpos)
        dv = np.abs(v1[e] - sim.qvel)
        eq.append(float((dq / np.maximum(np.abs(sim.qpos), 1.0)).max()))

        def vscale(lin_floor, ang_floor):
            # velocity error relative to the scale of the velocity block it belongs to (linear m/s; angular, hinge and
            # motor rad/s), floored: below the floor the comparison is absolute
            return np.concatenate([np.full(3, max(lin_floor, np.abs(sim.qvel[0:3]).max())),
                                   np.full(3, max(ang_floor, np.abs(sim.qvel[3:6]).max())),
                                   np.full(8, max(ang_floor, np.abs(sim.qvel[6:14]).max())),
                                   np.full(1, max(ang_floor, abs(sim.qvel[14])))])
        ev.append(float((dv / vscale(1e-2, 1.0)).max()))
        ev_tight.append(float((dv / vscale(1e-3, 0.1)).max()))
    eq, ev, ev_tight = np.array(eq), np.array(ev), np.array(ev_tight)
    print(f"single sub-step, {N} envs: skipped as knife-edge {skipped} ({100 * skipped / N:.2f} %); scaled qpos error max "
          f"{eq.max():.2e} p99 {np.quantile(eq, .99):.2e}; scaled qvel error max {ev.max():.2e} p99 {np.quantile(ev, .99):.2e} "
          f"(floors 1e-2 m/s, 1 rad/s); with floors 1e-3 m/s, 0.1 rad/s: max {ev_tight.max():.2e} p99 {np.quantile(ev_tight, .99):.2e}")
    assert skipped <= 0.02 * N, skipped     # measured: well under 1 % of envs have a contact within 1 um of the floor
    assert nbad_con == 0, f"{nbad_con} envs with a different active contact set"
    assert eq.max() < 1e-6, eq.max()        # north_star asks 1e-5 relative; measured 1.1e-7 on B200
    assert ev.max() < 1e-5, ev.max()        # measured 7.5e-7
    # a sub-step changes a velocity by h * acceleration (~0.2 - 2 rad/s for a sprung leg), so below ~0.1 rad/s the error
    # is absolute (Newton tolerance, fp32 rounding of h * a), not relative: with floors ten times lower the bulk is
    # unchanged (p99 ~4e-7) and the worst env of 4096 stays within a few 1e-5 (measured 5e-6 ... 1.4e-5 across builds)
    assert np.quantile(ev_tight, 0.99) < 2e-6, np.quantile(ev_tight, 0.99)
    assert ev_tight.max() < 5e-5, ev_tight.max()


def test_many_contacts_per_lane_on_device(compiled_model):
    """States with the core pressed into the floor (>= 12 contacts, up to 6 in the lanes that own the core boxes): the
    lanes that own the core boxes fill 6 of their 8 contact slots.  Active sets must equal the oracle's and one sub-step
    must agree (forces are huge here: velocities compared at 1e-4 of their scale)."""
    _need_gpu()
    O = _oracle()
    N = 96
    rng = np.random.default_rng(5)
    qpos = np.tile(compiled_model.qpos0, (N, 1))
    qpos[:, 2] = 0.021 - 0.004 * rng.random(N)
    ang = 0.02 * rng.standard_normal((N, 3))
    qpos[:, 3] = 1.0
    qpos[:, 4:7] = 0.5 * ang
    qpos[:, 3:7] /= np.linalg.norm(qpos[:, 3:7], axis=1, keepdims=True)
    qvel = 0.05 * rng.standard_normal((N, 15))
    env1 = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=1, control_timestep=0.0002, time_limit=float("inf"))
    env1.reset()
    env1.set_state(qpos, qvel, np.zeros((N, 15)), step_count=np.ones(N))
    qpos, qvel, warm = [t.cpu().numpy().astype(np.float64) for t in env1.get_state()]
    ncon, geom, dist = [t.cpu().numpy() for t in env1.contacts(32)]
    a = torch.full((N,), 0.3, device=env1.device)
    env1.step(a)
    q1, v1, _ = [t.cpu().numpy().astype(np.float64) for t in env1.get_state()]
    assert env1.counters()["contact_overflows"] == 0
    sim = O.OracleSim(compiled_model)
    checked = 0
    for e in range(N):
        sim.set_state(qpos[e], qvel[e], warm[e], ctrl=0.3)
        sim.step1()
        g2, d2, ex, _ = sim.contacts()
        if any(abs(dd) < 1e-6 for dd in d2):
            continue
        act = [int(gg) for gg, xx in zip(g2, ex) if not xx]
        assert [int(geom[e, i]) for i in range(ncon[e])] == act
        assert ncon[e] >= 12
        sim.step2()
        np.testing.assert_allclose(q1[e], sim.qpos, atol=2e-6)
        scale = max(1.0, np.abs(sim.qvel).max())
        assert np.abs(v1[e] - sim.qvel).max() < 1e-4 * scale
        checked += 1
    assert checked > N // 2


def test_control_step_parity_and_reward(compiled_model):
    """One control step (50 sub-steps) from identical states; reward within 1e-4, obs within 2e-3 (chaotic contact)."""
    _need_gpu()
    O = _oracle()
    N = 256
    for task in TASKS:
        env = BatchedJitterbugEnv(task, num_envs=N, seed=3)
        g = _free_run(env, 60, 11)
        qpos, qvel, warm = [t.cpu().numpy().astype(np.float64) for t in env.get_state()]
        tgt = env.get_target().cpu().numpy().astype(np.float64)
        a = (torch.rand(N, device=env.device, generator=g) * 2 - 1)
        obs, rew, st = env.step(a)
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        q1, v1, _ = [t.cpu().numpy().astype(np.float64) for t in env.get_state()]
        a = a.cpu().numpy()
        sim = O.OracleSim(compiled_model)
        tk = O.TaskOracle(sim, task)
        err_obs, err_rew = [], []
        for e in range(N):
            sim.set_state(qpos[e], qvel[e], warm[e])
            yaw = 2 * np.arctan2(tgt[e, 3], tgt[e, 2])
            sim.set_target(tgt[e, 0], tgt[e, 1], yaw)
            sim.step1()
            sim.control_step(a[e], 50)
            o_ref = np.concatenate([np.ravel(v) for v in tk.get_observation().values()])
            assert o_ref.shape[0] == OBS_DIM[task]
            err_obs.append(np.abs(obs[e, :OBS_DIM[task]] - o_ref).max())
            err_rew.append(abs(rew[e] - tk.get_reward()))
        err_obs, err_rew = np.array(err_obs), np.array(err_rew)
        # a few envs sit on a contact make/break edge inside the 50 sub-steps and diverge; the bulk must be tight
        assert np.median(err_obs) < 2e-5, (task, np.median(err_obs))
        assert np.quantile(err_obs, 0.95) < 2e-3, (task, np.quantile(err_obs, 0.95))
        assert np.quantile(err_rew, 0.95) < 1e-4, (task, np.quantile(err_rew, 0.95))


def test_obs_reward_from_identical_state(compiled_model):
    """Task kernels alone: obs / reward computed from an identical state (set_state -> reset-free read-back)."""
    _need_gpu()
    O = _oracle()
    N = 128
    for task in TASKS:
        env = BatchedJitterbugEnv(task, num_envs=N, seed=5, control_timestep=0.0002, time_limit=float("inf"))
        g = _free_run(env, 200, 13)
        a = torch.zeros(N, device=env.device)
        obs, rew, _ = env.step(a)
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        qpos, qvel, warm = [t.cpu().numpy().astype(np.float64) for t in env.get_state()]
        tgt = env.get_target().cpu().numpy().astype(np.float64)
        sim = O.OracleSim(compiled_model)
        tk = O.TaskOracle(sim, task)
        for e in range(N):
            sim.set_state(qpos[e], qvel[e], warm[e])
            sim.set_target(tgt[e, 0], tgt[e, 1], 2 * np.arctan2(tgt[e, 3], tgt[e, 2]))
            sim.step1()
            o_ref = np.concatenate([np.ravel(v) for v in tk.get_observation().values()])
            np.testing.assert_allclose(obs[e, :OBS_DIM[task]], o_ref, atol=2e-6, rtol=0)
            assert abs(rew[e] - tk.get_reward()) < 1e-5
            o_c, r_c = sim.obs_reward_c(TASKS.index(task))
            np.testing.assert_allclose(o_c, o_ref, atol=1e-12)
            assert abs(r_c - tk.get_reward()) < 1e-12


def test_episode_semantics():
    """FIRST/MID/LAST bookkeeping (control.Environment): 1000 steps, LAST at the limit, next step() is a reset."""
    _need_gpu()
    env = BatchedJitterbugEnv("move_from_origin", num_envs=64, seed=0, time_limit=0.05)   # 5 control steps
    a = torch.zeros(64, device=env.device)
    obs, rew, st = env.step(a)        # step before reset == reset
    assert (st == 0).all() and (rew == 0).all()
    for k in range(4):
        obs, rew, st = env.step(a)
        assert (st == 1).all()
    obs, rew, st = env.step(a)
    assert (st == 2).all()
    obs, rew, st = env.step(a)
    assert (st == 0).all()
    assert int(env.get_step_count().max()) == 0


def test_heuristic_kernel_matches_reference_policy(compiled_model):
    _need_gpu()
    O = _oracle()
    N = 256
    for task in TASKS:
        env = BatchedJitterbugEnv(task, num_envs=N, seed=9)
        _free_run(env, 30, 17)
        obs = env._obs.clone()
        act = env.heuristic_actions(obs).cpu().numpy()
        lay = env.observation_layout()
        obs = obs.cpu().numpy().astype(np.float64)
        for e in range(N):
            od = {k: obs[e, s] for k, s in lay.items()}
            ref = O.heuristic(task, od)
            assert abs(act[e] - ref) < 1e-5, (task, e, act[e], ref)


def test_long_horizon_return_distributions(compiled_model):
    """Beyond a few control steps trajectories are chaotic, so only distributions can agree (north_star): returns of
    the fp32 CUDA path and of the fp64 oracle from the same initial states under (a) the same open-loop random action
    sequences and (b) the closed-loop move_from_origin heuristic must be statistically indistinguishable."""
    _need_gpu()
    O = _oracle()
    import ctypes as C
    from scipy import stats
    N, T = 192, 120
    env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=21)
    env.reset()
    q0, v0, w0 = [t.cpu().numpy().astype(np.float64) for t in env.get_state()]
    g = torch.Generator(device=env.device).manual_seed(3)
    acts = torch.rand(T, N, device=env.device, generator=g) * 2 - 1
    ret_gpu = torch.zeros(N, device=env.device)
    for k in range(T):
        _, r, _ = env.step(acts[k])
        ret_gpu += r
    ret_gpu = ret_gpu.cpu().numpy()
    # oracle: same initial states, same action sequences, C rollout on all host threads
    L = O.lib()
    models = (O.JboModel * N)()
    datas = (O.JboData * N)()
    proto = O.make_model(compiled_model)
    for e in range(N):
        C.memmove(C.byref(models[e]), C.byref(proto), C.sizeof(O.JboModel))
        L.jbo_reset_data(C.byref(models[e]), C.byref(datas[e]))
        for i in range(16):
            datas[e].qpos[i] = q0[e, i]
        L.jbo_forward(C.byref(models[e]), C.byref(datas[e]))
    a_np = np.ascontiguousarray(acts.cpu().numpy().astype(np.float64))
    rew = np.zeros((T, N))
    dp = C.POINTER(C.c_double)
    import os
    L.jbo_rollout(models, datas, N, 0, a_np.ctypes.data_as(dp), T, 50, os.cpu_count() or 1, rew.ctypes.data_as(dp), None)
    ret_cpu = rew.sum(axis=0)
    # (a) random policy: same distribution of returns
    ks = stats.ks_2samp(ret_gpu, ret_cpu)
    assert ks.pvalue > 0.01, (ks, ret_gpu.mean(), ret_cpu.mean())
    se = np.sqrt(ret_gpu.var() / N + ret_cpu.var() / N)
    assert abs(ret_gpu.mean() - ret_cpu.mean()) < 4 * se + 1e-3, (ret_gpu.mean(), ret_cpu.mean(), se)
    # early steps are not yet chaotic: per-env rewards of the first control steps agree closely
    env2 = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=21)
    env2.reset()
    for k in range(3):
        _, r, _ = env2.step(acts[k])
        assert np.quantile(np.abs(r.cpu().numpy() - rew[k]), 0.9) < 1e-4
    # (b) closed-loop heuristic: device heuristic on the GPU vs reference-policy port on the oracle
    M = 48
    envh = BatchedJitterbugEnv("move_from_origin", num_envs=M, seed=33)
    obs = envh.reset()
    qh = envh.get_state()[0].cpu().numpy().astype(np.float64)
    rg = torch.zeros(M, device=envh.device)
    for k in range(T):
        obs, r, _ = envh.step(envh.heuristic_actions(obs))
        rg += r
    rg = rg.cpu().numpy()
    rc = np.zeros(M)
    for e in range(M):
        sim = O.OracleSim(compiled_model)
        sim.qpos[:] = qh[e]
        sim.forward()
        tk = O.TaskOracle(sim, "move_from_origin")
        for k in range(T):
            sim.control_step(O.heuristic("move_from_origin", tk.get_observation()))
            rc[e] += tk.get_reward()
    ksh = stats.ks_2samp(rg, rc)
    assert ksh.pvalue > 0.01, (ksh, rg.mean(), rc.mean())


def test_per_env_kernel_single_substep_parity_including_spread_contacts(compiled_model):
    """The per-env-model instantiation of the step kernel (which spreads a heavy lane's contacts over its quad) against
    the oracle, one sub-step from identical contact-rich states; every env carries the stock model, so the states and
    the oracle are the shared-model ones.  The states with a leg standing on its lower-cylinder cap (>= 4 contacts on
    one leg, the case the spread step exists for) must be present and must agree like the others."""
    _need_gpu()
    O = _oracle()
    N = 2048
    env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=1)
    g = _free_run(env, 200, 7)
    qpos, qvel, warm = [t.cpu().numpy().astype(np.float64) for t in env.get_state()]
    # N copies of the packed stock model (test-side pack helper: the same jb_pack_model the library uses)
    from jitterbug_b200.model import fill_struct, flatten_model
    import hostsim
    stock = hostsim.pack(fill_struct(flatten_model(compiled_model)), double=False)
    pk = torch.from_numpy(np.tile(stock, (N, 1))).to(env.device).contiguous()
    env1 = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=1, control_timestep=0.0002, time_limit=float("inf"),
                               packed_models=pk)
    env1.reset()
    env1.set_state(qpos, qvel, warm, step_count=np.ones(N))
    a = (torch.rand(N, device=env.device, generator=g) * 2 - 1)
    ncon, geom, dist = [t.cpu().numpy() for t in env1.contacts(32)]
    env1.step(a)
    q1, v1, _ = [t.cpu().numpy().astype(np.float64) for t in env1.get_state()]
    a = a.cpu().numpy()
    sim = O.OracleSim(compiled_model)
    worst_q = worst_v = 0.0
    heavy = checked = 0
    leg_of = {g_: (g_ - 5) // 4 for g_ in range(5, 21)}
    for e in range(N):
        sim.set_state(qpos[e], qvel[e], warm[e], ctrl=a[e])
        sim.step1()
        g2, d2, ex, _ = sim.contacts()
        if any(abs(dd) < 1e-6 for dd in d2):
            continue
        act = [int(gg) for gg, xx in zip(g2, ex) if not xx]
        assert [int(geom[e, i]) for i in range(ncon[e])] == act
        per_leg = np.bincount([leg_of[gg] for gg in act if gg in leg_of], minlength=4)
        is_heavy = per_leg.max() >= 4
        sim.step2()
        dq = (np.abs(q1[e] - sim.qpos) / np.maximum(np.abs(sim.qpos), 1.0)).max()
        scale = np.concatenate([np.full(3, max(1e-2, np.abs(sim.qvel[0:3]).max())), np.full(3, max(1.0, np.abs(sim.qvel[3:6]).max())),
                                np.full(8, max(1.0, np.abs(sim.qvel[6:14]).max())), np.full(1, max(1.0, abs(sim.qvel[14])))])
        dv = (np.abs(v1[e] - sim.qvel) / scale).max()
        worst_q, worst_v = max(worst_q, dq), max(worst_v, dv)
        heavy += int(is_heavy)
        checked += 1
    print(f"per-env kernel: {checked} envs checked, {heavy} with a 4-contact leg, worst qpos {worst_q:.2e}, qvel {worst_v:.2e}")
    assert checked > N * 0.9 and heavy >= 10
    assert worst_q < 1e-6 and worst_v < 1e-5


def test_mixed_batch_config3_control_step_parity(compiled_model):
    """BASELINE.json config 3: all five tasks in one batch of 65,536 envs (task = env id mod 5, observations padded to
    19), actions from the device heuristic-policy kernel.  After 100 free-running control steps, one more control step
    from identical states is checked against the oracle for a stratified sample of every task: contact set at the
    pre-step state, observation, reward, and the heuristic action itself."""
    _need_gpu()
    O = _oracle()
    N, PER_TASK = 65536, 64
    names = [TASKS[i % 5] for i in range(N)]
    env = BatchedJitterbugEnv(names, num_envs=N, seed=5)
    obs = env.reset()
    for _ in range(100):
        obs, _, _ = env.step(env.heuristic_actions(obs))
    qpos, qvel, warm = [t.cpu().numpy().astype(np.float64) for t in env.get_state()]
    tgt = env.get_target().cpu().numpy().astype(np.float64)
    ncon, geom, dist = [t.cpu().numpy() for t in env.contacts(32)]
    obs_pre = obs.clone()
    act = env.heuristic_actions(obs_pre)
    obs1, rew1, _ = env.step(act)
    obs_pre, act, obs1, rew1 = obs_pre.cpu().numpy(), act.cpu().numpy(), obs1.cpu().numpy(), rew1.cpu().numpy()
    assert env.counters()["diverged_resets"] == 0
    rng = np.random.default_rng(0)
    sim = O.OracleSim(compiled_model)
    for ti, task in enumerate(TASKS):
        ids = rng.choice(np.arange(ti, N, 5), PER_TASK, replace=False)
        tk = O.TaskOracle(sim, task)
        D = OBS_DIM[task]
        err_o, err_r, bad_sets, edge = [], [], 0, 0
        for e in ids:
            assert np.all(obs_pre[e, D:] == 0) and np.all(obs1[e, D:] == 0)         # padding
            sim.set_state(qpos[e], qvel[e], warm[e])
            sim.set_target(tgt[e, 0], tgt[e, 1], 2 * np.arctan2(tgt[e, 3], tgt[e, 2]))
            sim.step1()
            g2, d2, ex, _ = sim.contacts()
            if any(abs(dd) < 1e-6 for dd in d2):
                edge += 1
            elif [int(gg) for gg, xx in zip(g2, ex) if not xx] != [int(geom[e, i]) for i in range(ncon[e])]:
                bad_sets += 1
            o0 = np.concatenate([np.ravel(v) for v in tk.get_observation().values()])
            np.testing.assert_allclose(obs_pre[e, :D], o0, rtol=0, atol=4e-6)       # the state the policy saw
            lay = env.observation_layout(task)
            a_ref = O.heuristic(task, {k: obs_pre[e, sl].astype(np.float64) for k, sl in lay.items()})
            assert abs(float(np.ravel(a_ref)[0]) - act[e]) < 1e-5
            sim.control_step(act[e], 50)
            o_ref = np.concatenate([np.ravel(v) for v in tk.get_observation().values()])
            err_o.append(np.abs(obs1[e, :D] - o_ref).max())
            err_r.append(abs(rew1[e] - tk.get_reward()))
        err_o, err_r = np.array(err_o), np.array(err_r)
        print(f"config 3, {task}: {PER_TASK} envs, obs error median {np.median(err_o):.1e} p95 {np.quantile(err_o, .95):.1e}, "
              f"reward error p95 {np.quantile(err_r, .95):.1e}, knife-edge {edge}")
        assert bad_sets == 0
        assert np.median(err_o) < 2e-5, (task, np.median(err_o))
        assert np.quantile(err_o, 0.95) < 2e-3, (task, np.quantile(err_o, 0.95))
        assert np.quantile(err_r, 0.95) < 1e-4, (task, np.quantile(err_r, 0.95))


@pytest.mark.parametrize("task", TASKS)
def test_evaluate_policy_protocol_return_distributions(compiled_model, task):
    """The reference's evaluation protocol (benchmarks/evaluate_policy.py:10-35: reset, then int(_step_limit - 1) = 999
    control steps, reward per step) for every task under (a) its heuristic policy and (b) a uniform random policy:
    per-episode returns of the CUDA path and of the fp64 oracle started from the same reset states must be statistically
    indistinguishable (two-sample KS and a mean test).  Trajectories are chaotic, so nothing finer can agree."""
    _need_gpu()
    O = _oracle()
    from scipy import stats
    T, M = 999, 128
    ti = TASKS.index(task)
    for policy in ("heuristic", "random"):
        env = BatchedJitterbugEnv(task, num_envs=M, seed=41 + ti)
        obs = env.reset()
        q0 = env.get_state()[0].cpu().numpy().astype(np.float64)
        tgt = env.get_target().cpu().numpy().astype(np.float64)
        g = torch.Generator(device=env.device).manual_seed(7)
        acts = torch.rand(T, M, device=env.device, generator=g) * 2 - 1 if policy == "random" else None
        ret = torch.zeros(M, device=env.device)
        for k in range(T):
            obs, r, st = env.step(acts[k] if acts is not None else env.heuristic_actions(obs))
            ret += r
            assert int(st.max()) == 1                                  # 999 steps never reach LAST (App. C11)
        ret_gpu = ret.cpu().numpy().astype(np.float64)
        rew = O.batch_rollout(compiled_model, ti, q0, tgt, T, actions=None if acts is None else acts.cpu().numpy())
        ret_cpu = rew.sum(axis=0)
        ks = stats.ks_2samp(ret_gpu, ret_cpu)
        se = np.sqrt(ret_gpu.var() / M + ret_cpu.var() / M)
        print(f"{task} / {policy}: return GPU {ret_gpu.mean():.1f} +- {ret_gpu.std():.1f}, oracle {ret_cpu.mean():.1f} +- "
              f"{ret_cpu.std():.1f}, KS p = {ks.pvalue:.3f}")
        assert ks.pvalue > 0.005, (task, policy, ks, ret_gpu.mean(), ret_cpu.mean())
        assert abs(ret_gpu.mean() - ret_cpu.mean()) < 4 * se + 1e-2 * T * 1e-2, (task, policy, ret_gpu.mean(), ret_cpu.mean(), se)
        assert env.counters()["diverged_resets"] == 0


def test_diverged_env_restarts_with_first():
    """mj_checkPos / mj_checkVel: a state that blows up (here a NaN planted through set_state) restarts the env; the
    caller sees the boundary as step type FIRST with reward 0 and a fresh episode's observation."""
    _need_gpu()
    from jitterbug_b200.rollout import VecJitterbugEnv
    N = 64
    env = BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=2)
    env.reset()
    a = torch.zeros(N, device=env.device)
    for _ in range(3):
        env.step(a)
    qpos, qvel, warm = env.get_state()
    qvel[5, 2] = float("nan")
    qpos[9, 0] = 1e12
    env.set_state(qpos, qvel, warm)
    obs, rew, st = env.step(a)
    st = st.cpu().numpy()
    assert st[5] == 0 and st[9] == 0 and (np.delete(st, [5, 9]) == 1).all()
    assert float(rew[5]) == 0.0 and float(rew[9]) == 0.0
    assert torch.isfinite(obs).all()
    assert env.counters()["diverged_resets"] == 2
    assert env.get_step_count().cpu().numpy()[[5, 9]].tolist() == [0, 0]
    q2 = env.get_state()[0].cpu().numpy()
    np.testing.assert_allclose(q2[[5, 9], :3], [[0, 0, 0.035]] * 2, atol=1e-6)     # reset pose
    obs, rew, st = env.step(a)
    assert (st == 1).all()
    # the VecEnv facade reports the boundary as `done`
    vec = VecJitterbugEnv(BatchedJitterbugEnv("move_from_origin", num_envs=N, seed=2))
    vec.reset()
    vec.step(a)
    qpos, qvel, warm = vec.env.get_state()
    qvel[7, 0] = float("nan")
    vec.env.set_state(qpos, qvel, warm)
    _, _, done, _ = vec.step(a)
    assert bool(done[7]) and int(done.sum()) == 1


def test_two_simulators_on_two_devices():
    """Every C-ABI entry point makes its simulator's device current: two envs on two GPUs in one process."""
    _need_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    e0 = BatchedJitterbugEnv("move_from_origin", num_envs=256, seed=1, device="cuda:0")
    e1 = BatchedJitterbugEnv("move_from_origin", num_envs=256, seed=1, device="cuda:1")
    torch.cuda.set_device(0)
    e0.reset(); e1.reset()
    for k in range(5):
        a = torch.full((256,), 0.1 * k)
        o0, r0, _ = e0.step(a.to("cuda:0"))
        o1, r1, _ = e1.step(a.to("cuda:1"))
        assert torch.cuda.current_device() == 0
    assert torch.equal(o0.cpu(), o1.cpu()) and torch.equal(r0.cpu(), r1.cpu())
    assert e0.counters() == e1.counters()
