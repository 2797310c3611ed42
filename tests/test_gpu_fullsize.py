"""BASELINE.json's full sizes (1 Mi envs per GPU; 262144 augmented envs) through size-independent properties: an env's
trajectory must not depend on how many other envs are in the batch, on the CTA size the launcher picks for that batch
size, or on what the other quads of its warp / CTA are doing (finished quads take zero-length steps, which must be
bitwise no-ops) -- so the first envs of the full-size batch are compared bit for bit with a small batch that the parity
tests check against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from jitterbug_b200 import BatchedJitterbugEnv  # noqa: E402
from jitterbug_b200 import augmented_jitterbug as aug  # noqa: E402


def _actions(n, step, device):
    e = torch.arange(n, device=device, dtype=torch.float32)
    return torch.sin(0.37 * e + 1.3 * step)


def _run(env, steps):
    env.reset()
    out = []
    for k in range(steps):
        obs, rew, st = env.step(_actions(env.num_envs, k, env.device))
        out.append((obs.clone(), rew.clone(), st.clone()))
    return out


@pytest.mark.parametrize("task", ["move_from_origin", "move_to_pose"])
def test_one_mi_envs_equal_small_batch_bit_for_bit(task):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    big = BatchedJitterbugEnv(task, num_envs=1 << 20, seed=3)
    small = BatchedJitterbugEnv(task, num_envs=4096, seed=3)
    steps = 12
    a, b = _run(big, steps), _run(small, steps)
    for (ob, rb, sb), (os_, rs, ss) in zip(a, b):
        assert torch.equal(ob[:4096], os_) and torch.equal(rb[:4096], rs) and torch.equal(sb[:4096], ss)
        assert torch.isfinite(ob).all() and float(rb.min()) >= 0.0 and float(rb.max()) <= 1.0
    # the far end of the batch as well: same global env ids through env_id_offset
    tail = BatchedJitterbugEnv(task, num_envs=4096, seed=3, env_id_offset=(1 << 20) - 4096)
    tail.reset()
    e = torch.arange((1 << 20) - 4096, 1 << 20, device=tail.device, dtype=torch.float32)
    for k in range(steps):
        ot, rt, _ = tail.step(torch.sin(0.37 * e + 1.3 * k))
        assert torch.equal(a[k][0][-4096:], ot) and torch.equal(a[k][1][-4096:], rt)
    c = big.counters()
    assert c["diverged_resets"] == 0 and c["substeps"] == steps * 50 * (1 << 20)
    # a reset starts the next episode (Philox streams are keyed by the episode index): new poses, not a replay
    a2 = _run(big, 1)
    assert not torch.equal(a[0][0], a2[0][0])


def test_augmented_full_size_equals_small_batch_bit_for_bit():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = 262144
    big = aug.make_batched_env("move_in_direction", num_envs=n, seed=5)
    small = aug.make_batched_env("move_in_direction", num_envs=2048, seed=5)
    a, b = _run(big, 8), _run(small, 8)
    for (ob, rb, _), (os_, rs, _) in zip(a, b):
        assert torch.equal(ob[:2048], os_) and torch.equal(rb[:2048], rs)
        assert torch.isfinite(ob).all()
    # models differ from env to env: so do the trajectories of envs that start from the same pose statistics
    assert float((a[-1][0][:1024, 2] - a[-1][0][1024:2048, 2]).abs().max()) > 0


def test_results_do_not_depend_on_slot_order_or_segmentation():
    """The load-sorted slot order and the 10-sub-step segmentation of a control step are scheduling only: identity order
    with one launch per step, the automatic sort (five segments, re-sorted after each) and random explicit permutations
    give bit-identical observations, rewards, step types, states and device counters.  (The 1 Mi-env tests above compare
    the default sorted path with small identity-ordered batches as well.)"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    n = 16384
    envs = [BatchedJitterbugEnv("move_to_pose", num_envs=n, seed=9, time_limit=0.3) for _ in range(3)]   # 30-step episodes
    envs[0].set_slot_order("identity")
    envs[1].set_slot_order("auto")
    g = torch.Generator(device=envs[0].device).manual_seed(2)
    envs[2].set_slot_order(torch.randperm(n, device=envs[0].device, generator=g).int())
    for e in envs:
        e.reset()
    for k in range(70):                          # crosses two episode boundaries (LAST, then the reset step)
        a = _actions(n, k, envs[0].device)
        outs = [e.step(a) for e in envs]
        for o in outs[1:]:
            for x, y in zip(outs[0], o):
                assert torch.equal(x, y), k
        if k % 9 == 0:
            envs[2].set_slot_order(torch.randperm(n, device=envs[0].device, generator=g).int())
    for s0, s1, s2 in zip(*[e.get_state() for e in envs]):
        assert torch.equal(s0, s1) and torch.equal(s0, s2)
    assert envs[0].counters() == envs[1].counters() == envs[2].counters()
    perm = envs[1].get_slot_order().long()
    assert torch.equal(torch.sort(perm).values, torch.arange(n, device=perm.device))
    # the sort really orders by load: the load key of the envs is non-increasing along the slots (heaviest first)
    ld = envs[1].load().long()
    last = (ld[:, 3] >> 16)[perm]
    assert int(last[: n // 20].float().mean() * 100) >= int(last[-n // 20:].float().mean() * 100)
    assert envs[0].launch_count() < envs[1].launch_count()


def test_ragged_batch_sizes_equal_the_same_envs_of_a_bigger_batch():
    """Batches that do not fill a quad-warp, a CTA or a wave (1, 3, 97, 385 envs), and batches on either side of the size
    from which the step is cut into load-sorted segments, give every env the trajectory it has in a bigger batch, bit for
    bit; empty batches and wrong-length action tensors are rejected."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ref = BatchedJitterbugEnv("move_to_pose", num_envs=2048, seed=13, time_limit=0.2)      # 20-step episodes
    want = _run(ref, 45)                                                                  # crosses two episode ends
    for n in (1, 3, 97, 385):
        env = BatchedJitterbugEnv("move_to_pose", num_envs=n, seed=13, time_limit=0.2)
        env.reset()
        for k in range(45):
            o, r, st = env.step(_actions(2048, k, env.device)[:n])
            assert torch.equal(o, want[k][0][:n]) and torch.equal(r, want[k][1][:n]) and torch.equal(st, want[k][2][:n]), (n, k)
    # around the auto-sort threshold: identity order below it, sorted segments from it on
    import ctypes as C
    edge = 4 * 148 * 96
    big = BatchedJitterbugEnv("move_from_origin", num_envs=edge + 5, seed=21)
    a = _run(big, 6)
    for n in (edge - 1, edge):
        env = BatchedJitterbugEnv("move_from_origin", num_envs=n, seed=21)
        env.reset()
        for k in range(6):
            o, r, st = env.step(_actions(edge + 5, k, env.device)[:n])
            assert torch.equal(o, a[k][0][:n]) and torch.equal(r, a[k][1][:n]), (n, k)
    assert big.launch_count() > 6 * 5                                    # segmented and sorted ...
    small = BatchedJitterbugEnv("move_from_origin", num_envs=edge - 1, seed=21)
    small.reset(); small.step(torch.zeros(edge - 1, device=small.device))
    assert small.launch_count() <= 3                                     # ... one launch per step below the threshold
    with pytest.raises((ValueError, RuntimeError)):
        BatchedJitterbugEnv("move_from_origin", num_envs=0)
    with pytest.raises(ValueError):
        big.step(torch.zeros(7, device=big.device))
