import sys, torch, numpy as np
sys.path.insert(0, '.')
from jitterbug_b200 import BatchedJitterbugEnv, TASKS
N = 512
for task in TASKS:
    env = BatchedJitterbugEnv(task, num_envs=N, seed=11)
    obs = env.reset()
    tgt = env.get_target()
    rows = []
    ret = torch.zeros(N, device=env.device)
    for k in range(999):
        obs, rew, st = env.step(env.heuristic_actions(obs))
        ret += rew
        if k % 111 == 110:
            q, v, _ = env.get_state()
            d = torch.hypot(q[:, 0] - tgt[:, 0], q[:, 1] - tgt[:, 1])
            zz = 1 - 2 * (q[:, 4] ** 2 + q[:, 5] ** 2)
            rows.append(f"t={k+1}: r={rew.mean().item():.3f} dist={d.mean().item():.3f} |xy|={torch.hypot(q[:,0],q[:,1]).mean().item():.3f} upright={(zz>0.7).float().mean().item():.2f} mv={v[:,14].abs().mean().item():.0f}")
    print(task, "return 5/50/95 pct:", np.percentile(ret.cpu().numpy(), [5, 50, 95]).round(1), env.counters()["diverged_resets"])
    for r in rows[::2]: print("   ", r)
