"""Random inputs of the reference-executed golden vectors (tests/golden/make_reference_golden.py).  The fixtures store
the reference's OUTPUTS only; generator and tests regenerate the inputs from the seeds below (numpy's legacy
``RandomState`` streams are stable across numpy versions) and check a checksum stored next to the outputs."""
import numpy as np

TASKS = ("move_from_origin", "face_direction", "move_in_direction", "move_to_position", "move_to_pose")
EXTRA_KEYS = {"move_from_origin": [], "face_direction": ["angle_to_target"],
              "move_in_direction": ["angle_to_target", "speed_in_target_frame"],
              "move_to_position": ["target_in_jitterbug_frame"],
              "move_to_pose": ["target_in_jitterbug_frame", "angle_to_target"]}
KEY_DIMS = {"angle_to_target": 1, "speed_in_target_frame": 3, "target_in_jitterbug_frame": 3}
N_HEURISTIC, SEED_HEURISTIC = 12000, 20260101
N_TASK_STATES, SEED_TASK_STATES = 600, 20260102


def heuristic_observations(task, n=N_HEURISTIC, seed=SEED_HEURISTIC):
    """Normalised observation vectors [n, D] as float32 (so fp32 device code sees exactly the reference's inputs)."""
    rs = np.random.RandomState(seed + 17 * TASKS.index(task))
    D = 15 + sum(KEY_DIMS[k] for k in EXTRA_KEYS[task])
    obs = rs.uniform(-1.2, 1.2, size=(n, D))
    # a share of the samples in the interesting regions: tiny target offsets (the 0.01 norm test of move_to_pose), motor
    # angles around the kick threshold -- pi/4 compared against a *normalised* angle
    k = n // 4
    obs[:k, 15:] *= 0.02
    obs[k:2 * k, 13] = rs.choice([-1, 1], k) * (np.pi / 4 + rs.normal(0, 0.05, k))
    return obs.astype(np.float32)


def task_states(task, n=N_TASK_STATES, seed=SEED_TASK_STATES):
    """(qpos[n,16], qvel[n,15], target[n,3] = x, y, yaw): every field the task layer reads is exercised."""
    rs = np.random.RandomState(seed + 17 * TASKS.index(task))
    qpos = np.zeros((n, 16)); qvel = np.zeros((n, 15))
    qpos[:, 0:2] = rs.uniform(-1.5, 1.5, (n, 2)); qpos[:, 2] = rs.uniform(0.0, 0.08, n)
    q = rs.normal(size=(n, 4)); q[: n // 2, 1:3] *= 0.05             # half of them roughly upright
    qpos[:, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    qpos[:, 7:15] = rs.uniform(-0.2, 0.2, (n, 8))
    qpos[:, 15] = rs.uniform(-40, 40, n)                               # many motor revolutions
    qvel[:, 0:3] = rs.uniform(-0.5, 0.5, (n, 3)); qvel[:, 3:6] = rs.uniform(-20, 20, (n, 3))
    qvel[:, 6:14] = rs.uniform(-30, 30, (n, 8)); qvel[:, 14] = rs.uniform(-170, 170, n)
    tgt = np.zeros((n, 3))
    ang, rad = rs.uniform(0, 2 * np.pi, n), rs.uniform(.05, .2, n)
    tgt[:, 0], tgt[:, 1], tgt[:, 2] = rad * np.cos(ang), rad * np.sin(ang), rs.uniform(0, 2 * np.pi, n)
    # a share of the Jitterbugs close to their target (the gaussian position reward lives within a few cm)
    k = n // 3
    qpos[:k, 0:2] = tgt[:k, 0:2] + rs.normal(0, 0.03, (k, 2)); qpos[:k, 2] = 0.035 + rs.normal(0, 0.003, k)
    # slow ones too (the linear velocity reward saturates at 0.1 m/s)
    qvel[k:2 * k, 0:3] *= 0.2
    if task == "move_from_origin":
        tgt[:] = 0
    elif task in ("face_direction", "move_in_direction"):
        tgt[:, :2] = 0
    elif task == "move_to_position":
        tgt[:, 2] = 0
    return qpos, qvel, tgt
