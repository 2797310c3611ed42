"""The only MuJoCo-EXECUTED numbers the reference repository holds for this path: the training curves of its paper
(figure "fig-rl-perf"), stored as the rendered PNG inside ``/root/reference/fig-rl-perf.ipynb`` (cell 2's output).
Their LEFT EDGE is the per-episode return of a still-untrained policy on the real dm_control / MuJoCo environment --
for stable-baselines PPO2 (benchmarks/benchmark.py:290-325: fresh MlpPolicy, log-std 0, actions clipped to the action
space) that is a_t = clip(N(0, 1), -1, 1), independent per control step, which any simulator can replay.

    python tests/golden/read_published_curves.py          # writes tests/golden/published_training_start.json

The curves are drawn by ``benchmarks/plot_train_progress.py:73-131``: per-episode 10th / 50th / 90th percentile over the
monitor files of all seeds (40-80 files for PPO2, 10 for DDPG), then a 20-episode moving average, x = episode index
(one pixel column = 5.6 episodes), y from -50 to 1050.  This script digitises the first columns of every panel: the
median line (matplotlib C0) and the edges of the shaded band, calibrated on the figure's own grid lines (y = 1000, 800,
... rows) and checked on the red "solved" line (y = 900).  Reading error: about +-5 return units (line width, blending).

Only the PPO2 column is a clean "random policy" reading (flat for >= 1e6 steps on move_from_origin and
move_in_direction); the DDPG column trains from step 256 on (50 gradient steps per 100 env steps) and face_direction
climbs within the first pixels, so those are recorded but carry ``"usable": false``.
"""
import base64
import io
import json
import os
import re

import numpy as np

REF = os.environ.get("JB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
TASKS = ("move_from_origin", "face_direction", "move_in_direction", "move_to_position", "move_to_pose")
Y1000_ROW = (100.0, 649.0, 1198.0, 1746.5, 2295.0)       # pixel row of the y = 1000 grid line of each panel row
PX_PER_UNIT = 95.2 / 200.0                               # spacing of the 200-unit grid lines
X0_COL = {"ddpg": 204, "ppo2": 1204}                      # pixel column of the x = 0 grid line of each panel column
N_COLS = 12                                              # columns read: episodes ~10 .. ~80 of every monitor file


def figure_png():
    nb = json.load(open(os.path.join(REF, "fig-rl-perf.ipynb")))
    for cell in nb["cells"]:
        for out in cell.get("outputs", []):
            if "image/png" in out.get("data", {}):
                return base64.b64decode(out["data"]["image/png"])
    raise RuntimeError("no rendered figure in fig-rl-perf.ipynb")


def digitise():
    from PIL import Image
    im = np.asarray(Image.open(io.BytesIO(figure_png())).convert("RGB")).astype(int)
    assert im.shape == (2921, 2165, 3), im.shape
    dist = lambda rgb: np.abs(im - np.array(rgb)).sum(2)
    curve = dist((31, 119, 180)) < 120                                      # C0, anti-aliased
    band = (dist((214, 222, 236)) < 10) | (dist((233, 241, 248)) < 8)        # C0 at alpha 0.1 over the axes / a grid line
    red = (im[:, :, 0] > 200) & (im[:, :, 1] < 60) & (im[:, :, 2] < 60)
    out = {}
    for p, task in enumerate(TASKS):
        val = lambda row: 1000.0 - (row - Y1000_ROW[p]) / PX_PER_UNIT
        r0, r1 = int(Y1000_ROW[p] - 30), int(Y1000_ROW[p] + PX_PER_UNIT * 1050 + 5)
        solved = val(np.where(red[r0:r1, 300])[0].mean() + r0)
        assert abs(solved - 900.0) < 4.0, (task, solved)                    # calibration check on the red line
        for alg, x0 in X0_COL.items():
            rows = []
            for x in range(x0 + 2, x0 + 2 + N_COLS):
                cr = np.where(curve[r0:r1, x])[0]
                br = np.where((band | curve)[r0:r1, x])[0]
                if len(cr):
                    rows.append((val(cr.mean() + r0), val(br.max() + r0), val(br.min() + r0)))
            a = np.array(rows)
            usable = alg == "ppo2" and task != "face_direction"
            out[f"{task}/{alg}"] = dict(median=round(float(np.median(a[:, 0])), 1), p10=round(float(np.median(a[:, 1])), 1),
                                        p90=round(float(np.median(a[:, 2])), 1), first_column_median=round(float(a[0, 0]), 1),
                                        columns=len(a), usable=usable)
    return out


def digitise_over_training(step_px=8):
    """The whole PPO2 curve of move_from_origin and move_in_direction: [episode index / 1000, median, p10, p90] every
    ``step_px`` pixel columns (45 episodes).  During training the policy of these two panels does not improve (the
    paper reports PPO2 failing on them); what changes is mostly its std, so the curve traces a one-parameter family of
    (median, p10, p90) triples that a simulator can be asked to reproduce for SOME std at every point."""
    from PIL import Image
    im = np.asarray(Image.open(io.BytesIO(figure_png())).convert("RGB")).astype(int)
    dist = lambda rgb: np.abs(im - np.array(rgb)).sum(2)
    curve = dist((31, 119, 180)) < 120
    band = (dist((214, 222, 236)) < 10) | (dist((233, 241, 248)) < 8)
    out = {}
    x0 = X0_COL["ppo2"]
    for task in ("move_from_origin", "move_in_direction"):
        p = TASKS.index(task)
        val = lambda row: 1000.0 - (row - Y1000_ROW[p]) / PX_PER_UNIT
        r0, r1 = int(Y1000_ROW[p] - 30), int(Y1000_ROW[p] + PX_PER_UNIT * 1050 + 5)
        rows = []
        for x in range(x0 + 12, x0 + 900, step_px):
            cr = np.where(curve[r0:r1, x])[0]
            br = np.where((band | curve)[r0:r1, x])[0]
            if len(cr) == 0:
                break                                   # the curve ends (runs were stopped at different lengths)
            rows.append([round((x - x0) / 177.4, 3), round(float(val(cr.mean() + r0)), 1),
                         round(float(val(br.max() + r0)), 1), round(float(val(br.min() + r0)), 1)])
        out[task] = rows
    return out


def main():
    data = dict(
        source="fig-rl-perf.ipynb, cell 2, image/png output (2165 x 2921 px), digitised by tests/golden/read_published_curves.py",
        meaning="per-episode return percentiles over all monitor files, 20-episode moving average, first ~80 episodes",
        policy="ppo2: stable-baselines PPO2 at initialisation = clip(N(0,1), -1, 1) per control step; ddpg: trains from step 256, not a fixed policy",
        reading_error=5.0,
        curves=digitise(),
        ppo2_over_training=digitise_over_training())
    path = os.path.join(HERE, "published_training_start.json")
    text = json.dumps(data, indent=1)
    text = re.sub(r"\[\s*(-?[\d.]+),\s*(-?[\d.]+),\s*(-?[\d.]+),\s*(-?[\d.]+)\s*\]", r"[\1, \2, \3, \4]", text)   # one row per line
    with open(path, "w") as f:
        f.write(text + "\n")
    print(path)
    for k, v in data["curves"].items():
        print(f"  {k:28s} median {v['median']:6.1f}  p10 {v['p10']:6.1f}  p90 {v['p90']:6.1f}  usable {v['usable']}")


if __name__ == "__main__":
    main()
