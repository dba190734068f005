"""Planning-problem inputs: saved worlds, the armour.in text format and the synthetic generators of
the benchmark configurations.

File formats follow the reference: world CSVs (kinova_src/saved_worlds/random/*.csv: row 1 start,
row 2 goal, row 3 NaN, rows 4+ `cx,cy,cz,sx,sy,sz,NaN`; obstacle zonotope = [c, diag(s/2)],
simulator/worlds/obstacles/box_obstacle_zonotope.m:21-26) and `armour.in` (4 lines of 7 `%.10f`,
the obstacle count, one line of 12 `%.10f` per obstacle: kinova_simulator_interfaces/uarmtd_planner.m:158-185,
parsed by kinova_planner_realtime/armour_main.cu:53-76).
"""
from __future__ import annotations

import numpy as np

NF = 7
STATE_LB = np.array([-np.pi, -2.41, -np.pi, -2.66, -np.pi, -2.23, -np.pi])  # continuous joints sampled in [-pi, pi]
STATE_UB = -STATE_LB
SPEED_LIMITS = np.array([1.3963, 1.3963, 1.3963, 1.3963, 1.2218, 1.2218, 1.2218])
CONTINUOUS = np.array([True, False, True, False, True, False, True])


def quantize(x):
    """Round-trip through the `%.10f` text MATLAB writes, so every consumer sees the same doubles."""
    a = np.asarray(x, dtype=np.float64)
    return np.array([float("%.10f" % v) for v in a.ravel()]).reshape(a.shape)


def boxes_to_zonotopes(boxes):
    """[n, 6] (centre, side lengths) -> [n, 12] (centre, 3 generators = diag(side/2), column-major Z)."""
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 6)
    out = np.zeros((boxes.shape[0], 12))
    out[:, 0:3] = boxes[:, 0:3]
    for a in range(3):
        out[:, 3 + 4 * a] = boxes[:, 3 + a] / 2
    return out


def load_world_csv(path):
    rows = np.genfromtxt(path, delimiter=",")
    start, goal = rows[0, :NF].copy(), rows[1, :NF].copy()
    return start, goal, boxes_to_zonotopes(rows[3:, :6])


def angdiff(a, b):
    """MATLAB angdiff(a, b): b - a wrapped to [-pi, pi]."""
    d = np.asarray(b) - np.asarray(a)
    return (d + np.pi) % (2 * np.pi) - np.pi


def straight_line_waypoint(q_cur, q_goal, lookahead=0.1):
    """robot_arm_straight_line_HLP.get_waypoint (simulator/planners/high_level_planners/...:45-57)."""
    d = np.asarray(q_goal, dtype=np.float64) - np.asarray(q_cur, dtype=np.float64)
    d[CONTINUOUS] = angdiff(np.asarray(q_cur)[CONTINUOUS], np.asarray(q_goal)[CONTINUOUS])
    return np.asarray(q_cur) + lookahead * d / np.linalg.norm(d)


def write_armour_in(path, q0, qd0, qdd0, q_des, obstacles):
    obstacles = np.asarray(obstacles, dtype=np.float64).reshape(-1, 12)
    with open(path, "w") as f:
        for v in (q0, qd0, qdd0, q_des):
            f.write("".join("%.10f " % x for x in v) + "\n")
        f.write("%d\n" % obstacles.shape[0])
        for o in obstacles:
            f.write("".join("%.10f " % x for x in o) + "\n")


def read_armour_in(path):
    tok = open(path).read().split()
    vals = [float(x) for x in tok[:4 * NF]]
    q0, qd0, qdd0, q_des = (np.array(vals[i * NF:(i + 1) * NF]) for i in range(4))
    nobs = int(tok[4 * NF])
    obs = np.array([float(x) for x in tok[4 * NF + 1:4 * NF + 1 + nobs * 12]]).reshape(nobs, 12)
    return q0, qd0, qdd0, q_des, obs


def config1_problem(csv_path):
    """BASELINE config 1: one saved world, qd0 = qdd0 = 0, straight-line waypoint, %.10f-quantised."""
    start, goal, obs = load_world_csv(csv_path)
    q_des = straight_line_waypoint(start, goal, 0.1)
    z = np.zeros(NF)
    return quantize(start), z, z.copy(), quantize(q_des), quantize(obs)


def random_problems(nprob, nobs, seed=20261017, moving=True):
    """BASELINE config 2 generator: random start / goal / motion state and `nobs` random boxes per problem."""
    rng = np.random.default_rng(seed)
    q0 = rng.uniform(STATE_LB, STATE_UB, (nprob, NF))
    goal = rng.uniform(STATE_LB, STATE_UB, (nprob, NF))
    qd0 = rng.uniform(-0.5, 0.5, (nprob, NF)) * SPEED_LIMITS if moving else np.zeros((nprob, NF))
    qdd0 = rng.uniform(-1, 1, (nprob, NF)) if moving else np.zeros((nprob, NF))
    boxes = np.empty((nprob, nobs, 6))
    for p in range(nprob):
        n = 0
        while n < nobs:
            c = rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2])
            s = rng.uniform(0.01, 0.5, 3)
            if np.all(np.abs(c - np.array([0, 0, 0.05])) <= s / 2 + 0.1):  # would swallow the robot base
                continue
            boxes[p, n] = np.concatenate([c, s])
            n += 1
    obs = boxes_to_zonotopes(boxes.reshape(-1, 6)).reshape(nprob, nobs, 12)
    q_des = np.stack([straight_line_waypoint(q0[p], goal[p]) for p in range(nprob)])
    return quantize(q0), quantize(qd0), quantize(qdd0), quantize(q_des), quantize(obs)


def halton_k(n, dim=NF, skip=1):
    """Deterministic k schedule in [-1, 1]^7 (Halton points) for exercising eval_g / eval_jac_g."""
    primes = [2, 3, 5, 7, 11, 13, 17][:dim]
    out = np.empty((n, dim))
    for j, b in enumerate(primes):
        for i in range(n):
            f, r, x = 1.0, 0.0, i + skip
            while x > 0:
                f /= b
                r += f * (x % b)
                x //= b
            out[i, j] = 2 * r - 1
    return out
