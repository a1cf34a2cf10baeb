"""Synthetic inputs in the style of the reference's helper_scripts/create_artificial_maps.py:70-86 (rectangles of a
label over a default label; later rectangles overwrite earlier ones) and of the scout simulator's observation pattern
(mav_simulator/src/mav_simulator.cpp:156-159,438-479: one init square at the start, then fov x fov patches along the
flown trajectory). Recipe and seeds: SURVEY.md section 8(d). Pure numpy; used by tests and bench.py to make inputs."""
import numpy as np

MPP = 0.0625  # m_per_pixel of every shipped map (helper_scripts/create_map_config.py:131)


def make_labels(W, H, seed):
    rng = np.random.default_rng(seed)
    lab = np.full((H, W), 30, np.int32)
    R = max(4, int(round(64 * (W / 640.0) * (H / 480.0))))
    # keep the generation time bounded on very large maps: the rectangle count saturates, sizes scale with the map
    R = min(R, 4096)
    for _ in range(R):
        w = int(rng.integers(W // 40, W // 4 + 1))
        h = int(rng.integers(H // 40, H // 4 + 1))
        x = int(rng.integers(0, W))
        y = int(rng.integers(0, H))
        lab[y:y + h, x:x + w] = int(rng.integers(30, 121))
    for _ in range(max(1, R // 4)):
        w = int(rng.integers(max(1, W // 80), W // 10 + 1))
        h = int(rng.integers(max(1, H // 80), H // 10 + 1))
        x = int(rng.integers(0, W))
        y = int(rng.integers(0, H))
        lab[y:y + h, x:x + w] = 0
    # start px (8,8), goal px (W-8,H-8): clear a 33x33 square of label 30 around both
    for cx, cy in ((8, 8), (W - 8, H - 8)):
        lab[max(0, cy - 16):cy + 17, max(0, cx - 16):cx + 17] = 30
    return lab


def make_observed(W, H, seed, K, init=320, fov=65, step_m=2.0):
    rng = np.random.default_rng(seed + 7919)
    obs = np.zeros((H, W), np.uint8)
    obs[0:min(H, init), 0:min(W, init)] = 1
    x, y = 0.5, 0.5
    width, height = W * MPP, H * MPP
    for _ in range(K):
        a = rng.uniform(0, 2 * np.pi)
        x = float(np.clip(x + step_m * np.cos(a), 0.0, width - MPP))
        y = float(np.clip(y + step_m * np.sin(a), 0.0, height - MPP))
        ix, iy = int(x / MPP), int(y / MPP)
        x0, y0 = ix - fov // 2, iy - fov // 2
        obs[max(0, y0):max(0, y0 + fov), max(0, x0):max(0, x0 + fov)] = 1
    return obs


def make_candidates(W, H, seed, T):
    rng = np.random.default_rng(seed + 104729)
    xy = np.empty((T, 2), np.float64)
    xy[:, 0] = rng.uniform(0.0, W * MPP, T)
    xy[:, 1] = rng.uniform(0.0, H * MPP, T)
    cost = rng.uniform(0.2, 5.0, T)
    return xy, cost


def world(W, H, seed, K, desc_fn, init_cost="min", **kw):
    """labels, observed mask and a map description (desc_fn = any make_desc-compatible constructor)."""
    lab = make_labels(W, H, seed)
    obs = make_observed(W, H, seed, K)
    ids = sorted(set(np.unique(lab).tolist()) - {0})
    init = min(ids) if init_cost == "min" else (max(ids) if init_cost == "max" else float(init_cost))
    desc = desc_fn(W, H, W * MPP, H * MPP, (0.5, 0.5), (W * MPP - 0.5, H * MPP - 0.5), min(ids), max(ids), init, **kw)
    return lab, obs, desc


def make_tree(W, H, seed, T):
    """RRT-like candidate tree for config 1: T nodes, node 0 = the root at the start pose, every other node hangs off
    the nearest earlier node; renumbered in pre-order (children in creation order) as sipp_update_tree expects.
    Returns parent[T] (int32, parent[0] = -1), end_xy[T,2], start_xy[T,2] (= the parent's end point), cost[T] (segment length at
    v_max 1 m/s: SegmentTime, cfg/planners/example_config.yaml:28,45-46)."""
    rng = np.random.default_rng(seed + 15485863)
    pts = np.empty((T, 2), np.float64)
    pts[0] = (0.5, 0.5)
    pts[1:, 0] = rng.uniform(0.0, W * MPP, T - 1)
    pts[1:, 1] = rng.uniform(0.0, H * MPP, T - 1)
    par = np.full(T, -1, np.int64)
    for i in range(1, T):
        d2 = ((pts[:i] - pts[i]) ** 2).sum(axis=1)
        par[i] = int(np.argmin(d2))
    kids = [[] for _ in range(T)]
    for i in range(1, T):
        kids[par[i]].append(i)
    order, stack = [], [0]
    while stack:
        v = stack.pop()
        order.append(v)
        stack.extend(reversed(kids[v]))
    new = np.empty(T, np.int64)
    new[np.array(order)] = np.arange(T)
    parent = np.full(T, -1, np.int32)
    end = np.empty((T, 2), np.float64)
    for old in range(T):
        end[new[old]] = pts[old]
        if old:
            parent[new[old]] = new[par[old]]
    start = np.where(parent[:, None] >= 0, end[np.maximum(parent, 0)], end)
    cost = np.sqrt(((end - start) ** 2).sum(axis=1))
    cost[0] = 0.0
    return parent, end, start, cost
