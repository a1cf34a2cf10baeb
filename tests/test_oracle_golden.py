"""Pins the oracle's follower-path-cost step (A11) against the reference's golden vectors.

Golden values: docker/data/maps/cfg/map_*.yaml:15 (`optimal_prm_rov_path_cost`), produced by the reference's own
planexp_eval_planner (planexp_base_planners/src/planexp_eval_planner.cpp:201-303, printed with 6 significant digits,
utils.h:367-373) on the fully known map with obstacle ids = obstacle_ids_rov U {unknown_id} and init_cost = min cost
(utils.h:238-302).
"""
import math

import numpy as np
import pytest

from oracle import oracle as orc

# full-precision values of the same quantity — regression guard for ulp-level drift
FULL_PRECISION = {
    "map_102": 2207.3687883178195, "map_110": 2530.1453506242087, "map_111": 2692.8320950953685,
    "map_112": 2692.217363074229, "map_113": 3614.953975785665, "map_115": 1527.2913665038786,
    "map_122": 4145.212770161963, "map_212": 4270.462718314068, "map_312": 3057.4932302253173,
}


def eval_planner_desc(cfg, **kw):
    obst = list(cfg["obstacle_ids_rov"]) + [cfg["unknown_id"]]  # utils.h:272
    ids = [i for i in cfg["present_ids"] if i not in obst]
    return orc.make_desc(cfg["W"], cfg["H"], cfg["width"], cfg["height"], cfg["start"], cfg["goal"], min(ids), max(ids),
                         min(ids), obstacle_ids_mav=cfg["obstacle_ids_mav"], obstacle_ids_rov=obst,
                         unknown_id=cfg["unknown_id"], **kw)


def sig6(x):
    return float("%.6g" % x)  # default ostream precision


@pytest.mark.parametrize("name", ["map_102", "map_110", "map_111", "map_112", "map_113", "map_115", "map_116", "map_122",
                                  "map_212", "map_312"])
def test_golden_prm_cost(golden, name):
    labels, cfg = golden[name]
    o = orc.Oracle(eval_planner_desc(cfg))
    o.map_upload_full(labels.astype(np.int32))
    cost, status, path, n = o.plan()
    if cfg["golden_prm_cost"] < 0:
        assert cost == -1.0 and status == orc.PLAN_TERMINAL_INVALID and n == 0
        return
    assert sig6(cost) == cfg["golden_prm_cost"]
    assert cost == FULL_PRECISION[name]
    assert status == orc.PLAN_TERMINAL_VALID  # every path node explored on a fully known map
    assert o.oob_events() == 0
    # path endpoints are the snapped start / goal nodes
    g = o.geometry()
    ids = o.path_nodes()
    assert (ids[0] // cfg["H"], ids[0] % cfg["H"]) == g["start_node"]
    assert (ids[-1] // cfg["H"], ids[-1] % cfg["H"]) == g["goal_node"]
    # field is a fixed point on the path: d strictly increases start -> goal
    f = o.field_download()
    d = np.array([f[i % cfg["H"], i // cfg["H"]] for i in ids])
    assert d[0] == 0.0 and np.all(np.diff(d) > 0) and d[-1] == cost


def test_geometry_constants(golden):
    """640x480 geometry constants and the diagonal-admission count."""
    labels, cfg = golden["map_212"]
    o = orc.Oracle(eval_planner_desc(cfg))
    g = o.geometry()
    assert g["spacing"] == 0.062402496099843996
    assert g["off_x"] == 0.0
    assert g["off_y"] == -0.007800312012481214
    assert g["max_dist"] == 0.08825045631033356
    assert g["start_node"] == (8, 8) and g["goal_node"] == (633, 473)
    assert o.diag_admitted() == 164544


def test_round_half_away(golden):
    """map_115's start y lands on 240.5 before rounding: std::round -> 241 (prm_star_planner.h:81)."""
    _, cfg = golden["map_115"]
    g = orc.Oracle(eval_planner_desc(cfg)).geometry()
    assert g["start_node"] == (240, 241) and g["goal_node"] == (426, 241)


@pytest.mark.parametrize("W,expected", [(1024, 48608)])
def test_diag_admission_other_sizes(W, expected):
    d = orc.make_desc(W, W, W / 16.0, W / 16.0, (0.5, 0.5), (W / 16.0 - 0.5, W / 16.0 - 0.5), 30, 120, 30)
    assert orc.Oracle(d).diag_admitted() == expected
