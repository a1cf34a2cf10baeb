"""ctypes front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product (scouting-ipp_b200/, libsipp.so) never imports this module.

The structs mirror include/sipp.h (sipp_map_desc, sipp_eval_params); see sipp_oracle.cpp for the
reference file:line each function follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAX_IDS = 32
COUNTS_PER_CANDIDATE = 4

EVAL_RRT_STAR, EVAL_GOAL_DISTANCE, EVAL_EXPAND_REACHABLE, EVAL_AVERAGE_VIEW_COST, EVAL_EXPAND_LOW_COST, EVAL_NAIVE = range(6)
STATE_OCCUPIED, STATE_FREE, STATE_UNKNOWN = 0, 1, 2
PLAN_VALID, PLAN_INVALID, PLAN_TERMINAL_VALID, PLAN_TERMINAL_INVALID = 0, 1, 2, 3


class MapDesc(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("H", C.c_int32),
        ("width", C.c_double), ("height", C.c_double),
        ("start_x", C.c_double), ("start_y", C.c_double),
        ("goal_x", C.c_double), ("goal_y", C.c_double),
        ("min_rov_cost", C.c_double), ("max_rov_cost", C.c_double), ("unknown_init_cost", C.c_double),
        ("unknown_id", C.c_int32),
        ("n_obstacle_ids_mav", C.c_int32), ("obstacle_ids_mav", C.c_int32 * MAX_IDS),
        ("n_obstacle_ids_rov", C.c_int32), ("obstacle_ids_rov", C.c_int32 * MAX_IDS),
        ("compute_reachability", C.c_int32), ("compute_closest_observed", C.c_int32),
    ]


UPD_RRT_STAR, UPD_CONSTRAINED, UPD_DIRECT = range(3)
FOLLOW_UPDATE_ALL, FOLLOW_UPDATE_ALL_NON_TERMINAL = range(2)


class UpdaterParams(C.Structure):  # sipp_updater_params
    _fields_ = [
        ("updater", C.c_int32), ("following", C.c_int32),
        ("minimum_gain", C.c_double), ("update_range", C.c_double),
        ("update_gain", C.c_int32), ("update_cost", C.c_int32), ("update_value", C.c_int32), ("pad_", C.c_int32),
    ]


def make_updater(updater=UPD_RRT_STAR, following=FOLLOW_UPDATE_ALL, minimum_gain=-1e9, update_range=1e9, update_gain=True,
                 update_cost=True, update_value=True):
    u = UpdaterParams()
    u.updater, u.following = int(updater), int(following)
    u.minimum_gain, u.update_range = float(minimum_gain), float(update_range)
    u.update_gain, u.update_cost, u.update_value = int(bool(update_gain)), int(bool(update_cost)), int(bool(update_value))
    return u


class EvalParams(C.Structure):
    _fields_ = [
        ("evaluator", C.c_int32), ("half_fov", C.c_int32),
        ("non_path_voxel_gain", C.c_double),
        ("use_distance_discount", C.c_int32), ("do_binary_check", C.c_int32),
        ("discount_offset", C.c_double),
        ("max_gain", C.c_double),
        ("discount_factor", C.c_double),
        ("augment_unknown_with_mean", C.c_int32), ("use_bounding_volume", C.c_int32),
        ("bv_x_min", C.c_double), ("bv_x_max", C.c_double), ("bv_y_min", C.c_double), ("bv_y_max", C.c_double),
    ]


class CostParams(C.Structure):  # sipp_cost_params
    _fields_ = [("accumulate", C.c_int32), ("pad_", C.c_int32), ("max_sample_distance", C.c_double)]


TRAVEL_MAP_COST, TRAVEL_DISTANCE_COST, TRAVEL_TIME_COST = range(3)


class TravelParams(C.Structure):  # sipp_travel_params
    _fields_ = [("formulation", C.c_int32), ("pad_", C.c_int32), ("max_step_size", C.c_double), ("mav_speed", C.c_double)]


def make_travel_params(formulation=TRAVEL_MAP_COST, max_step_size=0.05, mav_speed=1.0):
    p = TravelParams()
    p.formulation, p.max_step_size, p.mav_speed = int(formulation), float(max_step_size), float(mav_speed)
    return p


def make_cost_params(accumulate=False, max_sample_distance=0.05):
    p = CostParams()
    p.accumulate, p.max_sample_distance = int(bool(accumulate)), float(max_sample_distance)
    return p


def make_desc(W, H, width, height, start, goal, min_cost, max_cost, init_cost, obstacle_ids_mav=(0,),
              obstacle_ids_rov=(0,), unknown_id=3599, compute_reachability=True, compute_closest_observed=False):
    d = MapDesc()
    d.W, d.H, d.width, d.height = int(W), int(H), float(width), float(height)
    d.start_x, d.start_y = float(start[0]), float(start[1])
    d.goal_x, d.goal_y = float(goal[0]), float(goal[1])
    d.min_rov_cost, d.max_rov_cost, d.unknown_init_cost = float(min_cost), float(max_cost), float(init_cost)
    d.unknown_id = int(unknown_id)
    d.n_obstacle_ids_mav = len(obstacle_ids_mav)
    for i, v in enumerate(obstacle_ids_mav):
        d.obstacle_ids_mav[i] = int(v)
    d.n_obstacle_ids_rov = len(obstacle_ids_rov)
    for i, v in enumerate(obstacle_ids_rov):
        d.obstacle_ids_rov[i] = int(v)
    d.compute_reachability = int(bool(compute_reachability))
    d.compute_closest_observed = int(bool(compute_closest_observed))
    return d


def make_params(evaluator=EVAL_RRT_STAR, half_fov=32, non_path_voxel_gain=0.0, use_distance_discount=False,
                do_binary_check=False, discount_offset=0.1, max_gain=100.0, discount_factor=0.1,
                augment_unknown_with_mean=False, bounding_volume=None):
    """Defaults are the reference's setupFromParamMap defaults (half_fov: shipped cfg value 32)."""
    p = EvalParams()
    p.evaluator, p.half_fov = int(evaluator), int(half_fov)
    p.non_path_voxel_gain = float(non_path_voxel_gain)
    p.use_distance_discount, p.do_binary_check = int(bool(use_distance_discount)), int(bool(do_binary_check))
    p.discount_offset, p.max_gain, p.discount_factor = float(discount_offset), float(max_gain), float(discount_factor)
    p.augment_unknown_with_mean = int(bool(augment_unknown_with_mean))
    if bounding_volume is None:
        p.use_bounding_volume = 0
    else:
        p.use_bounding_volume = 1
        p.bv_x_min, p.bv_x_max, p.bv_y_min, p.bv_y_max = [float(v) for v in bounding_volume]
    return p


def build(force=False):
    """Compile liboracle.so (g++, a second or two). Building the checker is not using it."""
    src = os.path.join(_HERE, "sipp_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(MapDesc)]
        L.orc_destroy.argtypes = [vp]
        L.orc_destroy.restype = None
        L.orc_map_clear.argtypes = [vp]
        L.orc_patch_origin.argtypes = [vp, f64, f64, i32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.orc_map_update_patch.argtypes = [vp, i32, i32, i32, i32, vp]
        L.orc_map_upload_full.argtypes = [vp, vp, vp]
        L.orc_update_unknown_cost.argtypes = [vp, f64]
        L.orc_map_stats.argtypes = [vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(C.c_int64)]
        L.orc_map_download.argtypes = [vp, vp, vp, vp, vp]
        L.orc_reach_ids.argtypes = [vp, vp]
        L.orc_closest_download.argtypes = [vp, vp, vp]
        L.orc_planner_download.argtypes = [vp, vp, vp, vp]
        L.orc_oob_events.argtypes = [vp]
        L.orc_geometry.argtypes = [vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64), C.POINTER(f64), vp]
        L.orc_diag_admitted.argtypes = [vp]
        L.orc_diag_admitted.restype = C.c_int64
        L.orc_plan.argtypes = [vp, C.POINTER(f64), C.POINTER(i32), vp, i32, C.POINTER(i32), i32]
        L.orc_field_download.argtypes = [vp, vp]
        L.orc_path_nodes.argtypes = [vp, vp, i32]
        L.orc_path_mask_generation.argtypes = [vp]
        L.orc_set_path_mask.argtypes = [vp, vp]
        L.orc_gain_batch.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, i32, i32]
        L.orc_value_select.argtypes = [i32, vp, vp, vp, vp, C.POINTER(i32), C.POINTER(f64)]
        L.orc_cost_integral.argtypes = [vp, C.POINTER(CostParams), i32, vp, vp, vp, vp, vp]
        L.orc_travel_cost.argtypes = [vp, vp, C.POINTER(TravelParams), i32, vp, vp, vp]
        L.orc_cycle.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp, C.POINTER(i32), C.POINTER(f64), i32, i32]
        L.orc_update_tree.argtypes = [vp, C.POINTER(EvalParams), C.POINTER(UpdaterParams), i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]
        L.orc_update_tree_views.argtypes = [vp, C.POINTER(EvalParams), C.POINTER(UpdaterParams), i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
        L.orc_gain_batch_views.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp, i32]
        L.orc_sample_viewpoints.argtypes = [vp, i32, f64, vp, i32]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU oracle context; the method names mirror scouting-ipp_b200/sipp's Context so parity tests read symmetrically."""

    def __init__(self, desc):
        self.desc = desc
        self.W, self.H = desc.W, desc.H
        self._L = lib()
        self._h = self._L.orc_create(C.byref(desc))
        if not self._h:
            raise ValueError("orc_create failed (bad map description)")

    def close(self):
        if self._h:
            self._L.orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- map store
    def map_clear(self):
        self._L.orc_map_clear(self._h)

    def patch_origin(self, px, py, rows, cols):
        x0, y0 = C.c_int32(), C.c_int32()
        self._L.orc_patch_origin(self._h, px, py, rows, cols, C.byref(x0), C.byref(y0))
        return x0.value, y0.value

    def map_update_patch(self, x0, y0, labels):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        rows, cols = labels.shape
        self._L.orc_map_update_patch(self._h, x0, y0, rows, cols, _ptr(labels))

    def map_upload_full(self, labels, observed=None):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        assert labels.shape == (self.H, self.W)
        if observed is not None:
            observed = np.ascontiguousarray(observed, dtype=np.uint8)
            assert observed.shape == (self.H, self.W)
        self._L.orc_map_upload_full(self._h, _ptr(labels), _ptr(observed))

    def update_unknown_cost(self, c):
        self._L.orc_update_unknown_cost(self._h, float(c))

    def map_stats(self):
        m, mx, n = C.c_double(), C.c_double(), C.c_int64()
        self._L.orc_map_stats(self._h, C.byref(m), C.byref(mx), C.byref(n))
        return m.value, mx.value, n.value

    def map_download(self):
        state = np.empty((self.H, self.W), np.uint8)
        cost = np.empty((self.H, self.W), np.int32)
        reach = np.empty((self.H, self.W), np.uint8)
        path = np.empty((self.H, self.W), np.uint8)
        self._L.orc_map_download(self._h, _ptr(state), _ptr(cost), _ptr(reach), _ptr(path))
        return state, cost, reach, path

    def reach_ids(self):
        out = np.empty((self.H, self.W), np.int32)
        self._L.orc_reach_ids(self._h, _ptr(out))
        return out

    def closest_download(self):
        cc = np.empty((self.H, self.W), np.float64)
        dc = np.empty((self.H, self.W), np.float64)
        if self._L.orc_closest_download(self._h, _ptr(cc), _ptr(dc)) != 0:
            return None, None
        return cc, dc

    def planner_download(self):
        pmap = np.empty((self.H, self.W), np.float64)
        explored = np.empty((self.H, self.W), np.uint8)
        removed = np.empty((self.H, self.W), np.uint8)
        self._L.orc_planner_download(self._h, _ptr(pmap), _ptr(explored), _ptr(removed))
        return pmap, explored, removed

    def oob_events(self):
        return self._L.orc_oob_events(self._h)

    def geometry(self):
        s, ox, oy, md = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        nodes = np.zeros(4, np.int32)
        self._L.orc_geometry(self._h, C.byref(s), C.byref(ox), C.byref(oy), C.byref(md), _ptr(nodes))
        return dict(spacing=s.value, off_x=ox.value, off_y=oy.value, max_dist=md.value,
                    start_node=(int(nodes[0]), int(nodes[1])), goal_node=(int(nodes[2]), int(nodes[3])))

    def diag_admitted(self):
        return int(self._L.orc_diag_admitted(self._h))

    # ---- follower path cost
    def plan(self, update_mask=True, path_cap=None):
        cap = path_cap if path_cap is not None else 4 * (self.W + self.H)
        cost, status, n = C.c_double(), C.c_int32(), C.c_int32()
        path = np.zeros((cap, 2), np.float64)
        self._L.orc_plan(self._h, C.byref(cost), C.byref(status), _ptr(path), cap, C.byref(n), int(update_mask))
        return cost.value, status.value, path[: min(n.value, cap)].copy(), n.value

    def field_download(self):
        f = np.empty((self.H, self.W), np.float64)
        if self._L.orc_field_download(self._h, _ptr(f)) != 0:
            raise RuntimeError("no field: call plan() first")
        return f

    def path_nodes(self):
        cap = 4 * (self.W + self.H)
        ids = np.zeros(cap, np.int32)
        n = self._L.orc_path_nodes(self._h, _ptr(ids), cap)
        return ids[:n].copy()

    def path_mask_generation(self):
        return self._L.orc_path_mask_generation(self._h)

    def set_path_mask(self, mask):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        assert mask.shape == (self.H, self.W)
        self._L.orc_set_path_mask(self._h, _ptr(mask))

    # ---- gain / value / cycle
    def gain_batch(self, params, xy, start_xy=None, variant=0, threads=1):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        n = xy.shape[0]
        if start_xy is not None:
            start_xy = np.ascontiguousarray(start_xy, dtype=np.float64).reshape(-1, 2)
        gain = np.zeros(n, np.float64)
        counts = np.zeros((n, COUNTS_PER_CANDIDATE), np.uint32)
        terminal = np.zeros(n, np.uint8)
        self._L.orc_gain_batch(self._h, C.byref(params), n, _ptr(xy), _ptr(start_xy), _ptr(gain), _ptr(counts),
                               _ptr(terminal), int(variant), int(threads))
        return gain, counts, terminal

    def gain_batch_views(self, params, view_off, view_xy, start_xy=None, variant=1):
        """segments scored from several sampled viewpoints each (sensor_model/sampling_time > 0): segment i has the views
        view_xy[view_off[i]:view_off[i+1]]"""
        off = np.ascontiguousarray(view_off, dtype=np.int32)
        n = off.shape[0] - 1
        vxy = np.ascontiguousarray(view_xy, dtype=np.float64).reshape(-1, 2)
        sxy = None if start_xy is None else np.ascontiguousarray(start_xy, dtype=np.float64).reshape(n, 2)
        gain = np.zeros(n, np.float64)
        counts = np.zeros((n, COUNTS_PER_CANDIDATE), np.uint32)
        terminal = np.zeros(n, np.uint8)
        self._L.orc_gain_batch_views(self._h, C.byref(params), n, _ptr(off), _ptr(vxy), _ptr(sxy), _ptr(gain), _ptr(counts), _ptr(terminal), int(variant))
        return gain, counts, terminal

    def cost_integral(self, cparams, point_offset, points_xyz, parent_cost=None):
        off = np.ascontiguousarray(point_offset, dtype=np.int32)
        n = off.shape[0] - 1
        pts = np.ascontiguousarray(points_xyz, dtype=np.float64).reshape(-1, 3)
        pc = None if parent_cost is None else np.ascontiguousarray(parent_cost, dtype=np.float64)
        cost, valid = np.zeros(n, np.float64), np.zeros(n, np.uint8)
        self._L.orc_cost_integral(self._h, C.byref(cparams), n, _ptr(off), _ptr(pts), _ptr(pc), _ptr(cost), _ptr(valid))
        return cost, valid

    def travel_cost(self, truth, tparams, start_xyz, goal_xyz):
        """MavSimulator::computeCost (mav_simulator.cpp:785-821) over the ground-truth label map `truth`"""
        truth = np.ascontiguousarray(truth, dtype=np.int32)
        a = np.ascontiguousarray(start_xyz, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(goal_xyz, dtype=np.float64).reshape(-1, 3)
        cost = np.zeros(len(a), np.float64)
        self._L.orc_travel_cost(self._h, _ptr(truth), C.byref(tparams), len(a), _ptr(a), _ptr(b), _ptr(cost))
        return cost

    def cycle(self, params, xy, cost, start_xy=None, variant=0, threads=1):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        n = xy.shape[0]
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        if start_xy is not None:
            start_xy = np.ascontiguousarray(start_xy, dtype=np.float64).reshape(-1, 2)
        gain = np.zeros(n, np.float64)
        terminal = np.zeros(n, np.uint8)
        value = np.zeros(n, np.float64)
        bi, bv = C.c_int32(), C.c_double()
        self._L.orc_cycle(self._h, C.byref(params), n, _ptr(xy), _ptr(start_xy), _ptr(cost), _ptr(gain), _ptr(terminal),
                          _ptr(value), C.byref(bi), C.byref(bv), int(variant), int(threads))
        return gain, terminal, value, bi.value, bv.value


    def update_tree(self, params, updater, parent, end_xy, gain, cost, value, terminal, cur_xy, path_changed, start_xy=None, variant=1,
                    cost_below=None, view_off=None, view_xy=None):
        """literal pre-order emulation of one re-evaluation pass (orc_update_tree); returns (gain, value, terminal)"""
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        n = parent.shape[0]
        end_xy = np.ascontiguousarray(end_xy, dtype=np.float64).reshape(n, 2)
        sxy = None if start_xy is None else np.ascontiguousarray(start_xy, dtype=np.float64).reshape(n, 2)
        g = np.array(gain, dtype=np.float64, copy=True)
        v = np.array(value, dtype=np.float64, copy=True)
        t = np.array(terminal, dtype=np.uint8, copy=True)
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        cur = np.ascontiguousarray(cur_xy, dtype=np.float64)
        cb = None if cost_below is None else np.ascontiguousarray(cost_below, dtype=np.float64)
        off = None if view_off is None else np.ascontiguousarray(view_off, dtype=np.int32)
        vxy = None if view_xy is None else np.ascontiguousarray(view_xy, dtype=np.float64).reshape(-1, 2)
        self._L.orc_update_tree_views(self._h, C.byref(params), C.byref(updater), n, _ptr(parent), _ptr(end_xy), _ptr(sxy), _ptr(g), _ptr(cost),
                                      _ptr(v), _ptr(t), _ptr(cur), int(bool(path_changed)), int(variant), _ptr(cb), _ptr(off), _ptr(vxy))
        return g, v, t


def sample_viewpoints(times_ns, sampling_time):
    """SensorModelPlanExp::sampleViewpoints (sensor_model_planexp.cpp:65-90): indices of the trajectory points a segment is viewed from"""
    t = np.ascontiguousarray(times_ns, dtype=np.int64)
    out = np.zeros(max(1, len(t)), np.int32)
    k = lib().orc_sample_viewpoints(_ptr(t), len(t), float(sampling_time), _ptr(out), len(out))
    return out[:k].copy()


def value_select(gain, cost, parent=None):
    gain = np.ascontiguousarray(gain, dtype=np.float64)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    n = gain.shape[0]
    if parent is not None:
        parent = np.ascontiguousarray(parent, dtype=np.int32)
    value = np.zeros(n, np.float64)
    bc, bv = C.c_int32(), C.c_double()
    lib().orc_value_select(n, _ptr(gain), _ptr(cost), _ptr(parent), _ptr(value), C.byref(bc), C.byref(bv))
    return value, bc.value, bv.value
