"""ctypes binding of libsipp.so — the B200-native scouting-ipp evaluation hot path.

This module is plumbing for tests and bench.py: it loads the C-ABI shared library declared in include/sipp.h and
exposes it as `Context`. There is no Python or CPU implementation of any result in here; if the library or a CUDA
device is missing every compute call raises `SippError`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("SIPP_LIB_PATH") or os.path.join(_PKG, "libsipp.so")   # the override is for A/B runs of two builds

MAX_IDS = 32
COUNTS_PER_CANDIDATE = 4
EVAL_RRT_STAR, EVAL_GOAL_DISTANCE, EVAL_EXPAND_REACHABLE, EVAL_AVERAGE_VIEW_COST, EVAL_EXPAND_LOW_COST, EVAL_NAIVE = range(6)
STATE_OCCUPIED, STATE_FREE, STATE_UNKNOWN = 0, 1, 2
PLAN_VALID, PLAN_INVALID, PLAN_TERMINAL_VALID, PLAN_TERMINAL_INVALID = 0, 1, 2, 3
OK, ERR_INVALID_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_PATH_MASK, ERR_ALLOC = 0, -1, -2, -3, -4, -5

# every symbol include/sipp.h declares (tests check that the library exports all of them)
ABI_SYMBOLS = [
    "sipp_abi_version", "sipp_create", "sipp_destroy", "sipp_last_error", "sipp_stream", "sipp_map_clear",
    "sipp_patch_origin", "sipp_map_update_patch", "sipp_map_upload_full", "sipp_update_unknown_cost", "sipp_map_stats",
    "sipp_map_download", "sipp_plan", "sipp_field_download", "sipp_path_mask_generation", "sipp_set_path_mask",
    "sipp_plan_stats", "sipp_gain_batch", "sipp_gain_batch_dev", "sipp_value_select", "sipp_value_select_dev", "sipp_value_preorder",
    "sipp_cycle", "sipp_cycle_dev", "sipp_merge_best_dev", "sipp_update_tree", "sipp_launch_count",
    "sipp_peer_export", "sipp_peer_attach", "sipp_peer_detach", "sipp_peer_status", "sipp_cycle_shard_dev", "sipp_cycle_shard",
    "sipp_peer_exchange_dev", "sipp_cycle_shard_pipelined_dev", "sipp_peer_flush_dev", "sipp_episode_run", "sipp_cost_integral", "sipp_set_incremental", "sipp_plan_info",
    "sipp_path_download", "sipp_truth_upload", "sipp_travel_cost", "sipp_plan_phases",
    "sipp_batch_create", "sipp_batch_destroy", "sipp_batch_last_error", "sipp_batch_launch_count", "sipp_batch_set_sharing", "sipp_batch_ingest", "sipp_batch_plan",
    "sipp_batch_cycle", "sipp_batch_episode_run", "sipp_gain_batch_views", "sipp_update_tree_views",
]
PEER_MAX_WORLD, PEER_HANDLE_BYTES = 16, 128


class MapDesc(C.Structure):  # sipp_map_desc
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


class EvalParams(C.Structure):  # sipp_eval_params
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


class SippError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sipp error %d: %s" % (code, msg))
        self.code = code


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


def build(force=False, verbose=False):
    """Compile libsipp.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(_PKG, "csrc", f) for f in os.listdir(os.path.join(_PKG, "csrc"))] + [
        os.path.join(os.path.dirname(_PKG), "include", "sipp.h")]
    newest = max(os.path.getmtime(f) for f in srcs if f.endswith((".cu", ".h", ".cuh")))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        out = None if verbose else subprocess.DEVNULL
        subprocess.check_call(["make", "-C", _PKG, "-j8", "libsipp.so"] + (["-B"] if force else []), stdout=out, stderr=out)
    return LIB_PATH


_lib = None


def lib():
    """Load libsipp.so. Raises OSError when it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("libsipp.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, i32, f64, i64 = C.c_void_p, C.c_int32, C.c_double, C.c_int64
        pi32, pf64 = C.POINTER(i32), C.POINTER(f64)
        L.sipp_abi_version.restype = C.c_int
        L.sipp_create.argtypes = [C.POINTER(MapDesc), C.c_int, C.POINTER(vp)]
        L.sipp_destroy.argtypes = [vp]
        L.sipp_destroy.restype = None
        L.sipp_last_error.argtypes = [vp]
        L.sipp_last_error.restype = C.c_char_p
        L.sipp_stream.argtypes = [vp]
        L.sipp_stream.restype = vp
        L.sipp_map_clear.argtypes = [vp]
        L.sipp_patch_origin.argtypes = [vp, f64, f64, i32, i32, pi32, pi32]
        L.sipp_map_update_patch.argtypes = [vp, i32, i32, i32, i32, vp]
        L.sipp_map_upload_full.argtypes = [vp, vp, vp]
        L.sipp_update_unknown_cost.argtypes = [vp, f64]
        L.sipp_map_stats.argtypes = [vp, pf64, pf64, C.POINTER(i64)]
        L.sipp_map_download.argtypes = [vp, vp, vp, vp, vp]
        L.sipp_plan.argtypes = [vp, pf64, pi32, vp, i32, pi32]
        L.sipp_field_download.argtypes = [vp, vp]
        L.sipp_path_download.argtypes = [vp, vp, i32, pi32]
        L.sipp_truth_upload.argtypes = [vp, vp]
        L.sipp_plan_phases.argtypes = [vp, vp]
        L.sipp_travel_cost.argtypes = [vp, C.POINTER(TravelParams), i32, vp, vp, vp]
        L.sipp_path_mask_generation.argtypes = [vp]
        L.sipp_set_path_mask.argtypes = [vp, vp]
        L.sipp_plan_stats.argtypes = [vp, vp]
        L.sipp_gain_batch.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp]
        L.sipp_gain_batch_dev.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp]
        L.sipp_value_select.argtypes = [vp, i32, vp, vp, vp, vp, pi32, pf64]
        L.sipp_value_select_dev.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.sipp_value_preorder.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.sipp_cycle.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp, pi32, pf64]
        L.sipp_cycle_dev.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp, vp, vp]
        L.sipp_merge_best_dev.argtypes = [vp, vp, i32, i64, vp, vp]
        L.sipp_update_tree.argtypes = [vp, C.POINTER(EvalParams), C.POINTER(UpdaterParams), i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, pi32, vp]
        L.sipp_update_tree_views.argtypes = [vp, C.POINTER(EvalParams), C.POINTER(UpdaterParams), i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, pi32, vp, vp, vp]
        L.sipp_gain_batch_views.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp]
        L.sipp_peer_export.argtypes = [vp, vp]
        L.sipp_peer_attach.argtypes = [vp, i32, i32, vp]
        L.sipp_peer_detach.argtypes = [vp]
        L.sipp_peer_status.argtypes = [vp, pi32, C.POINTER(i64)]
        L.sipp_cycle_shard_dev.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp, i64, vp, vp]
        L.sipp_cycle_shard.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, i64, vp, vp, vp, C.POINTER(i64), pf64]
        L.sipp_peer_exchange_dev.argtypes = [vp, vp, vp, vp]
        L.sipp_cycle_shard_pipelined_dev.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp, vp, i64, vp, vp]
        L.sipp_peer_flush_dev.argtypes = [vp, vp, vp]
        L.sipp_episode_run.argtypes = [vp, vp, C.POINTER(EvalParams), i32, i32, vp, vp, i32, i32, f64, f64, vp, vp, vp, vp, vp]
        L.sipp_cost_integral.argtypes = [vp, C.POINTER(CostParams), i32, vp, vp, vp, vp, vp]
        L.sipp_batch_create.argtypes = [vp, i32, C.POINTER(vp)]
        L.sipp_batch_destroy.argtypes = [vp]
        L.sipp_batch_destroy.restype = None
        L.sipp_batch_last_error.argtypes = [vp]
        L.sipp_batch_last_error.restype = C.c_char_p
        L.sipp_batch_launch_count.argtypes = [vp]
        L.sipp_batch_launch_count.restype = C.c_int64
        L.sipp_batch_set_sharing.argtypes = [vp, C.c_int32]
        L.sipp_batch_ingest.argtypes = [vp, vp, vp, i32, i32, vp]
        L.sipp_batch_plan.argtypes = [vp, vp, vp, vp]
        L.sipp_batch_cycle.argtypes = [vp, C.POINTER(EvalParams), i32, vp, vp, vp, vp, vp]
        L.sipp_batch_episode_run.argtypes = [vp, vp, C.POINTER(EvalParams), i32, i32, vp, vp, i32, i32, f64, f64, vp, vp, vp, vp, vp]
        L.sipp_set_incremental.argtypes = [vp, i32]
        L.sipp_plan_info.argtypes = [vp, vp]
        L.sipp_launch_count.argtypes = [vp]
        L.sipp_launch_count.restype = i64
        _lib = L
    return _lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One map + follower planner + evaluator state resident on one GPU (sipp_ctx)."""

    def __init__(self, desc, device=0):
        self.desc = desc
        self.W, self.H = desc.W, desc.H
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.sipp_create(C.byref(desc), int(device), C.byref(h))
        if rc != OK:
            raise SippError(rc, self._L.sipp_last_error(None).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.sipp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise SippError(rc, self._L.sipp_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    @property
    def stream(self):
        return self._L.sipp_stream(self._h)

    def launch_count(self):
        return int(self._L.sipp_launch_count(self._h))

    # ---- map store
    def map_clear(self):
        self._check(self._L.sipp_map_clear(self._h))

    def patch_origin(self, px, py, rows, cols):
        x0, y0 = C.c_int32(), C.c_int32()
        self._check(self._L.sipp_patch_origin(self._h, px, py, rows, cols, C.byref(x0), C.byref(y0)))
        return x0.value, y0.value

    def map_update_patch(self, x0, y0, labels):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        rows, cols = labels.shape
        self._check(self._L.sipp_map_update_patch(self._h, int(x0), int(y0), rows, cols, _ptr(labels)))

    def map_upload_full(self, labels, observed=None):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        assert labels.shape == (self.H, self.W)
        if observed is not None:
            observed = np.ascontiguousarray(observed, dtype=np.uint8)
            assert observed.shape == (self.H, self.W)
        self._check(self._L.sipp_map_upload_full(self._h, _ptr(labels), _ptr(observed)))

    def update_unknown_cost(self, c):
        self._check(self._L.sipp_update_unknown_cost(self._h, float(c)))

    def map_stats(self):
        m, mx, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self._L.sipp_map_stats(self._h, C.byref(m), C.byref(mx), C.byref(n)))
        return m.value, mx.value, n.value

    def map_download(self):
        state = np.empty((self.H, self.W), np.uint8)
        cost = np.empty((self.H, self.W), np.int32)
        reach = np.empty((self.H, self.W), np.uint8)
        path = np.empty((self.H, self.W), np.uint8)
        self._check(self._L.sipp_map_download(self._h, _ptr(state), _ptr(cost), _ptr(reach), _ptr(path)))
        return state, cost, reach, path

    # ---- follower path cost
    def plan(self, path_cap=None):
        cap = path_cap if path_cap is not None else 4 * (self.W + self.H)
        cost, status, n = C.c_double(), C.c_int32(), C.c_int32()
        path = np.zeros((cap, 2), np.float64)
        self._check(self._L.sipp_plan(self._h, C.byref(cost), C.byref(status), _ptr(path), cap, C.byref(n)))
        return cost.value, status.value, path[: min(n.value, cap)].copy(), n.value

    def path_download(self, cap=None):
        """the path of the last plan again (all of it by default)"""
        n = C.c_int32()
        self._check(self._L.sipp_path_download(self._h, None, 0, C.byref(n)))
        cap = n.value if cap is None else cap
        path = np.zeros((max(cap, 1), 2), np.float64)
        self._check(self._L.sipp_path_download(self._h, _ptr(path), cap, C.byref(n)))
        return path[: min(n.value, cap)].copy(), n.value

    def plan_stats(self):
        st = np.zeros(4, np.int64)
        self._check(self._L.sipp_plan_stats(self._h, _ptr(st)))
        return dict(pushes=int(st[0]), tile_activations=int(st[1]), tile_sweeps=int(st[2]), aborted=int(st[3]))

    def plan_phases(self):
        """device-timed (relaxation kernel ms, whole plan ms) of the last plan"""
        ms = np.zeros(2, np.float64)
        self._check(self._L.sipp_plan_phases(self._h, _ptr(ms)))
        return float(ms[0]), float(ms[1])

    def set_incremental(self, enable):
        self._check(self._L.sipp_set_incremental(self._h, int(bool(enable))))

    def plan_info(self):
        st = np.zeros(4, np.int64)
        self._check(self._L.sipp_plan_info(self._h, _ptr(st)))
        return dict(incremental=bool(st[0]), patch_tiles=int(st[1]), relax_seed_tiles=int(st[2]), invalidated_cells=int(st[3]))

    def field_download(self):
        f = np.empty((self.H, self.W), np.float64)
        self._check(self._L.sipp_field_download(self._h, _ptr(f)))
        return f

    def path_mask_generation(self):
        return self._L.sipp_path_mask_generation(self._h)

    def set_path_mask(self, mask):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        assert mask.shape == (self.H, self.W)
        self._check(self._L.sipp_set_path_mask(self._h, _ptr(mask)))

    # ---- gain / value / cycle (host buffers)
    def gain_batch(self, params, xy, start_xy=None):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        n = xy.shape[0]
        if start_xy is not None:
            start_xy = np.ascontiguousarray(start_xy, dtype=np.float64).reshape(-1, 2)
        gain = np.zeros(n, np.float64)
        counts = np.zeros((n, COUNTS_PER_CANDIDATE), np.uint32)
        terminal = np.zeros(n, np.uint8)
        self._check(self._L.sipp_gain_batch(self._h, C.byref(params), n, _ptr(xy), _ptr(start_xy), _ptr(gain), _ptr(counts),
                                            _ptr(terminal)))
        return gain, counts, terminal

    def gain_batch_views(self, params, view_off, view_xy, start_xy=None):
        """segments scored from several sampled viewpoints each: segment i has the views view_xy[view_off[i]:view_off[i+1]]"""
        off = np.ascontiguousarray(view_off, dtype=np.int32)
        n = off.shape[0] - 1
        vxy = np.ascontiguousarray(view_xy, dtype=np.float64).reshape(-1, 2)
        sxy = None if start_xy is None else np.ascontiguousarray(start_xy, dtype=np.float64).reshape(n, 2)
        gain = np.zeros(n, np.float64)
        counts = np.zeros((n, COUNTS_PER_CANDIDATE), np.uint32)
        terminal = np.zeros(n, np.uint8)
        self._check(self._L.sipp_gain_batch_views(self._h, C.byref(params), n, _ptr(off), _ptr(vxy), _ptr(sxy), _ptr(gain), _ptr(counts), _ptr(terminal)))
        return gain, counts, terminal

    def value_select(self, gain, cost, parent=None):
        gain = np.ascontiguousarray(gain, dtype=np.float64)
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        n = gain.shape[0]
        if parent is not None:
            parent = np.ascontiguousarray(parent, dtype=np.int32)
        value = np.zeros(n, np.float64)
        bc, bv = C.c_int32(), C.c_double()
        self._check(self._L.sipp_value_select(self._h, n, _ptr(gain), _ptr(cost), _ptr(parent), _ptr(value), C.byref(bc),
                                              C.byref(bv)))
        return value, bc.value, bv.value

    def value_preorder(self, gain_new, gain_old, cost, parent):
        gain_new = np.ascontiguousarray(gain_new, dtype=np.float64)
        gain_old = np.ascontiguousarray(gain_old, dtype=np.float64)
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        value = np.zeros(gain_new.shape[0], np.float64)
        self._check(self._L.sipp_value_preorder(self._h, gain_new.shape[0], _ptr(gain_new), _ptr(gain_old), _ptr(cost), _ptr(parent),
                                                _ptr(value)))
        return value

    def cycle(self, params, xy, cost, start_xy=None):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        n = xy.shape[0]
        cost = np.ascontiguousarray(cost, dtype=np.float64)
        if start_xy is not None:
            start_xy = np.ascontiguousarray(start_xy, dtype=np.float64).reshape(-1, 2)
        gain = np.zeros(n, np.float64)
        terminal = np.zeros(n, np.uint8)
        value = np.zeros(n, np.float64)
        bi, bv = C.c_int32(), C.c_double()
        self._check(self._L.sipp_cycle(self._h, C.byref(params), n, _ptr(xy), _ptr(start_xy), _ptr(cost), _ptr(gain),
                                       _ptr(terminal), _ptr(value), C.byref(bi), C.byref(bv)))
        return gain, terminal, value, bi.value, bv.value

    # ---- raw-pointer variants (device pointers as ints, e.g. torch.Tensor.data_ptr(); stream as int or None)
    def gain_batch_dev(self, params, n, d_xy, d_start_xy, d_gain, d_counts, d_terminal, stream=None):
        self._check(self._L.sipp_gain_batch_dev(self._h, C.byref(params), int(n), _ptr(d_xy), _ptr(d_start_xy), _ptr(d_gain),
                                                _ptr(d_counts), _ptr(d_terminal), _ptr(stream)))

    def cycle_dev(self, params, n, d_xy, d_start_xy, d_cost, d_gain, d_terminal, d_value, d_best, stream=None):
        self._check(self._L.sipp_cycle_dev(self._h, C.byref(params), int(n), _ptr(d_xy), _ptr(d_start_xy), _ptr(d_cost),
                                           _ptr(d_gain), _ptr(d_terminal), _ptr(d_value), _ptr(d_best), _ptr(stream)))

    def update_tree(self, params, updater, parent, end_xy, gain, cost, value, terminal, cur_xy, path_changed, start_xy=None, cost_below=None,
                    view_off=None, view_xy=None):
        """one re-evaluation pass over a pre-order tree; returns (gain, value, terminal, n_rescored) — inputs are not modified"""
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
        nres = C.c_int32()
        if view_off is not None:        # node i is scored from its sampled viewpoints (sensor_model/sampling_time > 0)
            off = np.ascontiguousarray(view_off, dtype=np.int32)
            vxy = np.ascontiguousarray(view_xy, dtype=np.float64).reshape(-1, 2)
            self._check(self._L.sipp_update_tree_views(self._h, C.byref(params), C.byref(updater), n, _ptr(parent), _ptr(end_xy), _ptr(sxy), _ptr(g),
                                                       _ptr(cost), _ptr(v), _ptr(t), _ptr(cur), int(bool(path_changed)), C.byref(nres), _ptr(cb),
                                                       _ptr(off), _ptr(vxy)))
            return g, v, t, nres.value
        self._check(self._L.sipp_update_tree(self._h, C.byref(params), C.byref(updater), n, _ptr(parent), _ptr(end_xy), _ptr(sxy), _ptr(g),
                                             _ptr(cost), _ptr(v), _ptr(t), _ptr(cur), int(bool(path_changed)), C.byref(nres), _ptr(cb)))
        return g, v, t, nres.value

    def merge_best_dev(self, d_recs, world, shard_stride, d_out, stream=None):
        self._check(self._L.sipp_merge_best_dev(self._h, _ptr(d_recs), int(world), int(shard_stride), _ptr(d_out), _ptr(stream)))

    def cost_integral(self, cparams, point_offset, points_xyz, parent_cost=None):
        """CostPlanExpMapIntegral for a batch of segments: returns (cost, valid)"""
        off = np.ascontiguousarray(point_offset, dtype=np.int32)
        n = off.shape[0] - 1
        pts = np.ascontiguousarray(points_xyz, dtype=np.float64).reshape(-1, 3)
        pc = None if parent_cost is None else np.ascontiguousarray(parent_cost, dtype=np.float64)
        cost, valid = np.zeros(n, np.float64), np.zeros(n, np.uint8)
        self._check(self._L.sipp_cost_integral(self._h, C.byref(cparams), n, _ptr(off), _ptr(pts), _ptr(pc), _ptr(cost), _ptr(valid)))
        return cost, valid

    def truth_upload(self, labels):
        labels = np.ascontiguousarray(labels, dtype=np.int32)
        assert labels.shape == (self.H, self.W)
        self._check(self._L.sipp_truth_upload(self._h, _ptr(labels)))

    def travel_cost(self, tparams, start_xyz, goal_xyz):
        """MavSimulator::computeCost for n (start, goal) pairs over the uploaded ground-truth map"""
        a = np.ascontiguousarray(start_xyz, dtype=np.float64).reshape(-1, 3)
        b = np.ascontiguousarray(goal_xyz, dtype=np.float64).reshape(-1, 3)
        cost = np.zeros(len(a), np.float64)
        self._check(self._L.sipp_travel_cost(self._h, C.byref(tparams), len(a), _ptr(a), _ptr(b), _ptr(cost)))
        return cost

    def episode_run(self, truth, params, cand_xy, cand_cost, fov=65, init=320, start=(0.5, 0.5)):
        """sipp_episode_run: cand_xy [cycles][n][2], cand_cost [cycles][n]; returns the per-cycle trace as a list of
        (path cost, plan status, path nodes, selected candidate, value) — the format of sipp.episode.run_episode"""
        truth = np.ascontiguousarray(truth, dtype=np.int32)
        cand_xy = np.ascontiguousarray(cand_xy, dtype=np.float64)
        cand_cost = np.ascontiguousarray(cand_cost, dtype=np.float64)
        cycles, n = cand_cost.shape
        assert truth.shape == (self.H, self.W) and cand_xy.shape == (cycles, n, 2)
        pc, bv = np.zeros(cycles, np.float64), np.zeros(cycles, np.float64)
        st, ln, bi = np.zeros(cycles, np.int32), np.zeros(cycles, np.int32), np.zeros(cycles, np.int32)
        self._check(self._L.sipp_episode_run(self._h, _ptr(truth), C.byref(params), cycles, n, _ptr(cand_xy), _ptr(cand_cost), int(fov), int(init),
                                             float(start[0]), float(start[1]), _ptr(pc), _ptr(st), _ptr(ln), _ptr(bi), _ptr(bv)))
        return [(float(pc[c]), int(st[c]), int(ln[c]), int(bi[c]), float(bv[c])) for c in range(cycles)]

    # ---- multi-GPU: shard winners exchanged through peer memory (sipp_peer_*)
    def peer_export(self):
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        self._check(self._L.sipp_peer_export(self._h, buf))
        return bytes(buf.raw)

    def peer_attach(self, rank, handles):
        """handles: list of the `world` byte strings returned by every rank's peer_export(), in rank order"""
        blob = b"".join(handles)
        assert len(blob) == PEER_HANDLE_BYTES * len(handles)
        self._check(self._L.sipp_peer_attach(self._h, int(rank), len(handles), C.c_char_p(blob)))

    def peer_detach(self):
        self._check(self._L.sipp_peer_detach(self._h))

    def peer_status(self):
        st, n = C.c_int32(), C.c_int64()
        self._check(self._L.sipp_peer_status(self._h, C.byref(st), C.byref(n)))
        return st.value, n.value

    def cycle_shard_dev(self, params, n, d_xy, d_start_xy, d_cost, d_gain, d_terminal, d_value, index_offset, d_best_global, stream=None):
        self._check(self._L.sipp_cycle_shard_dev(self._h, C.byref(params), int(n), _ptr(d_xy), _ptr(d_start_xy), _ptr(d_cost), _ptr(d_gain),
                                                 _ptr(d_terminal), _ptr(d_value), int(index_offset), _ptr(d_best_global), _ptr(stream)))

    def cycle_shard_pipelined_dev(self, params, n, d_xy, d_start_xy, d_cost, d_gain, d_terminal, d_value, index_offset, d_prev_best_global, stream=None):
        self._check(self._L.sipp_cycle_shard_pipelined_dev(self._h, C.byref(params), int(n), _ptr(d_xy), _ptr(d_start_xy), _ptr(d_cost), _ptr(d_gain),
                                                           _ptr(d_terminal), _ptr(d_value), int(index_offset), _ptr(d_prev_best_global), _ptr(stream)))

    def peer_flush_dev(self, d_best_global, stream=None):
        self._check(self._L.sipp_peer_flush_dev(self._h, _ptr(d_best_global), _ptr(stream)))

    def cycle_shard_host_ptr(self, params, n, xy, start_xy, cost, index_offset, gain, terminal, value):
        """sipp_cycle_shard on caller-owned host buffers given as raw addresses; returns the GLOBAL winner"""
        bi, bv = C.c_int64(), C.c_double()
        self._check(self._L.sipp_cycle_shard(self._h, C.byref(params), int(n), _ptr(xy), _ptr(start_xy), _ptr(cost), int(index_offset),
                                             _ptr(gain), _ptr(terminal), _ptr(value), C.byref(bi), C.byref(bv)))
        return bi.value, bv.value

    def peer_exchange_dev(self, d_mine, d_out, stream=None):
        self._check(self._L.sipp_peer_exchange_dev(self._h, _ptr(d_mine), _ptr(d_out), _ptr(stream)))

    def cycle_host_ptr(self, params, n, xy, start_xy, cost, gain, terminal, value):
        """sipp_cycle on caller-owned host buffers given as raw addresses (pinned memory for the e2e benchmark)."""
        bi, bv = C.c_int32(), C.c_double()
        self._check(self._L.sipp_cycle(self._h, C.byref(params), int(n), _ptr(xy), _ptr(start_xy), _ptr(cost), _ptr(gain),
                                       _ptr(terminal), _ptr(value), C.byref(bi), C.byref(bv)))
        return bi.value, bv.value


class Batch:
    """m maps of one size on one device that advance together (sipp_batch): one launch per kernel for all maps."""

    def __init__(self, contexts):
        self._L = lib()
        self.ctx = list(contexts)
        self.m = len(self.ctx)
        arr = (C.c_void_p * self.m)(*[c._h for c in self.ctx])
        h = C.c_void_p()
        rc = self._L.sipp_batch_create(arr, self.m, C.byref(h))
        if rc != 0:
            raise SippError(rc, self._L.sipp_last_error(self.ctx[0]._h).decode() if self.ctx else "empty batch")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.sipp_batch_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise SippError(rc, self._L.sipp_batch_last_error(self._h).decode())

    def _ptrs(self, arrays):
        return (C.c_void_p * self.m)(*[a.ctypes.data for a in arrays])

    def launch_count(self):
        return int(self._L.sipp_batch_launch_count(self._h))

    def set_sharing(self, batches_on_this_gpu):
        """how many batches run side by side on this GPU: the persistent workers of this one take that fraction of the SMs"""
        self._check(self._L.sipp_batch_set_sharing(self._h, int(batches_on_this_gpu)))

    def ingest(self, x0, y0, patches):
        """patches: m arrays [rows][cols] int32 (one shape)"""
        ps = [np.ascontiguousarray(p, dtype=np.int32) for p in patches]
        rows, cols = ps[0].shape
        x0 = np.ascontiguousarray(x0, dtype=np.int32)
        y0 = np.ascontiguousarray(y0, dtype=np.int32)
        self._check(self._L.sipp_batch_ingest(self._h, _ptr(x0), _ptr(y0), rows, cols, self._ptrs(ps)))

    def plan(self):
        """returns m tuples (cost, status, path nodes)"""
        cost, st, ln = np.zeros(self.m), np.zeros(self.m, np.int32), np.zeros(self.m, np.int32)
        self._check(self._L.sipp_batch_plan(self._h, _ptr(cost), _ptr(st), _ptr(ln)))
        return [(float(cost[i]), int(st[i]), int(ln[i])) for i in range(self.m)]

    def cycle(self, params, xy, cost, want_gain=False):
        """xy: m arrays [n][2], cost: m arrays [n]; returns (best_index[m], best_value[m][, gain m x n])"""
        xs = [np.ascontiguousarray(a, dtype=np.float64) for a in xy]
        cs = [np.ascontiguousarray(a, dtype=np.float64) for a in cost]
        n = len(cs[0])
        bi, bv = np.zeros(self.m, np.int64), np.zeros(self.m)
        gains = [np.zeros(n) for _ in range(self.m)] if want_gain else None
        self._check(self._L.sipp_batch_cycle(self._h, C.byref(params), n, self._ptrs(xs), self._ptrs(cs), _ptr(bi), _ptr(bv),
                                             self._ptrs(gains) if want_gain else None))
        return (bi, bv, gains) if want_gain else (bi, bv)

    def episode_run(self, truths, params, cand_xy, cand_cost, fov=65, init=320, start=(0.5, 0.5)):
        """sipp_batch_episode_run: truths m x [H][W]; cand_xy m x [cycles][n][2]; cand_cost m x [cycles][n]; returns m traces in the
        format of Context.episode_run"""
        ts = [np.ascontiguousarray(t, dtype=np.int32) for t in truths]
        xs = [np.ascontiguousarray(a, dtype=np.float64) for a in cand_xy]
        cs = [np.ascontiguousarray(a, dtype=np.float64) for a in cand_cost]
        cycles, n = cs[0].shape
        pc, bv = np.zeros((self.m, cycles)), np.zeros((self.m, cycles))
        st, ln, bi = (np.zeros((self.m, cycles), np.int32) for _ in range(3))
        self._check(self._L.sipp_batch_episode_run(self._h, self._ptrs(ts), C.byref(params), cycles, n, self._ptrs(xs), self._ptrs(cs), int(fov), int(init),
                                                   float(start[0]), float(start[1]), _ptr(pc), _ptr(st), _ptr(ln), _ptr(bi), _ptr(bv)))
        return [[(float(pc[k, c]), int(st[k, c]), int(ln[k, c]), int(bi[k, c]), float(bv[k, c])) for c in range(cycles)] for k in range(self.m)]
