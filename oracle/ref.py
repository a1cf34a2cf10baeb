"""ctypes front-end of oracle/_ref/libsipp_ref.so: the reference's OWN evaluator / updater / cost / map-adapter sources, compiled
unmodified from /root/reference by `make -C oracle _ref` against stand-in headers (oracle/ref_shim/) and wrapped by ref_driver.cpp.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__, bench.py's cpu_baseline leg). It pins the restated oracle (sipp_oracle.cpp): the two
must agree bit for bit on gains, terminal flags, unknown-voxel counts, updater passes and trajectory costs (tests/test_ref_pin.py).
The library is built where /root/reference exists (this container) and travels to the GPU box as a prebuilt file."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libsipp_ref.so")
REFERENCE_ROOT = "/root/reference"
_lib = None


def build():
    """(re)build when the reference tree is present; returns whether the library exists afterwards."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "advanced_planner_modules")):
        subprocess.check_call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def available():
    return os.path.exists(LIB_PATH) or build()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libsipp_ref.so is missing and /root/reference is not here to build it")
        L = C.CDLL(LIB_PATH)
        vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
        L.ref_create.argtypes = [C.POINTER(orc.MapDesc)]
        L.ref_create.restype = vp
        L.ref_destroy.argtypes = [vp]
        L.ref_destroy.restype = None
        L.ref_set_planes.argtypes = [vp, vp, vp, vp, vp, vp, f64, f64]
        L.ref_visible_voxels.argtypes = [vp, i32, f64, f64, vp, i32]
        L.ref_gain_batch.argtypes = [vp, C.POINTER(orc.EvalParams), i32, vp, vp, vp, vp, vp, f64, i32, vp, vp]
        L.ref_gain_batch_mt.argtypes = [vp, C.POINTER(orc.EvalParams), i32, vp, vp, vp, vp, vp, i32]
        L.ref_update_tree.argtypes = [vp, C.POINTER(orc.EvalParams), C.POINTER(orc.UpdaterParams), i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, C.POINTER(orc.CostParams)]
        L.ref_cost_integral.argtypes = [vp, C.POINTER(orc.CostParams), i32, vp, vp, vp, vp, vp]
        L.ref_has_module.argtypes = [C.c_char_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def has_module(name):
    return bool(lib().ref_has_module(name.encode()))


class Reference:
    """The reference's modules over a snapshot of the map planes (taken from an oracle context or given as arrays)."""

    def __init__(self, desc):
        self._L = lib()
        self.desc = desc
        self.W, self.H = desc.W, desc.H
        self._h = self._L.ref_create(C.byref(desc))

    def close(self):
        if self._h:
            self._L.ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_planes(self, state, cost, reach, path, closest=None, mean_cost=0.0, max_cost=0.0):
        state = np.ascontiguousarray(state, np.uint8)
        cost = np.ascontiguousarray(cost, np.int32)
        reach = np.ascontiguousarray(reach, np.uint8)
        path = np.ascontiguousarray(path, np.uint8)
        closest = None if closest is None else np.ascontiguousarray(closest, np.float64)
        self._L.ref_set_planes(self._h, _ptr(state), _ptr(cost), _ptr(closest), _ptr(reach), _ptr(path), float(mean_cost), float(max_cost))

    def snapshot(self, oracle_ctx):
        """copy the map state of an oracle context (map server planes, mean / max cost, current path mask)"""
        state, cost, reach, path = oracle_ctx.map_download()
        cc, _ = oracle_ctx.closest_download()
        mean, mx, _ = oracle_ctx.map_stats()
        self.set_planes(state, cost, reach, path, cc, mean, mx)
        return self

    def visible_voxels(self, half_fov, x, y):
        cap = (2 * half_fov + 1) ** 2 + 1
        out = np.zeros((cap, 3), np.float64)
        n = self._L.ref_visible_voxels(self._h, int(half_fov), float(x), float(y), _ptr(out), cap)
        return out[:max(n, 0)].copy()

    def gain_batch(self, params, xy, start_xy=None, sampling_time=0.0, traj_xy=None, times_ns=None):
        xy = np.ascontiguousarray(xy, np.float64)
        n = len(xy)
        start_xy = None if start_xy is None else np.ascontiguousarray(start_xy, np.float64)
        gain = np.zeros(n, np.float64)
        unk = np.zeros(n, np.uint32)
        term = np.zeros(n, np.uint8)
        npts = 0
        if traj_xy is not None:
            traj_xy = np.ascontiguousarray(traj_xy, np.float64)
            npts = traj_xy.shape[1]
            times_ns = np.ascontiguousarray(times_ns, np.int64)
        rc = self._L.ref_gain_batch(self._h, C.byref(params), n, _ptr(xy), _ptr(start_xy), _ptr(gain), _ptr(unk), _ptr(term), float(sampling_time),
                                    npts, _ptr(traj_xy), _ptr(times_ns))
        if rc != 0:
            raise RuntimeError("ref_gain_batch: evaluator %d is not part of the reference tree" % params.evaluator)
        return gain, unk, term

    def cycle(self, params, xy, cost, start_xy=None, threads=1):
        """gain of every candidate through the reference's own evaluator (worker threads with one module instance each), then the flat-tree
        value gain / cost and the first maximum — what the oracle's cycle() does; returns (gain, terminal, value, best_index, best_value)"""
        xy = np.ascontiguousarray(xy, np.float64)
        n = len(xy)
        start_xy = None if start_xy is None else np.ascontiguousarray(start_xy, np.float64)
        gain = np.zeros(n, np.float64)
        term = np.zeros(n, np.uint8)
        rc = self._L.ref_gain_batch_mt(self._h, C.byref(params), n, _ptr(xy), _ptr(start_xy), _ptr(gain), None, _ptr(term), int(threads))
        if rc != 0:
            raise RuntimeError("ref_gain_batch_mt failed (%d)" % rc)
        cost = np.asarray(cost, np.float64)
        value = np.where(cost > 0, (0.0 + gain) / np.where(cost > 0, cost, 1.0), 0.0)
        best = int(np.argmax(value)) if n else -1
        return gain, term, value, best, (float(value[best]) if n else 0.0)

    def update_tree(self, params, updater, parent, end_xy, gain, cost, value, terminal, cur_xy, path_changed, start_xy=None, cost_params=None):
        """one pass of the patched planner's tree re-evaluation; returns (gain, value, terminal, cost). cost_params: the cost computer
        is CostPlanExpMapIntegral (needed when the updater has update_cost set)"""
        n = len(parent)
        parent = np.ascontiguousarray(parent, np.int32)
        end_xy = np.ascontiguousarray(end_xy, np.float64)
        start_xy = None if start_xy is None else np.ascontiguousarray(start_xy, np.float64)
        g = np.array(gain, np.float64)
        v = np.array(value, np.float64)
        t = np.array(terminal, np.uint8)
        cost = np.array(cost, np.float64)
        cur = np.array(cur_xy, np.float64)
        rc = self._L.ref_update_tree(self._h, C.byref(params), C.byref(updater), n, _ptr(parent), _ptr(end_xy), _ptr(start_xy), _ptr(g), _ptr(cost),
                                     _ptr(v), _ptr(t), _ptr(cur), int(bool(path_changed)), None if cost_params is None else C.byref(cost_params))
        if rc != 0:
            raise RuntimeError("ref_update_tree failed (%d)" % rc)
        return g, v, t, cost

    def cost_integral(self, cparams, point_offset, points_xyz, parent_cost=None):
        off = np.ascontiguousarray(point_offset, np.int32)
        pts = np.ascontiguousarray(points_xyz, np.float64)
        n = len(off) - 1
        pc = None if parent_cost is None else np.ascontiguousarray(parent_cost, np.float64)
        cost = np.zeros(n, np.float64)
        valid = np.zeros(n, np.uint8)
        self._L.ref_cost_integral(self._h, C.byref(cparams), n, _ptr(off), _ptr(pts), _ptr(pc), _ptr(cost), _ptr(valid))
        return cost, valid
