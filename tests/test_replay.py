"""Saved-map replay (scouting-ipp_b200/plugin/sipp_replay.*): the loader of MapServerPlanExp::saveMap's files
(map_server_planexp.cpp:489-576, eigen_csv.h) on the CPU, and — on a GPU — the batched plan per saved map against the oracle, with
the result rows / path files in the format of planexp_base_planners/common/utils.h:367-386."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import sipp
from sipp import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def R():
    sipp.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "scouting-ipp_b200", "plugin")], stdout=subprocess.DEVNULL)
    L = C.CDLL(os.path.join(ROOT, "scouting-ipp_b200", "plugin", "libsipp_modules.so"))
    vp, i32 = C.c_void_p, C.c_int32
    L.sipp_replay_save_map.argtypes = [C.c_char_p, i32, i32, vp, vp, C.c_char_p, i32]
    L.sipp_replay_load_map.argtypes = [C.c_char_p, C.POINTER(i32), C.POINTER(i32), vp, vp, C.c_int64, C.c_char_p, i32]
    L.sipp_replay_saved_maps.argtypes = [vp, i32, C.POINTER(sipp.MapDesc), i32, i32, C.c_char_p, C.c_char_p, vp, vp, C.c_char_p, i32]
    return L


def save(R, prefix, lab, obs):
    lab = np.ascontiguousarray(lab, np.int32)
    obs = np.ascontiguousarray(obs, np.uint8)
    err = C.create_string_buffer(512)
    assert R.sipp_replay_save_map(prefix.encode(), lab.shape[1], lab.shape[0], lab.ctypes.data, obs.ctypes.data, err, 512) == 0, err.value


def load(R, prefix, cap=1 << 22):
    W, H = C.c_int32(), C.c_int32()
    lab, obs = np.zeros(cap, np.int32), np.zeros(cap, np.uint8)
    err = C.create_string_buffer(512)
    rc = R.sipp_replay_load_map(prefix.encode(), C.byref(W), C.byref(H), lab.ctypes.data, obs.ctypes.data, cap, err, 512)
    if rc != 0:
        raise RuntimeError(err.value.decode())
    n = W.value * H.value
    return lab[:n].reshape(H.value, W.value), obs[:n].reshape(H.value, W.value)


def test_saved_map_round_trip_and_reference_format(R, tmp_path):
    lab, obs, _ = synth.world(96, 64, 11, 3, sipp.make_desc)
    p = str(tmp_path / "00003")
    save(R, p, lab, obs)
    # the files are what Eigen's IOFormat(StreamPrecision, DontAlignCols, ", ", "\n") writes: ", " between cells, no trailing newline
    text = open(p + "_cost.csv").read()
    assert text.count("\n") == 63 and ", " in text and not text.endswith("\n")
    cost = np.loadtxt(p + "_cost.csv", delimiter=",").astype(np.int32)
    occ = np.loadtxt(p + "_occupancy.csv", delimiter=",").astype(np.int32)
    assert np.array_equal(cost, lab) and np.array_equal(occ == 2, obs == 0)       # MSR_UNKNOWN 2
    l2, o2 = load(R, p)
    assert np.array_equal(o2, obs) and np.array_equal(l2, np.where(obs != 0, lab, 0))
    # a file written by the reference itself would end every row but the last with "\n" as well; extra blanks and a final newline are tolerated
    open(p + "_cost.csv", "a").write("\n")
    assert np.array_equal(load(R, p)[1], obs)


def test_saved_map_loader_errors(R, tmp_path):
    with pytest.raises(RuntimeError, match="cannot open"):
        load(R, str(tmp_path / "missing"))
    p = str(tmp_path / "ragged")
    open(p + "_cost.csv", "w").write("1, 2, 3\n4, 5\n")
    open(p + "_occupancy.csv", "w").write("1, 1, 1\n1, 1, 1\n")
    with pytest.raises(RuntimeError, match="row 1 has 2 cells"):
        load(R, p)
    p = str(tmp_path / "shape")
    open(p + "_cost.csv", "w").write("1, 2, 3\n4, 5, 6\n")
    open(p + "_occupancy.csv", "w").write("1, 1\n1, 1\n")
    with pytest.raises(RuntimeError, match="occupancy map 2 x 2"):
        load(R, p)
    p = str(tmp_path / "text")
    open(p + "_cost.csv", "w").write("1, x\n")
    open(p + "_occupancy.csv", "w").write("1, 1\n")
    with pytest.raises(RuntimeError, match="not a number"):
        load(R, p)


def test_replay_reports_its_failures(R, tmp_path):
    """a missing saved map (with a GPU) or a missing GPU (here) ends in an error code and a message, never in made-up results"""
    _, _, desc = synth.world(96, 64, 11, 3, sipp.make_desc)
    arr = (C.c_char_p * 1)(str(tmp_path / "nothing_here").encode())
    cost, status = np.full(1, 123.0), np.zeros(1, np.int32)
    err = C.create_string_buffer(512)
    rc = R.sipp_replay_saved_maps(arr, 1, C.byref(desc), 0, 4, None, None, cost.ctypes.data, status.ctypes.data, err, 512)
    assert rc != 0 and err.value and cost[0] == 123.0, (rc, err.value)
    assert b"cannot open" in err.value or b"no CPU fallback" in err.value or b"CUDA" in err.value, err.value


@pytest.mark.gpu
def test_replay_plans_every_saved_map_like_the_oracle(R, tmp_path):
    from oracle import oracle as orc
    W, H, n = 640, 480, 7
    prefixes, want = [], []
    desc = None
    for k in range(n):
        lab = synth.make_labels(W, H, 4100 + k)
        obs = synth.make_observed(W, H, 4100 + k, 6 + 3 * k)
        if k == n - 1:
            obs[:] = 0
            obs[:40, :40] = 1            # the goal region was never seen: planned through unexplored cells at the init cost
        ids = sorted(set(np.unique(synth.make_labels(W, H, 4100)).tolist()) - {0})
        mk = lambda f: f(W, H, W * synth.MPP, H * synth.MPP, (0.5, 0.5), (W * synth.MPP - 0.5, H * synth.MPP - 0.5), min(ids), max(ids), max(ids))
        desc = mk(sipp.make_desc)
        p = str(tmp_path / ("%05d" % k))
        save(R, p, lab, obs)
        prefixes.append(p)
        o = orc.Oracle(mk(orc.make_desc))
        o.map_upload_full(lab, obs)
        want.append(o.plan())
    arr = (C.c_char_p * n)(*[p.encode() for p in prefixes])
    cost, status = np.zeros(n), np.zeros(n, np.int32)
    err = C.create_string_buffer(512)
    out_csv = str(tmp_path / "rrt_output.csv")
    os.mkdir(str(tmp_path / "rrt_paths"))
    rc = R.sipp_replay_saved_maps(arr, n, C.byref(desc), 0, 3, out_csv.encode(), str(tmp_path / "rrt_paths").encode(), cost.ctypes.data, status.ctypes.data, err, 512)
    assert rc == 0, err.value
    for k in range(n):
        wc, ws, wpath = want[k][0], want[k][1], want[k][2]
        assert cost[k] == wc and status[k] == ws, (k, cost[k], wc, status[k], ws)
    rows = [r.split(",") for r in open(out_csv).read().strip().splitlines()]
    assert len(rows) == n and all(len(r) == 9 for r in rows)
    for k, r in enumerate(rows):
        assert r[0] == "%05d" % k and int(r[7]) == int(cost[k] != -1.0) and float(r[8]) == pytest.approx(cost[k], rel=1e-5)
        if cost[k] != -1.0:
            lines = open(str(tmp_path / "rrt_paths" / ("%05d_path.csv" % k))).read().strip().splitlines()
            assert lines[0] == "PositionX,PositionY,LocationX,LocationY" and lines[1] == "m,m,px,px"
            wpath = np.asarray(want[k][2]).reshape(-1, 2)
            got = np.array([[float(v) for v in ln.split(",")[:2]] for ln in lines[2:]])
            assert got.shape == wpath.shape and np.allclose(got, wpath, rtol=1e-5)
