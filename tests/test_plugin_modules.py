"""The C++ module mirror (scouting-ipp_b200/plugin): builds against the interface shim, registers the reference-style
`type:` strings, and — on a GPU — reproduces the oracle's literal emulation of the reference's per-node update pass.
The checks themselves are C++ (tests/cpp/test_modules.cpp), like a test of the reference's modules would be."""
import os
import subprocess

import pytest

import sipp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def test_binary():
    sipp.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "scouting-ipp_b200", "plugin")], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "tests", "cpp", "test_modules")


def _run(binary, mode):
    p = subprocess.run([binary, mode], capture_output=True, text=True, timeout=600)
    print(p.stdout[-4000:], p.stderr[-2000:])
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-2000:]
    return p.stdout


def test_modules_host_logic(test_binary):
    out = _run(test_binary, "cpu")
    assert "cpu: 0 failure(s)" in out


def test_cfg_files_mirror_reference_keys():
    """every shipped GPU cfg selects only registered Gpu* module types and keeps the reference's key names"""
    import yaml
    cfg_dir = os.path.join(ROOT, "scouting-ipp_b200", "plugin", "cfg")
    src = open(os.path.join(ROOT, "scouting-ipp_b200", "plugin", "sipp_modules.cpp")).read()
    files = sorted(f for f in os.listdir(cfg_dir) if f.endswith(".yaml"))
    assert len(files) >= 5
    for f in files:
        y = yaml.safe_load(open(os.path.join(cfg_dir, f)))
        assert y["map"]["type"] == "GpuMapPlanExp"
        fe = y["trajectory_evaluator"]["following_evaluator"]
        assert fe["type"].startswith("Gpu") and ('"%s"' % fe["type"]) in src, fe["type"]
        up = fe["evaluator_updater"]
        assert up["type"] in ("RRTStarUpdater", "ConstrainedUpdater", "UpdateAll", "UpdateAllNonTerminal")
        leaf = up.get("following_updater", up)
        assert {"update_gain", "update_cost", "update_value"} <= set(leaf)


@pytest.mark.gpu
def test_modules_parity_on_gpu(test_binary):
    out = _run(test_binary, "gpu")
    assert "gpu: 0 failure(s)" in out


@pytest.mark.gpu
def test_cpp_host_threads_drive_sharded_contexts(test_binary):
    """tests/cpp/test_shard_threads.cpp: a C++ host with one std::thread per rank (contexts of one process, one per visible
    device) scores shards with sipp_cycle_shard; the ranks meet in the kernels' peer-memory exchange; winners match the oracle."""
    import torch
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "test_shard_threads"], stdout=subprocess.DEVNULL)
    binary = os.path.join(ROOT, "tests", "cpp", "test_shard_threads")
    ndev = max(1, torch.cuda.device_count())
    for ranks in sorted({1, min(2, ndev), min(4, ndev)}):      # one rank per device (the blocking host-pointer calls need that)
        p = subprocess.run([binary, str(ranks), str(ranks)], capture_output=True, text=True, timeout=300)
        print(p.stdout[-2000:], p.stderr[-2000:])
        assert p.returncode == 0 and "0 failure(s)" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
