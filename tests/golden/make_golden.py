#!/usr/bin/env python
"""Regenerate tests/golden/ref_maps.npz + ref_maps.json from the reference's own fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs  (reference fixtures, read-only):
    docker/data/maps/map_*.png            640x480 BGR label images
    docker/data/maps/cfg/map_*.yaml       map cfg incl. `optimal_prm_rov_path_cost` (line 15) = golden PRM* cost
Outputs (committed):
    ref_maps.npz   labels[name] uint8 (H=480, W=640), decoded with the reference's label codec
                   (planexp_base_planners/include/planexp_base_planners/common/utils.h:87-112,
                    helper_scripts/map_conversion_functions.py:6-15)
    ref_maps.json  per-map cfg: resolution, width/height, ids, start/goal, golden PRM*/RRT* costs
"""
import glob
import json
import os
import sys

import cv2
import numpy as np
import yaml

REF = "/root/reference/docker/data/maps"
OUT = os.path.dirname(os.path.abspath(__file__))


def decode_labels(png_path, num_classes=15):
    """utils.h:87-112: img(float32) * float32?(num_classes/255.0) -> std::round -> B + 15 G + 225 R.

    cv::Mat::convertTo(CV_32FC3, step) multiplies in double then saturate_casts to float; std::round is
    half-away-from-zero. For uint8 inputs c*15/255 = c/17 never lands on .5 (17 is odd), so the rounding
    mode cannot matter; we still use half-away to mirror the source.
    """
    img = cv2.imread(png_path, cv2.IMREAD_COLOR)  # BGR uint8, like cv::imread default
    assert img is not None, png_path
    step = num_classes / 255.0
    f = (img.astype(np.float64) * step).astype(np.float32)
    r = np.floor(np.abs(f) + np.float32(0.5)) * np.sign(f)  # std::round on float
    r = r.astype(np.int32)
    return r[:, :, 0] + r[:, :, 1] * num_classes + r[:, :, 2] * num_classes ** 2


def main():
    cfgs = {}
    arrays = {}
    for png in sorted(glob.glob(os.path.join(REF, "map_*.png"))):
        name = os.path.splitext(os.path.basename(png))[0]
        with open(os.path.join(REF, "cfg", name + ".yaml")) as fh:
            cfg = yaml.safe_load(fh.read().replace("\t", " "))  # the shipped cfgs use tabs before comments
        lab = decode_labels(png, cfg["n_data_classes"])
        assert lab.shape == (cfg["resolution"][1], cfg["resolution"][0]), lab.shape
        assert lab.max() <= 255 and lab.min() >= 0
        assert sorted(np.unique(lab).tolist()) == sorted(set(np.unique(lab).tolist()) & set(cfg["present_ids"]))
        arrays[name] = lab.astype(np.uint8)
        cfgs[name] = {
            "W": cfg["resolution"][0], "H": cfg["resolution"][1],
            "width": cfg["width"], "height": cfg["height"],
            "n_data_classes": cfg["n_data_classes"],
            "present_ids": cfg["present_ids"],
            "obstacle_ids_mav": cfg["obstacle_ids_mav"],
            "obstacle_ids_rov": cfg["obstacle_ids_rov"],
            "unknown_id": cfg["unknown_id"],
            "start": cfg["start_location"], "goal": cfg["goal_location"],
            "golden_prm_cost": cfg["optimal_prm_rov_path_cost"],
            "golden_rrt_cost": cfg["optimal_rrt_rov_path_cost"],
            "source": f"docker/data/maps/cfg/{name}.yaml:15",
        }
    np.savez_compressed(os.path.join(OUT, "ref_maps.npz"), **arrays)
    with open(os.path.join(OUT, "ref_maps.json"), "w") as fh:
        json.dump(cfgs, fh, indent=1, sort_keys=True)
    print("wrote", len(arrays), "maps;", os.path.getsize(os.path.join(OUT, "ref_maps.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
