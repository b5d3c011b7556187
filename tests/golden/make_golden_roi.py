#!/usr/bin/env python
"""Golden vectors for the RoI input pipeline, produced by the REFERENCE'S OWN functions (tools/dataset_utils.py:
get_2d_coord_np, crop_resize_by_warp_affine, get_affine_transform) running on the cv2 of the build image, following
evaluation/load_data_eval.py:256-289.  Run in the build container (needs /root/reference and cv2); writes tests/golden/roi.npz."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, "/root/reference")
from tools.dataset_utils import crop_resize_by_warp_affine, get_2d_coord_np, get_affine_transform  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
rng = np.random.default_rng(20261017)
H, W = 120, 160
image = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
inst = rng.integers(0, 4, (H, W), dtype=np.uint8)           # instance-id map (training-style mask)
coord_2d = get_2d_coord_np(W, H).transpose(1, 2, 0)
# (cx, cy, scale, img_size, out_res, inst_id): inside, partly outside, fully outside, half-integer centres, scale = n * 1.5
cases = [(80.0, 60.0, 90.0, 64, 16, -1), (10.5, 7.0, 75.0, 64, 16, 2), (150.25, 110.75, 120.0, 64, 16, 1), (79.5, 59.5, 160.0, 64, 16, -1),
         (400.0, 300.0, 30.0, 32, 8, 3), (33.3, 44.4, 17.7, 48, 12, 0), (80.0, 60.0, 96.0, 256, 64, 1), (-5.0, 130.0, 61.5, 64, 16, 2)]
out = {"image": image, "inst": inst, "cases": np.array(cases, dtype=np.float64), "cv2_version": np.array(cv2.__version__)}
for i, (cx, cy, s, S, R, iid) in enumerate(cases):
    c = np.array([cx, cy])
    S, R, iid = int(S), int(R), int(iid)
    roi_img = crop_resize_by_warp_affine(image, c, s, S, interpolation=cv2.INTER_NEAREST)
    roi_img = ((roi_img / 255.0 - MEAN) / STD).transpose(2, 0, 1)
    roi_coord = crop_resize_by_warp_affine(coord_2d, c, s, R, interpolation=cv2.INTER_NEAREST).transpose(2, 0, 1)
    mt = inst.astype(np.float32) if iid < 0 else (inst == iid).astype(np.float32)
    roi_mask = crop_resize_by_warp_affine(mt, c, s, S, interpolation=cv2.INTER_NEAREST)[None]
    out[f"c{i}/roi_img"] = roi_img.astype(np.float32)
    out[f"c{i}/roi_mask"] = roi_mask.astype(np.float32)
    out[f"c{i}/roi_coord_2d"] = roi_coord.astype(np.float32)
    out[f"c{i}/trans_img"] = get_affine_transform(c, (s, s), 0, (S, S))     # the tuple form crop_resize_by_warp_affine passes on
    out[f"c{i}/trans_out"] = get_affine_transform(c, (s, s), 0, (R, R))
# full_img (load_data_eval.py:336-338 with FLAGS.resize_full, the default): cv2.resize -> normalise -> CHW
frame = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
full = cv2.resize(frame, (64, 48))
out["full/frame"] = frame
out["full/resized_u8"] = full
out["full/full_img"] = ((full / 255.0 - MEAN) / STD).transpose(2, 0, 1).astype(np.float32)
np.savez_compressed(os.path.join(HERE, "roi.npz"), **out)
print("wrote", os.path.join(HERE, "roi.npz"), os.path.getsize(os.path.join(HERE, "roi.npz")), "bytes")
