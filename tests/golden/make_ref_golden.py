"""Generates tests/golden/ref_pin.npz + ref_*.json|txt from the REFERENCE'S OWN OBJECT CODE (oracle/_ref/libref_pipeline.so:
/root/reference/src/{preprocess,postprocess,mask2polygon}.cpp compiled unmodified against the OpenCV stub; run here, where
/root/reference exists).

    make -C oracle/ref_build && python tests/golden/make_ref_golden.py

These fixtures pin P1c (preprocess_raw), P5 (postprocess_mask), P7a-c (extract_contours, map_contour_points, generate_json)
and P7e (process_single_mask's file protocol) to what the reference's compiled statements produce; tests/test_ref_pin.py
checks oracle/pipeline.py against them on CPU and tests/test_gpu_ref_pin.py checks the CUDA path against them."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref  # noqa: E402
from cases import contour_case_masks, postprocess_case_masks, preprocess_cases, json_cases, ref_mask_cases, MAP_SCALES  # noqa: E402


def main():
    assert ref.available(), "build oracle/_ref/libref_pipeline.so first (make -C oracle/ref_build)"
    rec, text = {}, {}

    # P1c: Preprocess::preprocess_raw -- pixels (sha256 + the first case in full) and the sidecar text
    for name, src in preprocess_cases().items():
        px, side = ref.preprocess_raw(src, "slice_007.raw")
        text[f"pre_sha_{name}"] = hashlib.sha256(px.tobytes()).hexdigest()
        text[f"pre_sidecar_{name}"] = side
    rec["pre_full_rand_333x517"] = ref.preprocess_raw(preprocess_cases()["rand_333x517"])[0]

    # P5: ::postprocess_mask on the small cases and on 512 x 512 class masks
    pm = postprocess_case_masks() + ref_mask_cases()
    for i, m in enumerate(pm):
        rec[f"post_{i}"] = np.packbits(ref.postprocess_mask(m) == 2)
    rec["post_n"] = np.array(len(pm))

    # P7a / P7b: extract_contours, then map_contour_points at several scales, on every contour case
    masks = contour_case_masks()
    for i, m in enumerate(masks):
        cs = ref.extract_contours(m)
        rec[f"cnt_len_{i}"] = np.array([len(c) for c in cs], np.int32)
        rec[f"cnt_xy_{i}"] = (np.concatenate(cs) if cs else np.zeros((0, 2), np.int32)).astype(np.int32)
        for k, (sx, sy) in enumerate(MAP_SCALES):
            mp = ref.map_contour_points(cs, sx, sy)
            rec[f"map_xy_{i}_{k}"] = (np.concatenate(mp) if mp else np.zeros((0, 2), np.int32)).astype(np.int32)
    rec["cnt_n"] = np.array(len(masks))

    # P7c: generate_json
    for name, (base, w, h, contours) in json_cases().items():
        text[f"labelme_{name}"] = ref.generate_json([np.array(c, np.int32) for c in contours], base, w, h)

    # P7e: process_single_mask through files (stub PNG codec), for a post-processed 512 x 512 mask of a 600 x 400 slice
    with tempfile.TemporaryDirectory() as td:
        m = ref.postprocess_mask(ref_mask_cases()[0])
        vis = np.where(m == 2, 255, 0).astype(np.uint8)
        norm = (np.add.outer(np.arange(512), np.arange(512)) % 251).astype(np.uint8)
        ref.write_png_gray(os.path.join(td, "s_mask.png"), vis)
        ref.write_png_gray(os.path.join(td, "s_normalized.png"), norm)
        with open(os.path.join(td, "s_original_sizes.json"), "w") as f:
            f.write('{"s.raw":{"original_height":400,"original_width":600,"scaled_height":512,"scaled_width":512}}\n')
        ref.process_single_mask(os.path.join(td, "s_mask.png"), td, os.path.join(td, "s_original_sizes.json"),
                                os.path.join(td, "s_normalized.png"), "s")
        text["psm_json"] = open(os.path.join(td, "s.json")).read()
        ov = ref.read_png_bgr(os.path.join(td, "s_contour_overlay.png"))
        red = (ov[:, :, 2] == 255) & (ov[:, :, 1] == 0) & (ov[:, :, 0] == 0)
        rec["psm_overlay_red"] = np.packbits(red)
        text["psm_overlay_sha"] = hashlib.sha256(ov.tobytes()).hexdigest()

    np.savez_compressed(os.path.join(HERE, "ref_pin.npz"), **rec)
    with open(os.path.join(HERE, "ref_pin_text.json"), "w") as f:
        json.dump(text, f, indent=1, sort_keys=True)
    print("ref_pin: %d arrays, %d texts" % (len(rec), len(text)))


if __name__ == "__main__":
    main()
