"""Generates the committed golden fixtures from the oracle (run here, where /root/reference exists).

    python tests/golden/make_golden.py

Fixtures (all small, all seeded):
  contours_small.npz   bit-packed random/adversarial masks + cv2 findContours(RETR_EXTERNAL, SIMPLE) output
  postprocess_small.npz  class masks {0,1,2} + oracle postprocess_mask output (cv2 CCL / morphology)
  preprocess_hashes.json  sha256 of oracle preprocess_raw output for seeded u16 inputs of several sizes
  labelme_*.json / sidecar_*.json  text written by the reference's vendored nlohmann::json 3.12.0
                       (oracle/_ref/gen_json, compiled against /root/reference/include in place)
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pipeline as op  # noqa: E402

sys.path.insert(0, HERE)
from cases import contour_case_masks, postprocess_case_masks, preprocess_cases, json_cases  # noqa: E402


def main():
    # contours
    masks = contour_case_masks()
    rec = {}
    for i, m in enumerate(masks):
        cs = op.extract_contours(m)
        rec[f"shape_{i}"] = np.array(m.shape, np.int32)
        rec[f"bits_{i}"] = np.packbits(m > 127)
        rec[f"xy_{i}"] = (np.concatenate(cs) if cs else np.zeros((0, 2), np.int32)).astype(np.int16)
        rec[f"len_{i}"] = np.array([len(c) for c in cs], np.int32)
    rec["n"] = np.array(len(masks))
    np.savez_compressed(os.path.join(HERE, "contours_small.npz"), **rec)
    print("contours:", len(masks), "masks")

    # postprocess
    rec = {}
    pm = postprocess_case_masks()
    for i, m in enumerate(pm):
        out = op.postprocess_mask(m)
        rec[f"shape_{i}"] = np.array(m.shape, np.int32)
        rec[f"in_{i}"] = m
        rec[f"out_{i}"] = np.packbits(out == 2)
    rec["n"] = np.array(len(pm))
    np.savez_compressed(os.path.join(HERE, "postprocess_small.npz"), **rec)
    print("postprocess:", len(pm), "masks")

    # preprocess
    hashes = {}
    for name, src in preprocess_cases().items():
        hashes[name] = hashlib.sha256(op.preprocess_raw(src).tobytes()).hexdigest()
    with open(os.path.join(HERE, "preprocess_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("preprocess:", len(hashes), "cases")

    # JSON via the reference's vendored nlohmann (needs /root/reference)
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle", "ref_json")], check=True, capture_output=True)
    gen = os.path.join(ROOT, "oracle", "_ref", "gen_json")
    for name, (base, w, h, contours) in json_cases().items():
        inp = f"{base} {w} {h} {len(contours)}\n" + "\n".join(
            f"{len(c)} " + " ".join(f"{x} {y}" for x, y in c) for c in contours) + "\n"
        txt = subprocess.run([gen, "labelme"], input=inp.encode(), capture_output=True, check=True).stdout
        with open(os.path.join(HERE, f"labelme_{name}.json"), "wb") as f:
            f.write(txt)
    txt = subprocess.run([gen, "sidecar"], input=b"slice_007.raw 600 400 512 512\n", capture_output=True, check=True).stdout
    with open(os.path.join(HERE, "sidecar_slice_007.json"), "wb") as f:
        f.write(txt)
    print("json: done")


if __name__ == "__main__":
    main()
