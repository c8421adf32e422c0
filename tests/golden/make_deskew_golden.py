"""Mint tests/golden/deskew_golden.json from the REFERENCE's own `deskew` (src/preprocessing/normalise.py:19-57, imported
unchanged from /root/reference; run in the build container only):  python tests/golden/make_deskew_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference")

import ref_preproc as P  # noqa: E402
from src.preprocessing.normalise import deskew  # noqa: E402

out = {"cv2": __import__("cv2").__version__, "cases": []}
for (h, w, seed, ang) in P.DESKEW_CASES:
    img = P.tooth_image(h, w, seed, ang)
    rot, a = deskew(img)
    out["cases"].append({"h": h, "w": w, "seed": seed, "tilt": ang, "input": hashlib.sha1(img.tobytes()).hexdigest(),
                         "angle": float(a), "output": hashlib.sha1(np.ascontiguousarray(rot).tobytes()).hexdigest()})
    print(out["cases"][-1])
json.dump(out, open(os.path.join(HERE, "deskew_golden.json"), "w"), indent=1)
