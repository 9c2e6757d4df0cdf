"""TEST INFRASTRUCTURE — golden vectors for the bilinear resize, produced by PIL itself (the library the reference's
GroupScale / torchvision Resize call, data/transforms.py:79-92).

    python oracle/make_golden_resize.py        # writes tests/golden/resize_pil.npz (a few KB)

Inputs are regenerated from seeds (numpy PCG64), so the fixture stores, per case, the SHA-256 of PIL's 224 x 224 output,
its main diagonal and its first row: enough to pin an implementation bit for bit without shipping images."""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [(256, 256, 1), (360, 640, 2), (112, 160, 3), (225, 223, 4), (224, 224, 5), (480, 854, 6), (720, 1280, 7)]


def make_input(h, w, seed):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # half noise, half smooth ramps: exercises both the rounding of flat areas and the clamping of edges
    yy, xx = np.mgrid[0:h, 0:w]
    ramp = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
    base[:, : w // 2] = ramp[:, : w // 2]
    return base


def main():
    out = {"cases": np.array(CASES)}
    for h, w, seed in CASES:
        img = make_input(h, w, seed)
        pil = np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
        key = f"{h}x{w}"
        out[key + "_sha256"] = np.frombuffer(hashlib.sha256(pil.tobytes()).digest(), np.uint8)
        out[key + "_diag"] = pil[np.arange(224), np.arange(224)]
        out[key + "_row0"] = pil[0]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_pil.npz"), **out)
    print("written", {k: v.shape for k, v in out.items() if k != "cases"})


if __name__ == "__main__":
    sys.exit(main())
