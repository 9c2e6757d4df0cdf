"""CPU: the resize oracle (oracle/resize_oracle.py, Pillow's ImagingResample restated) against PIL itself and against the
committed PIL goldens (tests/golden/resize_pil.npz, oracle/make_golden_resize.py)."""
import hashlib

import numpy as np
import pytest


def test_resize_oracle_matches_pil_goldens(golden_dir):
    from oracle import resize_oracle as ro
    from oracle.make_golden_resize import make_input
    g = np.load(f"{golden_dir}/resize_pil.npz")
    for h, w, seed in g["cases"].tolist():
        img = make_input(h, w, seed)
        got = ro.resize_bilinear_u8(img, 224, 224)
        key = f"{h}x{w}"
        assert hashlib.sha256(got.tobytes()).digest() == g[key + "_sha256"].tobytes(), key
        assert np.array_equal(got[np.arange(224), np.arange(224)], g[key + "_diag"])
        assert np.array_equal(got[0], g[key + "_row0"])


def test_resize_oracle_matches_pil_live():
    """The same against the PIL of this image, on sizes the fixture does not hold (up- and down-scaling, odd sizes)."""
    Image = pytest.importorskip("PIL.Image")
    from oracle import resize_oracle as ro
    rng = np.random.default_rng(0)
    for h, w in [(100, 300), (299, 224), (224, 100), (448, 448), (31, 47)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
        assert np.array_equal(ro.resize_bilinear_u8(img, 224, 224), want), (h, w)


def test_coefficient_tables():
    """Fixed-point weights: non-negative, sum within rounding of 2^22, identity when nothing changes."""
    from oracle import resize_oracle as ro
    for n_in in (224, 256, 360, 640, 112):
        b, k = ro.precompute_coeffs(n_in, 224)
        assert (k >= 0).all() and (np.abs(k.sum(1) - (1 << 22)) <= k.shape[1]).all()
        assert (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= n_in).all()
    b, k = ro.precompute_coeffs(224, 224)
    assert (b[:, 0] == np.arange(224)).all() and (k[:, 0] == 1 << 22).all() and (k[:, 1:] == 0).all()
