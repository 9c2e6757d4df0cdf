"""GPU: the fused resize + pre-processing kernel (csrc/resize.cu) against the oracle — BIT-EXACT on the uint8 result (it
is integer arithmetic) and on the normalised stem input, and through the engine's uint8 entry points."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w", [(256, 256), (360, 640), (112, 160), (225, 223), (224, 224), (480, 854), (1080, 1920)])
def test_resize_u8_bit_exact(h, w):
    from oracle import resize_oracle as ro
    from oracle.make_golden_resize import make_input
    from vcg_b200 import ops
    frames = np.stack([make_input(h, w, 10 + i) for i in range(3)])
    want = ro.resize_frames(frames)
    got, stem = ops.resize_u8(torch.from_numpy(frames).cuda(), want_u8=True, stem_dtype=torch.bfloat16)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), want), (h, w)
    # the fused normalise + cast equals the plain pre-processing kernel on the resized frames, bit for bit
    ref_stem = ops.preprocess_u8(torch.from_numpy(want).cuda(), dtype=torch.bfloat16)
    assert torch.equal(stem.view(torch.int16), ref_stem.view(torch.int16))
    _, stem32 = ops.resize_u8(torch.from_numpy(frames).cuda(), want_u8=False, stem_dtype=torch.float32)
    assert torch.equal(stem32, ops.preprocess_u8(torch.from_numpy(want).cuda(), dtype=torch.float32))


def test_resize_matches_pil_golden(golden_dir):
    import hashlib
    from oracle.make_golden_resize import make_input
    from vcg_b200 import ops
    g = np.load(f"{golden_dir}/resize_pil.npz")
    for h, w, seed in g["cases"].tolist():
        got, _ = ops.resize_u8(torch.from_numpy(make_input(h, w, seed)[None]).cuda())
        assert hashlib.sha256(got[0].cpu().numpy().tobytes()).digest() == g[f"{h}x{w}_sha256"].tobytes(), (h, w)


def test_engine_scores_resized_frames_like_pre_resized_ones():
    """256 x 256 source frames through score_video_u8 / score_clips_u8 / the host-buffer call == the same calls on frames
    resized by the oracle: identical logits (the resized uint8 frames are identical, everything after is the same code)."""
    from oracle import resize_oracle as ro
    from oracle import weights as W
    from vcg_b200.engine import Engine
    T, L, n_frames = 8, 24, 28
    sd = W.make_state_dict(T, "mlp", seed=123)
    rng = np.random.default_rng(3)
    big = rng.integers(0, 256, (n_frames, 256, 320, 3), dtype=np.uint8)
    small = torch.from_numpy(ro.resize_frames(big))
    big = torch.from_numpy(big)
    starts = W.clip_starts(n_frames, T)
    B = len(starts)
    ids, mask = W.make_text(B, L, seed=4)
    st = torch.tensor(starts, dtype=torch.int32)
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=L, max_batch=4)
    eng.load_state_dict(sd)
    want, _ = eng.score_video_u8(small.cuda(), 0, 4, ids.cuda(), mask.cuda())
    want_g, _ = eng.score_clips_u8(small.cuda(), st.cuda(), ids.cuda(), mask.cuda())
    got, _ = eng.score_video_u8(big.cuda(), 0, 4, ids.cuda(), mask.cuda())
    got_g, _ = eng.score_clips_u8(big.cuda(), st.cuda(), ids.cuda(), mask.cuda())
    got_h, _ = eng.score_clips_u8_host(big.pin_memory(), st, ids, mask)
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(got_g, want_g) and torch.equal(got_h.cuda(), want)
    # and back to 224 x 224 frames on the same engine
    again, _ = eng.score_video_u8(small.cuda(), 0, 4, ids.cuda(), mask.cuda())
    assert torch.equal(again, want)
    eng.close()
