"""The oracle (test infrastructure) against the committed golden vectors, which oracle/make_golden.py produced by
running the UNMODIFIED reference in the build container.  Runs on CPU."""
import numpy as np
import pytest
import torch

from oracle import two_stream_oracle as orc
from oracle import weights as W


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.fixture(scope="module")
def golden_attn(golden_dir):
    return np.load(f"{golden_dir}/two_stream_attn_T8_L32_B2.npz")


def test_oracle_reproduces_reference_outputs(golden_attn):
    g = golden_attn
    T, L, B, seed = [int(x) for x in g["meta"][:4]]
    sd = W.make_state_dict(T, "attn", seed=seed)
    ids, mask = W.make_text(B, L, seed=seed)
    assert np.array_equal(ids.numpy(), g["text_ids"])
    frames = W.make_frames_u8(4 * (B - 1) + T, seed=seed)
    img = orc.gather_clips(orc.preprocess_u8(frames), [int(s) for s in g["clip_starts"]], T)
    taps = {}
    with torch.no_grad():
        logits, probs, vis, lang = orc.two_stream_forward(sd, img, ids, mask, T, 128, "attn", 8, taps=taps)
    assert rel(logits, torch.from_numpy(g["logits"])) <= 1e-5
    assert rel(probs, torch.from_numpy(g["probs"])) <= 1e-5
    assert rel(vis, torch.from_numpy(g["vision_emb"])) <= 1e-5
    assert rel(lang, torch.from_numpy(g["lang_emb"])) <= 1e-5
    assert orc.predict_labels(logits) == g["labels"].tolist()
    for k, v in taps.items():
        ref = g["tap/" + k]
        t = v.float().reshape(-1)
        got = np.array([t.mean().item(), t.abs().mean().item(), t.abs().max().item(),
                        t.double().pow(2).sum().sqrt().item()])
        assert np.allclose(got, ref, rtol=1e-4, atol=1e-6), (k, got, ref)


def test_cut_points_known_answers(golden_dir):
    g = np.load(f"{golden_dir}/cut_points.npz", allow_pickle=True)
    for labels, T, cuts in zip(g["labels"], g["T"], g["cuts"]):
        assert orc.convert_clip_label2cut_point(list(labels), int(T), 2) == list(cuts)
    vec = [1, 0, 0, 0, 1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
    assert orc.convert_clip_label2cut_point(vec, 8, 2) == [4, 22, 44, 60]
    assert orc.convert_clip_label2cut_point(vec, 16, 2) == [8, 26, 48, 64]
    assert orc.convert_clip_label2cut_point(vec, 32, 2) == [16, 34, 56, 72]
    assert orc.convert_clip_label2cut_point([0, 1, 1], 16, 2) == []
    assert orc.calculate_pr([10, 50, 100], [10, 52, 96, 200]) == tuple(g["pr"])
    with pytest.raises(ZeroDivisionError):
        orc.calculate_pr([], [1])


def test_temporal_shift_restatement():
    x = torch.arange(2 * 4 * 16 * 1 * 1, dtype=torch.float32).view(8, 16, 1, 1)
    y = orc.temporal_shift(x, 4, 8).view(2, 4, 16)
    xv = x.view(2, 4, 16)
    assert torch.equal(y[:, :-1, :2], xv[:, 1:, :2]) and torch.all(y[:, -1, :2] == 0)
    assert torch.equal(y[:, 1:, 2:4], xv[:, :-1, 2:4]) and torch.all(y[:, 0, 2:4] == 0)
    assert torch.equal(y[:, :, 4:], xv[:, :, 4:])


def test_preprocess_matches_totensor_normalize():
    frames = W.make_frames_u8(2, seed=5)
    x = orc.preprocess_u8(frames)
    assert x.shape == (2, 3, 224, 224)
    f = frames[1, 10, 20].float() / 255.0
    exp = (f - torch.tensor(orc.IMAGENET_MEAN)) / torch.tensor(orc.IMAGENET_STD)
    assert torch.allclose(x[1, :, 10, 20], exp, atol=1e-6)


@pytest.mark.parametrize("case,kind", [("r50tsm_T8_B2", "r50tsm"), ("r50_T8_B2", "r50"), ("bert_L48_B3", "bert")])
def test_unimodal_oracle_matches_reference_golden(golden_dir, case, kind):
    """Single-modality scorers (--data_mode image / text): restatement vs the reference's own outputs."""
    import numpy as np
    import torch
    from oracle import two_stream_oracle as orc
    from oracle import weights as W
    g = np.load(f"{golden_dir}/unimodal_{case}.npz")
    T, L, B, seed = [int(x) for x in g["meta"]]
    sd = W.make_unimodal_state_dict(kind, clip_frames=max(T, 1), seed=123)
    with torch.no_grad():
        if kind == "bert":
            ids, mask = W.make_text(B, L, seed=seed)
            assert np.array_equal(ids.numpy(), g["text_ids"])
            logits, probs, _ = orc.text_only_forward(sd, ids, mask)
        else:
            frames = W.make_frames_u8(4 * (B - 1) + T, seed=seed)
            img = orc.gather_clips(orc.preprocess_u8(frames), [int(s) for s in g["clip_starts"]], T)
            logits, probs, _ = orc.vision_only_forward(sd, img, T, 8 if kind == "r50tsm" else 0)
    assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-5 * np.abs(g["logits"]).max()
    assert np.abs(probs.numpy() - g["probs"]).max() <= 1e-5
    assert orc.predict_labels(logits) == g["labels"].tolist()


def test_vision_emb_io_roundtrip(tmp_path):
    """convert2vision_emb.py's file layout: vision_emb_{start}_{end}.npy, fp32 [T,2048] per clip."""
    import os
    import sys
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video-chapter-generation_b200"))
    from vcg_b200 import vision_emb_io as vio
    emb = torch.randn(3, 16, 2048)
    infos = [{"vid": "abc", "clip_start_end": (4 * i, 4 * i + 16)} for i in range(3)]
    paths = vio.save_vision_embs(str(tmp_path), infos, emb)
    assert [os.path.basename(p) for p in paths] == ["vision_emb_0_16.npy", "vision_emb_4_20.npy", "vision_emb_8_24.npy"]
    assert np.load(paths[1]).shape == (16, 2048) and np.load(paths[1]).dtype == np.float32
    back = vio.load_vision_embs(str(tmp_path), "abc", [0, 4, 8], 16)
    assert back.shape == (3, 16, 2048, 1, 1) and torch.equal(back.view(3, 16, 2048), emb)


@pytest.mark.parametrize("case,head", [("cross_attn_T8_w1_L24_B2", "cross_attn"), ("mlp_T8_w1_L24_B2", "mlp"),
                                       ("bilinear_T8_w1_L24_B2", "bilinear"),
                                       ("multiplication_T8_w1_L24_B2", "multiplication"),
                                       ("self_attn_T8_w1_L24_B2", "self_attn")])
def test_window_oracle_matches_reference_golden(golden_dir, case, head):
    """Window ("update") model: restatement vs the reference's own outputs (oracle/make_golden_window.py)."""
    import numpy as np
    import torch
    from oracle import weights as W
    from oracle import window_oracle as worc
    from oracle.make_golden_window import make_inputs
    g = np.load(f"{golden_dir}/window_{case}.npz")
    T, window, L, B, seed = [int(x) for x in g["meta"]]
    sd = W.make_window_state_dict(T, window, head, seed=123)
    img, ids, mask = make_inputs(T, window, L, B, seed)
    assert np.array_equal(ids.numpy(), g["text_ids"])
    with torch.no_grad():
        logits, probs = worc.window_forward(sd, img, ids, mask, T, head)
    assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-5 * np.abs(g["logits"]).max()
    assert np.abs(probs.numpy() - g["probs"]).max() <= 1e-5


def test_metrics_oracle_matches_sklearn():
    """roc_auc / average_precision restatement vs the scikit-learn of this image (the reference's own dependency), with
    ties and the single-class cases; and the reference loop's clip grouping (first clip of each video twice)."""
    import numpy as np
    import warnings
    sk = pytest.importorskip("sklearn.metrics")
    from oracle import metrics_oracle as mo
    rng = np.random.RandomState(0)
    cases = [([0, 0, 0], [.1, .2, .3]), ([1, 1], [.3, .4]), ([0, 1, 1, 0], [.5, .5, .5, .5]), ([1], [0.2])]
    for n in (2, 7, 50, 400):
        for _ in range(6):
            y = (rng.rand(n) < 0.2).astype(int)
            s = rng.rand(n).astype(np.float32)
            s = np.round(s, 1) if rng.rand() < 0.5 else s       # heavy ties half of the time
            cases.append((y.tolist(), s.tolist()))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for y, s in cases:
            fpr, tpr, _ = sk.roc_curve(y, s, pos_label=1)
            want_auc, want_ap = sk.auc(fpr, tpr), sk.average_precision_score(y, s)
            got_auc, got_ap = mo.roc_auc(y, s), mo.average_precision(y, s)
            assert (np.isnan(want_auc) and np.isnan(got_auc)) or abs(got_auc - want_auc) <= 1e-12, (y, s)
            assert abs(got_ap - want_ap) <= 1e-12, (y, s)
    assert mo.reference_video_groups(["a", "a", "b", "b", "b", "c"]) == [[0, 0, 1], [2, 2, 3, 4], [5, 5]]


def test_variant_oracles_match_reference_golden(golden_dir):
    """The reference's two unused fusion variants (two_stream_domain_specific.py, window_self_attention.py):
    restatements vs the reference's own outputs (oracle/make_golden_window.py)."""
    import numpy as np
    import torch
    from oracle import weights as W
    from oracle import window_oracle as worc
    from oracle.make_golden_window import make_inputs
    g = np.load(f"{golden_dir}/window_domain_T8_w1_L24_B2.npz")
    T, window, L, B, seed = [int(x) for x in g["meta"]]
    sd = W.make_domain_state_dict(T, window, seed=123)
    img, ids, mask = make_inputs(T, window, L, B, seed)
    with torch.no_grad():
        logits, probs = worc.domain_specific_forward(sd, img, ids, mask, T)
    assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-5 * np.abs(g["logits"]).max()
    assert np.abs(probs.numpy() - g["probs"]).max() <= 1e-5
    for window in (1, 2):
        g = np.load(f"{golden_dir}/window_single_block_w{window}_B5.npz")
        sd = W.make_single_block_state_dict(window, seed=123)
        with torch.no_grad():
            logits, probs = worc.single_block_classifier(sd, torch.from_numpy(g["x"]))
        assert np.abs(logits.numpy() - g["logits"]).max() <= 1e-5 * np.abs(g["logits"]).max()
        assert np.abs(probs.numpy() - g["probs"]).max() <= 1e-5


def test_oracle_reproduces_video_scale_goldens(golden_dir):
    """Video-scale fixtures (oracle/make_golden_video.py, UNMODIFIED reference): the oracle restatement on a sample of the
    146 clips of the 10-minute video (first clips, a scene change, the tail) and on all 256 clips of configs[1]; the
    committed labels / timestamps are what the reference's topk + convert_clip_label2cut_point give."""
    g = np.load(f"{golden_dir}/video_mlp_T16_L100_600f.npz")
    T, L, B, seed, _, _, n_frames, _ = [int(x) for x in g["meta"]]
    sd = W.make_state_dict(T, "mlp", seed=seed)
    sd["fusion_head.head.bias"] = torch.from_numpy(g["head_bias"]).clone()
    frames, scenes = W.make_video_u8(n_frames, seed=seed)
    starts = W.clip_starts(n_frames, T)
    assert scenes == g["scene_starts"].tolist() and starts == g["clip_starts"].tolist()
    ids, mask = W.make_video_text(starts, scenes, T, L, seed=seed)
    ref = torch.from_numpy(g["logits"])
    sample = [0, 16, 17, B - 1]                      # clip 16/17: the first label change of the video
    with torch.no_grad():
        img = orc.gather_clips(orc.preprocess_u8(frames), [starts[i] for i in sample], T)
        logits = orc.two_stream_forward(sd, img, ids[sample], mask[sample], T, 128, "mlp", 8)[0]
    assert float((logits - ref[sample]).abs().max() / ref.abs().max()) <= 1e-4
    assert orc.predict_labels(logits) == g["labels"][sample].tolist()
    labels = orc.predict_labels(ref)
    assert labels == g["labels"].tolist() and 0 < sum(labels) < B
    assert orc.convert_clip_label2cut_point(labels, T, 2) == g["cut_points"].tolist()
    assert len(g["cut_points"]) >= 3
    m = ref[:, 1] - ref[:, 0]
    assert abs(float(m.abs().min()) - float(g["min_abs_margin"])) < 1e-9

    # the same video through the attention head (fusion_head.head.proj.bias re-centred)
    ga = np.load(f"{golden_dir}/video_attn_T16_L100_600f.npz")
    sda = W.make_state_dict(T, "attn", seed=seed)
    sda["fusion_head.head.proj.bias"] = torch.from_numpy(ga["head_bias"]).clone()
    refa = torch.from_numpy(ga["logits"])
    with torch.no_grad():
        la = orc.two_stream_forward(sda, img[:2], ids[sample[:2]], mask[sample[:2]], T, 128, "attn", 8)[0]
    assert float((la - refa[sample[:2]]).abs().max() / max(float(ga["raw_logit_absmax"]), float(refa.abs().max()))) <= 1e-4
    labels_a = orc.predict_labels(refa)
    assert labels_a == ga["labels"].tolist() and 0 < sum(labels_a) < B
    assert orc.convert_clip_label2cut_point(labels_a, T, 2) == ga["cut_points"].tolist() and len(ga["cut_points"]) >= 3

    g2 = np.load(f"{golden_dir}/video_emb_mlp_T16_L100_B256.npz")
    T, L, B, seed = [int(x) for x in g2["meta"][:4]]
    sd2 = W.make_state_dict(T, "mlp", seed=seed, include_vision=False)
    sd2["fusion_head.head.bias"] = torch.from_numpy(g2["head_bias"]).clone()
    emb, ids2, mask2 = W.make_precomputed_inputs(B, T, L, seed=seed)
    with torch.no_grad():
        l2 = orc.two_stream_forward(sd2, None, ids2, mask2, T, 128, "mlp", 8, vision_emb=emb)[0]
    ref2 = torch.from_numpy(g2["logits"])
    assert float((l2 - ref2).abs().max() / ref2.abs().max()) <= 1e-4
    assert orc.predict_labels(l2) == g2["labels"].tolist()
    assert orc.convert_clip_label2cut_point(g2["labels"].tolist(), T, 2) == g2["cut_points"].tolist()
