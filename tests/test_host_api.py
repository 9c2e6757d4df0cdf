"""Host-side mirror of the reference API (model/, ops/, eval_utils/, common_utils/): import paths, constructor
signatures, checkpoint key schema and error behaviour — everything that does not need the GPU."""
import numpy as np
import pytest
import torch


def build(T=16, head="mlp"):
    from model.fusion import two_stream
    from model.lang import bert_hugface
    from model.vision import resnet50_tsm
    lang = bert_hugface.BertHugface(pretrain_stage=False)
    vis = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
    m = two_stream.TwoStream(lang.base_model, vis.base_model, lang.embed_size, vis.feature_dim, T, 128)
    m.build_chapter_head(output_size=2, head_type=head)
    return m, lang, vis


@pytest.fixture(scope="module")
def built():
    return build()


def test_constructor_attributes(built):
    m, lang, vis = built
    assert lang.embed_size == 768 and lang.vocab_size == 30522 and vis.feature_dim == 2048
    assert hasattr(m, "vision_model") and hasattr(m, "lang_model")
    assert m.eval() is m
    m.vision_model = m.vision_model.eval()
    m.lang_model = m.lang_model.eval()
    # reference caller #1 nulls running stats of every nn.BatchNorm2d (test_video_segment_point.py:116-122):
    # there must be none, so folded eval-mode statistics survive that loop
    assert not any(isinstance(x, torch.nn.BatchNorm2d) for x in m.modules())
    assert sum(p.numel() for p in m.parameters()) == 133355074


def test_state_dict_schema_matches_reference(built):
    from oracle import weights as W
    m, _, _ = built
    sd = W.make_state_dict(16, "mlp")      # make_golden.py loads this dict into the real reference with strict=True
    ours = m.state_dict()
    assert set(ours) == set(sd)
    assert all(tuple(ours[k].shape) == tuple(sd[k].shape) for k in sd)
    assert "vision_model.layer1.0.conv1.net.weight" in ours          # TemporalShift wrapper key
    assert "fusion_head.head.weight" in ours and ours["fusion_head.head.weight"].shape == (2, 17 * 128)
    m.load_state_dict(sd, strict=True)


def test_attn_head_schema_and_unknown_head():
    from model.fusion import two_stream
    from ops.basic_ops import Identity
    m = two_stream.TwoStream(Identity(), Identity(), 768, 2048, 8, 128)
    m.build_chapter_head(output_size=2, head_type="attn")
    keys = set(m.state_dict())
    assert {"fusion_head.head.key.weight", "fusion_head.head.query.bias", "fusion_head.head.value.weight",
            "fusion_head.head.proj.weight"} <= keys
    with pytest.raises(RuntimeError, match="Unknown head_type"):
        m.build_chapter_head(output_size=2, head_type="bilinear")


def test_forward_without_gpu_raises(built):
    m, _, _ = built
    m.eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 16, 3, 224, 224), torch.zeros(1, 8, dtype=torch.long), torch.ones(1, 8, dtype=torch.long))


def test_eval_utils_mirror(golden_dir):
    from eval_utils.eval_utils import calculate_pr, convert_clip_label2cut_point
    g = np.load(f"{golden_dir}/cut_points.npz", allow_pickle=True)
    for labels, T, cuts in zip(g["labels"], g["T"], g["cuts"]):
        assert convert_clip_label2cut_point(list(labels), int(T), 2) == list(cuts)
    assert calculate_pr([10, 50, 100], [10, 52, 96, 200]) == tuple(g["pr"])
    assert calculate_pr([5], []) == (0.0, 0.0, 0.0, None, None, None)
    with pytest.raises(ZeroDivisionError):
        calculate_pr([], [3])


def test_temporal_shift_mirror():
    from ops.temporal_shift import TemporalShift
    from oracle import two_stream_oracle as orc
    x = torch.randn(12, 32, 3, 3)
    assert torch.equal(TemporalShift.shift(x, 4, fold_div=8), orc.temporal_shift(x, 4, 8))


def test_seed_helper():
    from common_utils import set_random_seed
    set_random_seed.use_fix_random_seed()
    a = torch.rand(3)
    set_random_seed.use_fix_random_seed()
    assert torch.equal(a, torch.rand(3))


def test_flat_clip_reader(tmp_path):
    """Flat-clip JSON (flat_video2clip_for_quick_infer.py:112-119) -> frame table with every JPEG decoded once, per-clip
    start indices and the reference's tokenisation ("[CLS] " + text, truncate, [PAD], mask 1/0)."""
    import json
    import os
    import sys
    import numpy as np
    import torch
    from PIL import Image
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video-chapter-generation_b200"))
    from vcg_b200 import flat_clips as fc

    T, L, n_frames = 8, 12, 20
    rng = np.random.RandomState(0)
    vid_dir = tmp_path / "vidA"
    vid_dir.mkdir()
    pixels = {}
    for i in range(n_frames):
        arr = rng.randint(0, 256, size=(224, 224, 3), dtype=np.uint8)
        p = str(vid_dir / f"{i + 1:05d}.png")            # PNG: lossless, so pixel equality can be asserted
        Image.fromarray(arr).save(p)
        pixels[p] = arr
    paths = sorted(pixels)
    clips = []
    for s in range(0, n_frames - T, 4):
        clips.append({"image_paths": paths[s:s + T], "text_clip": "hello world " * (s + 1), "clip_label": int(s == 4),
                      "clip_start_end": [s, s + T], "cut_points": [6], "vid": "vidA"})
    jf = tmp_path / "clips.json"
    jf.write_text(json.dumps(clips))

    class Tok:   # minimal stand-in for BertTokenizer
        def tokenize(self, text):
            return text.split()

        def convert_tokens_to_ids(self, toks):
            return [{"[CLS]": 101, "[PAD]": 0, "hello": 7592, "world": 2088}[t] for t in toks]

    videos = list(fc.iter_videos(str(jf), Tok(), T, L, pin=False))
    assert len(videos) == 1
    v = videos[0]
    assert v.vid == "vidA" and len(v) == len(clips) == 3
    assert v.frames.shape == (16, 224, 224, 3) and v.frames.dtype == torch.uint8       # 3 clips x 8 frames, 16 distinct
    assert v.clip_start.tolist() == [0, 4, 8]
    for c, info in enumerate(clips):
        for t, p in enumerate(info["image_paths"]):
            assert np.array_equal(v.frames[v.clip_start[c] + t].numpy(), pixels[p])
    assert v.text_ids[0].tolist() == [101, 7592, 2088] + [0] * 9 and v.attention_mask[0].tolist() == [1] * 3 + [0] * 9
    assert v.attention_mask[2].sum() == L and v.text_ids[2, 0] == 101                    # truncated at max_text_len
    assert v.labels.tolist() == [0, 1, 0] and v.cut_points == [6]


def test_evaluation_host_logic():
    """Host-side pieces of vcg_b200.evaluate / postprocess that need no GPU: the reference loop's clip grouping (first
    clip of every video twice) and the averaging of test_video_segment_point.py:345-358."""
    import math
    from oracle import metrics_oracle as mo
    from vcg_b200 import evaluate as ev
    from vcg_b200 import postprocess as pp
    vids = ["a", "a", "a", "b", "c", "c"]
    idx, off = pp.reference_video_groups(vids)
    assert idx.tolist() == [0, 0, 1, 2, 3, 3, 4, 4, 5] and off.tolist() == [0, 4, 6, 9]
    assert [idx[off[v]:off[v + 1]].tolist() for v in range(3)] == mo.reference_video_groups(vids)
    assert pp.reference_video_groups([])[1].tolist() == [0]
    out = {}
    ev._summarise(out, [(1.0, 1.0, 1.0, 0.5, 0.5, 1.0), (0.0, 0.5, 0.5, None, None, None)], "")
    assert out["recall"] == 0.5 and out["recall@3"] == 0.75 and out["precision"] == 0.5 and out["precision@5"] == 1.0
    assert out["f-score"] == 0.5 and abs(out["f-score@5"] - 2 * 0.75 * 1.0 / 1.75) < 1e-15
    ev._summarise(out, [(0.0, 0.0, 0.0, None, None, None)], "_rand")
    assert out["recall_rand"] == 0.0 and math.isnan(out["precision_rand"]) and math.isnan(out["f-score_rand"])
