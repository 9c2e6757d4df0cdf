"""VIDEO-SCALE parity against the reference's own forward (the north_star's "chapter timestamps identical to the
reference"): tests/golden/video_*.npz were produced by oracle/make_golden_video.py from the UNMODIFIED reference
(model/fusion/two_stream.py TwoStream.forward + eval_utils.convert_clip_label2cut_point).

  configs[0]  one synthetic 10-minute video, 600 uint8 frames -> 146 clips (data/infer_youtube_video_dataset.py:117),
              T=16, L=100: logits [146,2] within tolerance, labels identical, chapter timestamps IDENTICAL
              (5 timestamps, both labels present in runs), in fp32 AND bf16, through every entry point:
                A. the callers' loop: TwoStream mirror, fp32 normalised img_clip, batch 16
                   (test_whole_pipeline_per_video.py:121,145-166)
                B. device-resident uint8 frames (vcg_score_video_u8: shared stem)
                C. host buffers (vcg_score_clips_u8_host: H2D/D2H inside)
                D. two ranks, clip-sharded, one all-gather of the logits (vcg_b200.distributed.score_sharded)
  configs[1]  precomputed vision embeddings, one batch of 256 clips.

Tolerances (BASELINE.json): logits max|delta| / max|ref| <= 1e-4 (fp32 mode), 2e-2 (bf16 mode), max|ref| being the
magnitude of the reference's logits with its OWN head bias (``raw_logit_absmax`` in the fixture): the fixtures then add
one constant per logit to the head bias so that the decisions straddle zero, which leaves every delta unchanged but
would shrink max|ref| 4x; the error against the re-centred logits is printed as well.
Margin safety: a label can only flip if the error of l1 - l0 reaches the clip's |margin|; the tests assert the
largest margin error stays below HALF the smallest reference |margin| (histogram printed), i.e. no clip sits inside
the error band of either precision.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = {"fp32": 1e-4, "bf16": 2e-2}


BIAS_KEY = {"mlp": "fusion_head.head.bias", "attn": "fusion_head.head.proj.bias"}


def _video_case(golden_dir, head="mlp"):
    from oracle import weights as W
    g = np.load(f"{golden_dir}/video_{head}_T16_L100_600f.npz")
    T, L, B, seed, _, _, n_frames, _ = [int(x) for x in g["meta"]]
    sd = W.make_state_dict(T, head, seed=seed)
    sd[BIAS_KEY[head]] = torch.from_numpy(g["head_bias"]).clone()
    frames, scenes = W.make_video_u8(n_frames, seed=seed)
    assert scenes == g["scene_starts"].tolist()
    starts = W.clip_starts(n_frames, T)
    assert starts == g["clip_starts"].tolist() and len(starts) == B == 146
    ids, mask = W.make_video_text(starts, scenes, T, L, seed=seed)
    return g, T, L, B, sd, frames, starts, ids, mask


def _check(g, logits, precision, T, what):
    """logits [B,2] (any device) against the reference golden: tolerance, labels, timestamps, margin safety."""
    from eval_utils.eval_utils import convert_clip_label2cut_point
    from vcg_b200 import postprocess as pp
    ref = torch.from_numpy(g["logits"])
    got = logits.float().cpu()
    # relative to the magnitude of the reference's logits BEFORE the fixture's additive bias re-centring (the same
    # constant is added on both sides, so it changes neither the computation nor the error, only max|ref|)
    scale = max(float(g["raw_logit_absmax"]), float(ref.abs().max()))
    err = float((got - ref).abs().max() / scale)
    err_recentred = float((got - ref).abs().max() / ref.abs().max())
    m_ref, m_got = (ref[:, 1] - ref[:, 0]).double(), (got[:, 1] - got[:, 0]).double()
    m_err = float((m_got - m_ref).abs().max())
    min_margin = float(m_ref.abs().min())
    hist = np.histogram(m_ref.abs().numpy(), bins=[0, 0.5 * min_margin, min_margin, 2 * min_margin, 4 * min_margin, 8 * min_margin, 1e9])[0]
    print(f"{what} [{precision}]: logits rel err {err:.3e} (tol {TOL[precision]:.0e}; {err_recentred:.3e} against the re-centred logits); margin err {m_err:.3e} vs min |margin| "
          f"{min_margin:.3e} (ratio {m_err / min_margin:.3f}); |margin| histogram in units of min: {hist.tolist()}")
    assert abs(min_margin - float(g["min_abs_margin"])) < 1e-9
    assert err <= TOL[precision], err
    assert m_err < 0.5 * min_margin, (m_err, min_margin)                     # nobody inside the error band
    labels = got.topk(1, 1, True, True)[1].view(-1).tolist()                 # test_whole_pipeline_per_video.py:156-158
    assert labels == g["labels"].tolist()
    assert 0 < sum(labels) < len(labels)                                     # both labels occur
    cuts = convert_clip_label2cut_point(labels, T, 2)                        # :165-166, the mirror's host function
    assert cuts == g["cut_points"].tolist() and len(cuts) >= 3               # IDENTICAL chapter timestamps
    off = torch.tensor([0, len(labels)], dtype=torch.int32)
    dlabels, dcuts = pp.cut_points_device(logits.cuda().float(), off, T, 2)  # and the device peak picker
    assert dlabels.cpu().tolist() == labels and dcuts[0] == cuts
    return err, m_err


def _build_mirror_model(sd, T, precision, vision=True):
    from model.fusion import two_stream
    from model.lang import bert_hugface
    from model.vision import resnet50_tsm
    from ops.basic_ops import Identity
    lang = bert_hugface.BertHugface(pretrain_stage=False)
    if vision:
        vis = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
        model = two_stream.TwoStream(lang.base_model, vis.base_model, lang.embed_size, vis.feature_dim, T, 128)
    else:
        model = two_stream.TwoStream(lang.base_model, Identity(), lang.embed_size, 2048, T, 128)
    model.build_chapter_head(output_size=2, head_type="mlp")
    model.load_state_dict(sd, strict=True)
    model = model.to(0).eval()
    model.precision = precision
    return model


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_video_timestamps_identical_callers_loop(golden_dir, precision):
    """Flow A: exactly what test_whole_pipeline_per_video.py does — DataLoader batches of 16 normalised fp32 clips
    through model(img_clip, text_ids, attention_mask), labels accumulated over the video, then the peak picker."""
    from oracle import two_stream_oracle as orc
    g, T, L, B, sd, frames, starts, ids, mask = _video_case(golden_dir)
    model = _build_mirror_model(sd, T, precision)
    pre = orc.preprocess_u8(frames)                                           # ToTensor + Normalize, as the dataset does
    out = []
    for b0 in range(0, B, 16):
        sl = slice(b0, min(b0 + 16, B))
        img = orc.gather_clips(pre, starts[sl], T).cuda()
        logits, prob = model(img, ids[sl].cuda(), mask[sl].cuda())
        out.append(logits)
    _check(g, torch.cat(out), precision, T, "callers' loop")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_video_timestamps_identical_u8_entry_points(golden_dir, precision):
    """Flows B and C: uint8 frames on the device (shared stem) and the host-buffer C-ABI call."""
    from vcg_b200.engine import Engine
    g, T, L, B, sd, frames, starts, ids, mask = _video_case(golden_dir)
    eng = Engine(T, "mlp", precision, vision=True, max_tokens=L, max_batch=32)
    eng.load_state_dict(sd)
    logits, _ = eng.score_video_u8(frames.cuda(), 0, 4, ids.cuda(), mask.cuda())
    _check(g, logits, precision, T, "score_video_u8")
    st = torch.tensor(starts, dtype=torch.int32)
    logits_h, probs_h = eng.score_clips_u8_host(frames.pin_memory(), st, ids, mask)
    _check(g, logits_h, precision, T, "score_clips_u8_host")
    assert torch.allclose(probs_h, torch.softmax(logits_h, 1), atol=1e-6)
    # per-clip gather path (device-side starts, no shared stem) gives the same timestamps too
    logits_g, _ = eng.score_clips_u8(frames.cuda(), st.cuda(), ids.cuda(), mask.cuda())
    _check(g, logits_g, precision, T, "score_clips_u8")
    eng.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_config1_batch256_matches_reference_golden(golden_dir, precision):
    """BASELINE.json configs[1]: precomputed vision embeddings [256,16,2048], through the unchanged TwoStream.forward
    with vision_model = Identity and img_clip shaped [B,T,2048,1,1]."""
    from oracle import weights as W
    g = np.load(f"{golden_dir}/video_emb_mlp_T16_L100_B256.npz")
    T, L, B, seed = [int(x) for x in g["meta"][:4]]
    sd = W.make_state_dict(T, "mlp", seed=seed, include_vision=False)
    sd["fusion_head.head.bias"] = torch.from_numpy(g["head_bias"]).clone()
    emb, ids, mask = W.make_precomputed_inputs(B, T, L, seed=seed)
    model = _build_mirror_model(sd, T, precision, vision=False)
    logits, probs = model(emb.view(B, T, 2048, 1, 1).cuda(), ids.cuda(), mask.cuda())
    _check(g, logits, precision, T, "configs[1] B=256")
    assert float((probs.cpu() - torch.from_numpy(g["probs"])).abs().max()) <= TOL[precision]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_video_timestamps_identical_attention_head(golden_dir, precision):
    """The same video through the attention ChapterHead (two_stream.py:31-48): reference logits / labels / timestamps from
    the unmodified reference's fusion head on its own embeddings (oracle/make_golden_video.py)."""
    from vcg_b200.engine import Engine
    g, T, L, B, sd, frames, starts, ids, mask = _video_case(golden_dir, "attn")
    eng = Engine(T, "attn", precision, vision=True, max_tokens=L, max_batch=32)
    eng.load_state_dict(sd)
    logits, _ = eng.score_video_u8(frames.cuda(), 0, 4, ids.cuda(), mask.cuda())
    _check(g, logits, precision, T, "score_video_u8, attn head")
    eng.close()


def _sharded_worker(rank, world, port, golden_dir, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "video-chapter-generation_b200"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from vcg_b200 import distributed as vd
    from vcg_b200.engine import Engine
    # one process per rank; both ranks share cuda:0 here (the GPU test box has one GPU), so the all-gather runs on gloo —
    # on a multi-GPU box the same code runs on NCCL (bench.py --gpus N)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    g, T, L, B, sd, frames, starts, ids, mask = _video_case(golden_dir)
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=L, max_batch=16)
    eng.load_state_dict(sd)
    frames_d = frames.cuda()
    st = torch.tensor(starts, dtype=torch.int32)

    def score(lo, hi):          # this rank's contiguous shard of the video's clips
        out = eng.score_clips_u8_host(frames, st[lo:hi].contiguous(), ids[lo:hi].contiguous(), mask[lo:hi].contiguous())
        return out[0].clone()

    logits = vd.score_sharded(score, B)
    torch.save({"logits": logits, "range": vd.shard_range(B, rank, world)}, os.path.join(out_dir, f"r{rank}.pt"))
    del frames_d
    eng.close()
    dist.destroy_process_group()


def test_video_timestamps_identical_two_rank_sharded(golden_dir, tmp_path):
    """Flow D: the clips of the video sharded over two ranks, one all-gather of the [73,2] logits; every rank must derive
    the reference's timestamps."""
    import torch.multiprocessing as mp
    mp.spawn(_sharded_worker, args=(2, _free_port(), golden_dir, str(tmp_path)), nprocs=2, join=True)
    g = np.load(f"{golden_dir}/video_mlp_T16_L100_600f.npz")
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(2)]
    assert r[0]["range"] == (0, 73) and r[1]["range"] == (73, 146)
    assert torch.equal(r[0]["logits"], r[1]["logits"])                         # every rank holds all scores
    _check(g, r[0]["logits"], "bf16", 16, "2-rank sharded")
