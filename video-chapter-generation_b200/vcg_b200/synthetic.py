"""Deterministic random weights and synthetic inputs for the chapter-boundary scorer (no arithmetic of the path).

The reference ships no checkpoint and there is no network, so parity runs on random-init weights.  This module builds
a state dict with exactly the key schema and shapes of the reference's ``TwoStream.state_dict()``
(video_chapter_generation/model/fusion/two_stream.py:99-124 wiring ``BertModel`` + torchvision ResNet-50 with
``TemporalShift`` wrappers (ops/temporal_shift.py:138 -> ``conv1.net.weight``) + ``ChapterHead``), from a seeded CPU
generator, so the same tensors can be regenerated on any machine with the same torch build instead of committing
533 MB of weights.  ``oracle/make_golden.py`` loads this dict into the unmodified reference with
``load_state_dict(strict=True)``, which is what pins the schema.

BatchNorm statistics/affine and LayerNorm affine are randomised so that activations stay O(1) through
the 16 residual blocks and every parameter influences the output.

Data generation only: bench.py, the tools and the tests all draw their synthetic workloads from here; oracle/weights.py
re-exports it for the oracle-side scripts.
"""
import math

import torch

BERT_HIDDEN = 768
BERT_FFN = 3072
VOCAB = 30522
MAX_POS = 512


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def make_state_dict(clip_frames=16, head_type="mlp", hidden_size=128, seed=123, bert_layers=12, tsm=True,
                    include_vision=True):
    """Returns an OrderedDict-like dict of fp32 CPU tensors keyed like the reference TwoStream.state_dict()."""
    g = _gen(seed)
    sd = {}

    def normal(shape, std):
        return torch.randn(shape, generator=g) * std

    def uniform(shape, lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    # ---------------- text stream: transformers BertModel (bert-base-uncased geometry)
    lm = "lang_model."
    sd[lm + "embeddings.word_embeddings.weight"] = normal((VOCAB, BERT_HIDDEN), 0.02)
    sd[lm + "embeddings.position_embeddings.weight"] = normal((MAX_POS, BERT_HIDDEN), 0.02)
    sd[lm + "embeddings.token_type_embeddings.weight"] = normal((2, BERT_HIDDEN), 0.02)
    sd[lm + "embeddings.LayerNorm.weight"] = uniform((BERT_HIDDEN,), 0.8, 1.2)
    sd[lm + "embeddings.LayerNorm.bias"] = normal((BERT_HIDDEN,), 0.05)
    for i in range(bert_layers):
        p = f"{lm}encoder.layer.{i}."
        for name in ("query", "key", "value"):
            # std 0.05 (not 0.02) so that attention is not uniform and the softmax path is really exercised
            sd[p + f"attention.self.{name}.weight"] = normal((BERT_HIDDEN, BERT_HIDDEN), 0.05)
            sd[p + f"attention.self.{name}.bias"] = normal((BERT_HIDDEN,), 0.02)
        sd[p + "attention.output.dense.weight"] = normal((BERT_HIDDEN, BERT_HIDDEN), 0.02)
        sd[p + "attention.output.dense.bias"] = normal((BERT_HIDDEN,), 0.02)
        sd[p + "attention.output.LayerNorm.weight"] = uniform((BERT_HIDDEN,), 0.8, 1.2)
        sd[p + "attention.output.LayerNorm.bias"] = normal((BERT_HIDDEN,), 0.05)
        sd[p + "intermediate.dense.weight"] = normal((BERT_FFN, BERT_HIDDEN), 0.02)
        sd[p + "intermediate.dense.bias"] = normal((BERT_FFN,), 0.02)
        sd[p + "output.dense.weight"] = normal((BERT_HIDDEN, BERT_FFN), 0.02)
        sd[p + "output.dense.bias"] = normal((BERT_HIDDEN,), 0.02)
        sd[p + "output.LayerNorm.weight"] = uniform((BERT_HIDDEN,), 0.8, 1.2)
        sd[p + "output.LayerNorm.bias"] = normal((BERT_HIDDEN,), 0.05)
    sd[lm + "pooler.dense.weight"] = normal((BERT_HIDDEN, BERT_HIDDEN), 0.02)
    sd[lm + "pooler.dense.bias"] = normal((BERT_HIDDEN,), 0.02)

    # ---------------- vision stream: torchvision resnet50 (v1.5), fc = Identity, TemporalShift on every conv1
    if include_vision:
        vm = "vision_model."

        def conv(key, cout, cin, k):
            sd[key] = normal((cout, cin, k, k), math.sqrt(2.0 / (cout * k * k)))   # kaiming_normal_(fan_out, relu)

        def bn(prefix, c, last=False):
            # the last BN of a block gets a smaller gain so the residual sum does not blow up over 16 blocks
            sd[prefix + ".weight"] = uniform((c,), 0.2, 0.6) if last else uniform((c,), 0.5, 1.5)
            sd[prefix + ".bias"] = normal((c,), 0.1)
            sd[prefix + ".running_mean"] = normal((c,), 0.1)
            sd[prefix + ".running_var"] = uniform((c,), 0.5, 1.5)
            sd[prefix + ".num_batches_tracked"] = torch.tensor(100, dtype=torch.long)

        conv(vm + "conv1.weight", 64, 3, 7)
        bn(vm + "bn1", 64)
        inplanes = 64
        for stage, (blocks, planes) in enumerate(zip((3, 4, 6, 3), (64, 128, 256, 512)), start=1):
            for i in range(blocks):
                p = f"{vm}layer{stage}.{i}."
                conv(p + ("conv1.net.weight" if tsm else "conv1.weight"), planes, inplanes, 1)
                bn(p + "bn1", planes)
                conv(p + "conv2.weight", planes, planes, 3)
                bn(p + "bn2", planes)
                conv(p + "conv3.weight", planes * 4, planes, 1)
                bn(p + "bn3", planes * 4, last=True)
                if i == 0:
                    conv(p + "downsample.0.weight", planes * 4, inplanes, 1)
                    bn(p + "downsample.1", planes * 4)
                inplanes = planes * 4

    # ---------------- fusion head (two_stream.py:51-68)
    fh = "fusion_head."

    def linear_w(out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        return uniform((out_f, in_f), -b, b)

    sd[fh + "lang_proj_head.weight"] = linear_w(hidden_size, BERT_HIDDEN)
    sd[fh + "vision_proj_head.weight"] = linear_w(hidden_size, 2048)
    if head_type == "mlp":
        n_in = (clip_frames + 1) * hidden_size
        sd[fh + "head.weight"] = linear_w(2, n_in)
        sd[fh + "head.bias"] = uniform((2,), -1.0 / math.sqrt(n_in), 1.0 / math.sqrt(n_in))
    elif head_type == "attn":
        for name in ("key", "query", "value"):
            sd[fh + f"head.{name}.weight"] = linear_w(hidden_size, hidden_size)
            sd[fh + f"head.{name}.bias"] = uniform((hidden_size,), -0.08, 0.08)
        sd[fh + "head.proj.weight"] = linear_w(2, hidden_size)
        sd[fh + "head.proj.bias"] = uniform((2,), -0.08, 0.08)
    else:
        raise RuntimeError(f"Unknown head_type {head_type}")
    return sd


def make_unimodal_state_dict(kind, clip_frames=16, seed=123):
    """State dict of a single-modality reference model (--data_mode image / text), keyed like the reference modules:
      kind "r50tsm": Resnet50TSM (base_model.* with '.conv1.net.' keys, head Linear(T*2048, 2))
      kind "r50":    Resnet50    (base_model.*, plain conv1 keys, same head)
      kind "bert":   BertHugface (base_model.* = BertModel keys, head Linear(768, 2))
    The backbone values are those of make_state_dict (same seed), so the backbones agree with the two-stream cases."""
    full = make_state_dict(clip_frames, "mlp", seed=seed, tsm=(kind != "r50"), include_vision=(kind != "bert"))
    g = _gen(seed + 77)
    sd = {}
    if kind in ("r50tsm", "r50"):
        for k, v in full.items():
            if k.startswith("vision_model."):
                sd["base_model." + k[len("vision_model."):]] = v
        d = clip_frames * 2048
    elif kind == "bert":
        for k, v in full.items():
            if k.startswith("lang_model."):
                sd["base_model." + k[len("lang_model."):]] = v
        d = 768
    else:
        raise ValueError(kind)
    bound = 1.0 / d ** 0.5                       # nn.Linear default init range
    sd["head.weight"] = (torch.rand(2, d, generator=g) * 2 - 1) * bound * 4   # x4: keeps the two logits well apart
    sd["head.bias"] = (torch.rand(2, generator=g) * 2 - 1) * bound
    return sd


def make_window_state_dict(clip_frames=8, window_size=1, head_type="cross_attn", hidden_size=128, seed=123):
    """State dict of the reference's window model (model/fusion/two_stream_window.py TwoStream + build_chapter_head):
    backbones as in make_state_dict (same seed) plus window_mlp.*, fusion_head.* (per-position heads) and window_attn.*.
    Linear weights ~ U(+-1.5/sqrt(in)), biases ~ U(+-0.1), LayerNorm weight ~ U(0.5, 1.5), bias ~ U(+-0.1)."""
    full = make_state_dict(clip_frames, "mlp", hidden_size, seed=seed)
    sd = {k: v for k, v in full.items() if not k.startswith("fusion_head.")}
    g = _gen(seed + 501)
    h, W = hidden_size, 2 * window_size + 1

    def u(shape, a):
        return (torch.rand(shape, generator=g) * 2 - 1) * a

    def linear(prefix, d_in, d_out):
        sd[prefix + ".weight"] = u((d_out, d_in), 1.5 / d_in ** 0.5)
        sd[prefix + ".bias"] = u((d_out,), 0.1)

    def norm(prefix, d):
        sd[prefix + ".weight"] = torch.rand(d, generator=g) + 0.5
        sd[prefix + ".bias"] = u((d,), 0.1)

    def seq(prefix, dims, norm_last=False):
        """nn.Sequential(Linear, LayerNorm, act, Dropout, ...): module index 4*j for Linear j, 4*j+1 for its LayerNorm"""
        for j in range(len(dims) - 1):
            linear(f"{prefix}.{4 * j}", dims[j], dims[j + 1])
            if j < len(dims) - 2 or norm_last:
                norm(f"{prefix}.{4 * j + 1}", dims[j + 1])

    seq("window_mlp", [h * W, h, h // 2, h // 4, h // 8, h // 16, 2])
    for i in range(W):
        seq(f"fusion_head.lang_proj_heads.{i}", [768, 384, h])
        seq(f"fusion_head.vision_proj_heads.{i}", [2048, 8 * h, 4 * h, h])
        if head_type == "mlp":
            seq(f"fusion_head.head.{i}", [(clip_frames + 1) * h, 8 * h, 4 * h, h])
        elif head_type == "bilinear":       # nn.Bilinear(h, T*h, 2h) + Sequential(LN, ReLU, Drop, Linear, LN, ReLU, Drop, Linear)
            sd[f"fusion_head.bilinear_layers.{i}.weight"] = u((2 * h, h, clip_frames * h), 1.5 / (h * clip_frames * h) ** 0.5)
            sd[f"fusion_head.bilinear_layers.{i}.bias"] = u((2 * h,), 0.1)
            norm(f"fusion_head.head.{i}.0", 2 * h)
            linear(f"fusion_head.head.{i}.3", 2 * h, h)
            norm(f"fusion_head.head.{i}.4", h)
            linear(f"fusion_head.head.{i}.7", h, h)
        elif head_type == "multiplication":
            seq(f"fusion_head.lang_expand_layers.{i}", [h, 8 * h, clip_frames * h], norm_last=True)
            seq(f"fusion_head.head.{i}", [clip_frames * h, 8 * h, 4 * h, h])
    if head_type == "self_attn":
        for n in ("key", "query", "value", "proj"):
            linear(f"fusion_head.head.{n}", h, h)
    if head_type == "cross_attn":
        for n in ("query_proj", "key_proj", "value_proj", "out_proj"):
            linear(f"fusion_head.head.{n}", h, h)
        norm("fusion_head.head.lang_norm", h)
        norm("fusion_head.head.vision_norm", h)
        linear("fusion_head.head.frame_pos_encoding", 1, h)
        linear("fusion_head.output_proj", h, 2)
    for l in range(6):
        p = f"window_attn.layers.{l}"
        norm(p + ".attention_norm", h)
        norm(p + ".ffn_norm", h)
        for n in ("query", "key", "value", "out_proj"):
            linear(f"{p}.attention.{n}", h, h)
        linear(p + ".attention.position_encoding", 1, h)
        sd[p + ".attention.window_pos_bias"] = u((1, 16, 1, W), 0.5)
        for j, (a, b) in enumerate(((h, 2 * h), (2 * h, 4 * h), (4 * h, 2 * h), (2 * h, h))):
            linear(f"{p}.ffn.{3 * j}", a, b)      # Sequential(Linear, GELU, Dropout, ...): Linear j at index 3*j
    norm("window_attn.final_layer_norm", h)
    seq("window_attn.classifier", [h, h, h, h // 2, h // 4, 2])
    return sd


def _param_factory(sd, g):
    def u(shape, a):
        return (torch.rand(shape, generator=g) * 2 - 1) * a

    def linear(prefix, d_in, d_out):
        sd[prefix + ".weight"] = u((d_out, d_in), 1.5 / d_in ** 0.5)
        sd[prefix + ".bias"] = u((d_out,), 0.1)

    def norm(prefix, d):
        sd[prefix + ".weight"] = torch.rand(d, generator=g) + 0.5
        sd[prefix + ".bias"] = u((d,), 0.1)

    def seq(prefix, dims, step=4):
        """nn.Sequential(Linear, LayerNorm, act, Dropout, ..., Linear): Linear j at step*j, its LayerNorm at step*j+1"""
        for j in range(len(dims) - 1):
            linear(f"{prefix}.{step * j}", dims[j], dims[j + 1])
            if j < len(dims) - 2:
                norm(f"{prefix}.{step * j + 1}", dims[j + 1])
    return u, linear, norm, seq


def make_domain_state_dict(clip_frames=8, window_size=1, hidden_size=128, seed=123):
    """State dict of the reference's model/fusion/two_stream_domain_specific.py TwoStream + build_chapter_head:
    backbones as in make_state_dict (same seed) plus fusion_head.{lang,vision}_proj_heads, the two WindowSelfAttention
    blocks, the (unused) cross_attn member and the classifier."""
    full = make_state_dict(clip_frames, "mlp", hidden_size, seed=seed)
    sd = {k: v for k, v in full.items() if not k.startswith("fusion_head.")}
    u, linear, norm, seq = _param_factory(sd, _gen(seed + 701))
    h, W = hidden_size, 2 * window_size + 1
    for i in range(W):
        seq(f"fusion_head.lang_proj_heads.{i}", [768, 384, h])
        seq(f"fusion_head.vision_proj_heads.{i}", [2048, 8 * h, 4 * h, h])
    for name in ("lang_window_attn", "vision_window_attn", "cross_attn"):
        p = f"fusion_head.{name}"
        for n in ("query_proj", "key_proj", "value_proj"):
            linear(f"{p}.{n}", h, h)
        seq(f"{p}.out_proj", [h, 2 * h, 2 * h, 2 * h, h])
        if name == "cross_attn":
            norm(p + ".vision_norm", h)
            norm(p + ".lang_norm", h)
        else:
            norm(p + ".norm", h)
            sd[p + ".window_pos_bias"] = u((1, 16, W, W), 0.5)
            linear(p + ".position_encoding.0", 1, h)
            norm(p + ".position_encoding.1", h)
    seq("fusion_head.classifier", [2 * h, 2 * h, h, h // 2, h // 4, 2])
    return sd


def make_single_block_state_dict(window_size=1, hidden_size=128, seed=123):
    """State dict of the reference's model/fusion/window_self_attention.py VideoChapterClassifier."""
    sd = {}
    u, linear, norm, seq = _param_factory(sd, _gen(seed + 801))
    h, W = hidden_size, 2 * window_size + 1
    b = "window_block"
    norm(b + ".attention_norm", h)
    norm(b + ".ffn_norm", h)
    for n in ("query", "key", "value", "out_proj"):
        linear(f"{b}.attention.{n}", h, h)
    linear(b + ".attention.position_encoding.0", 1, h)
    norm(b + ".attention.position_encoding.1", h)
    sd[b + ".attention.window_pos_bias"] = u((1, 16, 1, W), 0.5)
    linear(b + ".ffn.1", h, 4 * h)
    linear(b + ".ffn.4", 4 * h, h)
    norm("classifier.0", h)
    linear("classifier.1", h, h // 2)
    linear("classifier.4", h // 2, 2)
    return sd


def make_text(batch, max_len, seed=123):
    """Synthetic token ids / attention mask: [CLS]=101 first, len ~ U{10..L}, pad id 0."""
    g = _gen(seed + 1)
    lo = min(10, max_len)
    lens = torch.randint(lo, max_len + 1, (batch,), generator=g)
    ids = torch.randint(1000, VOCAB, (batch, max_len), generator=g)
    ids[:, 0] = 101
    ar = torch.arange(max_len)[None, :]
    mask = (ar < lens[:, None]).long()
    ids = ids * mask
    return ids, mask


def make_precomputed_inputs(batch, clip_frames, max_len, seed=123, full_length=False):
    """BASELINE.json configs[1] inputs: precomputed vision embeddings (post-ReLU average-pooled features are
    non-negative and O(1): uniform [0,2)) + text.  full_length=True: every attention mask is all ones."""
    ids, mask = make_text(batch, max_len, seed=seed)
    if full_length:
        g2 = _gen(seed + 11)
        ids = torch.randint(1000, VOCAB, (batch, max_len), generator=g2)
        ids[:, 0] = 101
        mask = torch.ones_like(mask)
    g = _gen(seed + 7)
    emb = torch.rand(batch, clip_frames, 2048, generator=g) * 2.0
    return emb, ids, mask


def make_frames_u8(n_frames, seed=123):
    """uint8 HWC frames [n,224,224,3], uniform in [0,255]."""
    g = _gen(seed + 2)
    return torch.randint(0, 256, (n_frames, 224, 224, 3), generator=g, dtype=torch.uint8)


def make_video_u8(n_frames, seed=123, scene_min=24, scene_max=72, noise=16):
    """A synthetic VIDEO (not i.i.d. noise): scenes of U{scene_min..scene_max} frames.  A scene has its own base colour
    (uniform over the frame, U{0..255} per channel: scenes differ in brightness and hue the way shots do) and its own
    7x7 mosaic of 32-pixel cells (+-48 grey levels around the base); every frame adds uniform noise of +-``noise``.
    Neighbouring clips therefore see related frames and a scene change moves the vision embeddings, so that the scores
    of a video form RUNS of labels (what the reference's peak picker, eval_utils/eval_utils.py:3-18, turns into
    timestamps).  Integer arithmetic only -> identical on every machine.
    -> (frames [n,224,224,3] uint8, scene start indices)."""
    g = _gen(seed + 3)
    frames = torch.empty(n_frames, 224, 224, 3, dtype=torch.uint8)
    starts, f = [], 0
    while f < n_frames:
        n = int(torch.randint(scene_min, scene_max + 1, (1,), generator=g))
        n = min(n, n_frames - f)
        colour = torch.randint(0, 256, (1, 1, 3), generator=g, dtype=torch.int16)
        mosaic = torch.randint(-48, 49, (7, 7, 3), generator=g, dtype=torch.int16) + colour
        base = mosaic.repeat_interleave(32, 0).repeat_interleave(32, 1)                       # [224,224,3]
        jitter = torch.randint(-noise, noise + 1, (n, 224, 224, 3), generator=g, dtype=torch.int16)
        frames[f:f + n] = (base[None] + jitter).clamp_(0, 255).to(torch.uint8)
        starts.append(f)
        f += n
    return frames, starts


def make_video_text(starts, scene_starts, clip_frames, max_len, seed=123):
    """Subtitle tokens of the clips of a synthetic video: the subtitles of a real video change with its shots, and
    neighbouring clips (stride 4 s, window T s) see mostly the same words.  Every scene of make_video_u8 gets one token
    sequence of U{10..max_len} tokens ([CLS] first, pad id 0); a clip carries the text of the scene its centre frame
    lies in.  -> (ids [B,L] int64, attention_mask [B,L] int64)."""
    import bisect
    n_scenes = len(scene_starts)
    s_ids, s_mask = make_text(n_scenes, max_len, seed=seed + 5)
    which = [max(0, bisect.bisect_right(scene_starts, st + clip_frames // 2) - 1) for st in starts]
    idx = torch.tensor(which, dtype=torch.long)
    return s_ids[idx].contiguous(), s_mask[idx].contiguous()


def clip_starts(n_frames, clip_frames, stride=4):
    """Candidate clips of a video: range(0, n_frames - T, stride) (infer_youtube_video_dataset.py:117)."""
    return list(range(0, n_frames - clip_frames, stride))
