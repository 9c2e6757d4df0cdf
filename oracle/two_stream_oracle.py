"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch fp32, functional) of the reference's chapter-boundary scoring
path.  It is the checker for the CUDA path, never the thing measured or shipped: only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import it.

Parity status: PINNED against the reference itself.  ``oracle/make_golden.py`` (run in the build container, where
/root/reference is mounted) imports the unmodified reference modules, loads the same seeded state dict with
``load_state_dict(strict=True)``, and checks every function below against the reference's own forward
(max |delta| / max |ref| <= 1e-5 on logits, embeddings and per-layer taps); the reference's outputs are committed as
tests/golden/*.npz so the pin travels to machines without /root/reference.

Every function cites the reference file:line it restates (paths relative to /root/reference/video_chapter_generation
unless they name the un-vendored dependencies: transformers 5.5.0 modeling_bert.py, torchvision 0.26.0 resnet.py).
"""
import math

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# ----------------------------------------------------------------------------------------------- preprocessing
def preprocess_u8(frames_u8):
    """ToTensor + Normalize (test_video_segment_point.py:142-145; torchvision ToTensor = HWC u8 -> CHW fp32 / 255).
    frames_u8 [n,224,224,3] uint8 -> [n,3,224,224] fp32."""
    x = frames_u8.permute(0, 3, 1, 2).float().div(255.0)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def gather_clips(frames, starts, clip_frames):
    """Clip b = frames[start_b : start_b + T] (infer_youtube_video_dataset.py:117,186-200). -> [B,T,...]"""
    return torch.stack([frames[s:s + clip_frames] for s in starts], dim=0)


# ----------------------------------------------------------------------------------------------- text stream
def bert_forward(sd, ids, mask, prefix="lang_model.", taps=None):
    """transformers BertModel.forward -> pooler_output (modeling_bert.py:628-703), as called at
    model/fusion/two_stream.py:174-179 (token_type_ids = 0, position_ids = arange(L), eval mode: no dropout)."""
    B, L = ids.shape
    p = prefix
    # BertEmbeddings, modeling_bert.py:72-113
    x = sd[p + "embeddings.word_embeddings.weight"][ids] + sd[p + "embeddings.token_type_embeddings.weight"][0]
    x = x + sd[p + "embeddings.position_embeddings.weight"][:L][None]
    x = F.layer_norm(x, (768,), sd[p + "embeddings.LayerNorm.weight"], sd[p + "embeddings.LayerNorm.bias"], 1e-12)
    if taps is not None:
        taps["bert.embeddings"] = x
    # additive key mask: 0 where attention_mask == 1, -inf elsewhere (modeling_bert.py:115-140 via sdpa).  A row whose mask
    # is ALL zero -- the window dataset's padding clips, infer_youtube_video_dataset.py:488-499 -- gets a ZERO attention
    # context: torch's scaled_dot_product_attention (2.11, what the pinned oracle environment runs) uses a "safe softmax"
    # that returns 0 for fully masked rows instead of NaN.  Verified against transformers 5.5 BertModel on CPU.
    add = torch.zeros(B, 1, 1, L, device=ids.device).masked_fill(mask[:, None, None, :] == 0, float("-inf"))
    n_layers = 0
    while f"{p}encoder.layer.{n_layers}.attention.self.query.weight" in sd:
        n_layers += 1
    for i in range(n_layers):
        q_ = f"{p}encoder.layer.{i}."
        # BertSelfAttention, modeling_bert.py:168-207 (12 heads x 64, scale 1/8)
        def heads(t):
            return t.view(B, L, 12, 64).transpose(1, 2)
        q = heads(F.linear(x, sd[q_ + "attention.self.query.weight"], sd[q_ + "attention.self.query.bias"]))
        k = heads(F.linear(x, sd[q_ + "attention.self.key.weight"], sd[q_ + "attention.self.key.bias"]))
        v = heads(F.linear(x, sd[q_ + "attention.self.value.weight"], sd[q_ + "attention.self.value.bias"]))
        att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(64.0) + add, dim=-1)
        att = torch.nan_to_num(att, nan=0.0)      # fully masked rows: safe softmax -> 0
        ctx = (att @ v).transpose(1, 2).reshape(B, L, 768)
        # BertSelfOutput, modeling_bert.py:294-298
        y = F.linear(ctx, sd[q_ + "attention.output.dense.weight"], sd[q_ + "attention.output.dense.bias"])
        x = F.layer_norm(y + x, (768,), sd[q_ + "attention.output.LayerNorm.weight"],
                         sd[q_ + "attention.output.LayerNorm.bias"], 1e-12)
        # BertIntermediate (erf GELU), modeling_bert.py:339-342; BertOutput :352-356
        h = F.gelu(F.linear(x, sd[q_ + "intermediate.dense.weight"], sd[q_ + "intermediate.dense.bias"]))
        y = F.linear(h, sd[q_ + "output.dense.weight"], sd[q_ + "output.dense.bias"])
        x = F.layer_norm(y + x, (768,), sd[q_ + "output.LayerNorm.weight"], sd[q_ + "output.LayerNorm.bias"], 1e-12)
        if taps is not None:
            taps[f"bert.layer{i}"] = x
    # BertPooler, modeling_bert.py:456-468
    return torch.tanh(F.linear(x[:, 0], sd[p + "pooler.dense.weight"], sd[p + "pooler.dense.bias"]))


# ----------------------------------------------------------------------------------------------- vision stream
def temporal_shift(x, n_segment, fold_div=8):
    """TemporalShift.shift, ops/temporal_shift.py:34-51 (out-of-place branch)."""
    nt, c, h, w = x.shape
    x = x.view(nt // n_segment, n_segment, c, h, w)
    fold = c // fold_div
    out = torch.zeros_like(x)
    out[:, :-1, :fold] = x[:, 1:, :fold]
    out[:, 1:, fold:2 * fold] = x[:, :-1, fold:2 * fold]
    out[:, :, 2 * fold:] = x[:, :, 2 * fold:]
    return out.view(nt, c, h, w)


BN_BATCH_STATS = False   # set by batch_stat_bn(): restates caller #1's quirk, test_video_segment_point.py:116-122


class batch_stat_bn:
    """Context manager: BatchNorm2d normalises with the statistics of the batch it is given — what an eval-mode
    nn.BatchNorm2d does once running_mean / running_var are None and track_running_stats is False
    (torch.nn.modules.batchnorm._BatchNorm.forward: bn_training = running_mean is None and running_var is None)."""

    def __enter__(self):
        global BN_BATCH_STATS
        self.prev, BN_BATCH_STATS = BN_BATCH_STATS, True

    def __exit__(self, *exc):
        global BN_BATCH_STATS
        BN_BATCH_STATS = self.prev


def _bn(sd, prefix, x):
    """BatchNorm2d in eval mode with running statistics (standard .eval(): test_whole_pipeline_per_video.py:105), or with
    batch statistics under batch_stat_bn()."""
    if BN_BATCH_STATS:
        return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.0, 1e-5)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.0, 1e-5)


def resnet50_tsm_forward(sd, x, n_segment, shift_div=8, prefix="vision_model.", taps=None):
    """torchvision ResNet._forward_impl with Bottleneck v1.5 (stride on the 3x3), fc = Identity
    (model/vision/resnet50_tsm.py:15-19), TemporalShift in front of every conv1 when the keys carry '.net.'
    (ops/temporal_shift.py:127-144).  x [B*T,3,224,224] -> [B*T,2048]."""
    p = prefix
    x = F.relu(_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"], stride=2, padding=3)))
    if taps is not None:
        taps["vision.stem"] = x
    x = F.max_pool2d(x, 3, 2, 1)
    for stage, blocks in enumerate((3, 4, 6, 3), start=1):
        for i in range(blocks):
            b = f"{p}layer{stage}.{i}."
            stride = 2 if (i == 0 and stage > 1) else 1
            identity = x
            if b + "conv1.net.weight" in sd:
                out = F.conv2d(temporal_shift(x, n_segment, shift_div), sd[b + "conv1.net.weight"])
            else:
                out = F.conv2d(x, sd[b + "conv1.weight"])
            out = F.relu(_bn(sd, b + "bn1", out))
            out = F.relu(_bn(sd, b + "bn2", F.conv2d(out, sd[b + "conv2.weight"], stride=stride, padding=1)))
            out = _bn(sd, b + "bn3", F.conv2d(out, sd[b + "conv3.weight"]))
            if b + "downsample.0.weight" in sd:
                identity = _bn(sd, b + "downsample.1", F.conv2d(x, sd[b + "downsample.0.weight"], stride=stride))
            x = F.relu(out + identity)
        if taps is not None:
            taps[f"vision.layer{stage}"] = x
    return F.adaptive_avg_pool2d(x, 1).flatten(1)


# ----------------------------------------------------------------------------------------------- fusion head
def chapter_head(sd, lang_emb, vision_emb, segment_size, hidden_size=128, head_type="mlp", prefix="fusion_head."):
    """ChapterHead.forward, model/fusion/two_stream.py:71-95; SelfAttention.forward :31-48 for head_type='attn'."""
    p = prefix
    B = lang_emb.shape[0]
    lang_out = F.relu(F.linear(lang_emb, sd[p + "lang_proj_head.weight"])).unsqueeze(1)
    vis = F.linear(vision_emb.reshape(-1, vision_emb.shape[-1]), sd[p + "vision_proj_head.weight"])
    vis = F.relu(vis.view(B, segment_size, hidden_size))
    fusion = torch.cat([vis, lang_out], dim=1)
    if head_type == "mlp":
        return F.linear(fusion.reshape(B, -1), sd[p + "head.weight"], sd[p + "head.bias"])
    if head_type != "attn":
        raise RuntimeError(f"Unknown head_type {head_type}")
    Tn, C, nh = fusion.shape[1], hidden_size, 4
    k = F.linear(fusion, sd[p + "head.key.weight"], sd[p + "head.key.bias"]).view(B, Tn, nh, C // nh).transpose(1, 2)
    q = F.linear(fusion, sd[p + "head.query.weight"], sd[p + "head.query.bias"]).view(B, Tn, nh, C // nh).transpose(1, 2)
    v = F.linear(fusion, sd[p + "head.value.weight"], sd[p + "head.value.bias"]).view(B, Tn, nh, C // nh).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(k.size(-1))), dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, Tn, C)
    return F.linear(y[:, 0, :], sd[p + "head.proj.weight"], sd[p + "head.proj.bias"])


def two_stream_forward(sd, img_clip, text_ids, attention_mask, segment_size, hidden_size=128, head_type="mlp",
                       shift_div=8, vision_emb=None, taps=None):
    """TwoStream.forward, model/fusion/two_stream.py:172-194 -> (logits, probs, vision_emb, lang_emb).
    ``vision_emb`` given = the precomputed-embedding configuration (vision_model = Identity)."""
    lang_emb = bert_forward(sd, text_ids, attention_mask, taps=taps)
    if vision_emb is None:
        B = img_clip.shape[0]
        x = img_clip.reshape(B * segment_size, *img_clip.shape[2:]).contiguous()
        vision_emb = resnet50_tsm_forward(sd, x, segment_size, shift_div, taps=taps).view(B, segment_size, -1)
    logits = chapter_head(sd, lang_emb, vision_emb, segment_size, hidden_size, head_type)
    return logits, F.softmax(logits, dim=1), vision_emb, lang_emb


# ----------------------------------------------------------------------------------------------- single modality
def vision_only_forward(sd, img_clip, segment_size, shift_div=8):
    """Resnet50TSM.forward / Resnet50.forward (model/vision/resnet50_tsm.py:68-77, resnet50.py:64-73):
    backbone -> view [B, T*2048] -> head Linear -> softmax.  shift_div = 0 (or plain conv1 keys): no temporal shift."""
    B = img_clip.shape[0]
    x = img_clip.reshape(B * segment_size, *img_clip.shape[2:]).contiguous()
    emb = resnet50_tsm_forward(sd, x, segment_size, shift_div, prefix="base_model.").view(B, -1)
    logits = F.linear(emb, sd["head.weight"], sd["head.bias"])
    return logits, F.softmax(logits, dim=1), emb.view(B, segment_size, -1)


def text_only_forward(sd, text_ids, attention_mask):
    """BertHugface.forward with pretrain_stage=False (model/lang/bert_hugface.py:98-132): pooler output -> head."""
    pooled = bert_forward(sd, text_ids, attention_mask, prefix="base_model.")
    logits = F.linear(pooled, sd["head.weight"], sd["head.bias"])
    return logits, F.softmax(logits, dim=1), pooled


# ----------------------------------------------------------------------------------------------- post-processing
def predict_labels(logits):
    """pred_label = logits.topk(1) index (test_video_segment_point.py:201-203)."""
    return logits.topk(1, 1, True, True)[1].view(-1).tolist()


def convert_clip_label2cut_point(clip_label_array, clip_frame_num, max_offset):
    """eval_utils/eval_utils.py:3-18: midpoint of every maximal run of 1-labels that is followed by a 0."""
    enter = False
    begin_sec = 0
    cut_points = []
    for i, lab in enumerate(clip_label_array):
        if lab == 1 and not enter:
            enter = True
            begin_sec = i * max_offset * 2
        if lab == 0 and enter:
            enter = False
            end_sec = (i - 1) * max_offset * 2 + clip_frame_num
            cut_points.append(round((begin_sec + end_sec - 1) / 2))
    return cut_points


def calculate_pr(gt_cut_points, pred_cut_points):
    """eval_utils/eval_utils.py:21-92: recall / precision at 0, 3 and 5 s tolerance."""
    def hits(a_list, b_list):
        h0 = h3 = h5 = 0
        for a in a_list:
            h0 += any(a == b for b in b_list)
            h3 += any(a - 3 <= b <= a + 3 for b in b_list)
            h5 += any(a - 5 <= b <= a + 5 for b in b_list)
        return h0, h3, h5
    n = len(gt_cut_points)
    r0, r3, r5 = hits(gt_cut_points, pred_cut_points)
    recall = (r0 / n, r3 / n, r5 / n)   # ZeroDivisionError on an empty ground truth, like the reference (:51)
    precision = (None, None, None)
    if len(pred_cut_points) > 0:
        m = len(pred_cut_points)
        h0 = h3 = h5 = 0
        for pc in pred_cut_points:
            h0 += any(pc == g for g in gt_cut_points)
            h3 += any(g - 3 <= pc <= g + 3 for g in gt_cut_points)
            h5 += any(g - 5 <= pc <= g + 5 for g in gt_cut_points)
        precision = (h0 / m, h3 / m, h5 / m)
    return recall + precision
