"""TEST INFRASTRUCTURE — CPU restatement (plain PyTorch fp32, functional) of the reference's window model:
model/fusion/two_stream_window.py (ChapterHead :134-290, CrossAttention :11-88, TwoStream.forward :392-445) and
model/fusion/stacked_window_self_attention.py (:6-224), plus the two unused variants model/fusion/two_stream_domain_specific.py
and model/fusion/window_self_attention.py.  Pinned against the unmodified reference by
oracle/make_golden_window.py.  Only tests may import this."""
import math

import torch
import torch.nn.functional as F

from oracle import two_stream_oracle as orc


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _mlp(sd, prefix, x, n_linear, act=F.relu, step=4):
    """nn.Sequential(Linear, LayerNorm, act, Dropout) x (n-1) + Linear (two_stream_window.py:146-185)."""
    for j in range(n_linear):
        x = _lin(sd, f"{prefix}.{step * j}", x)
        if j < n_linear - 1:
            x = act(_ln(sd, f"{prefix}.{step * j + 1}", x))
    return x


def cross_attention(sd, p, lang, vision, num_heads=16):
    """CrossAttention.forward, two_stream_window.py:53-88: lang [B,H] queries vision [B,T,H]."""
    B, T, H = vision.shape
    hd = H // num_heads
    lang = _ln(sd, p + ".lang_norm", lang)
    vision = _ln(sd, p + ".vision_norm", vision)
    pos = (torch.arange(T).float() / (T - 1)).unsqueeze(-1)
    vision = vision + _lin(sd, p + ".frame_pos_encoding", pos)
    q = _lin(sd, p + ".query_proj", lang).view(B, 1, num_heads, hd).transpose(1, 2)
    k = _lin(sd, p + ".key_proj", vision).view(B, T, num_heads, hd).transpose(1, 2)
    v = _lin(sd, p + ".value_proj", vision).view(B, T, num_heads, hd).transpose(1, 2)
    att = F.softmax(q @ k.transpose(-2, -1) / math.sqrt(hd), dim=-1)
    ctx = (att @ v).transpose(1, 2).contiguous().view(B, 1, H)
    return _lin(sd, p + ".out_proj", ctx).squeeze(1)


def self_attention_first(sd, p, x, num_heads=4):
    """SelfAttention.forward, two_stream_window.py:114-131: unmasked attention over the tokens of x [B,N,C]; only the
    first token's context goes through the output projection."""
    B, N, C = x.shape
    hd = C // num_heads
    q, k, v = (_lin(sd, f"{p}.{nm}", x).view(B, N, num_heads, hd).transpose(1, 2) for nm in ("query", "key", "value"))
    att = F.softmax(q @ k.transpose(-2, -1) * (1.0 / math.sqrt(hd)), dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, N, C)
    return _lin(sd, p + ".proj", y[:, 0, :])


def chapter_head(sd, lang_emb, vision_emb, i, T, head_type, H=128):
    """ChapterHead.forward for window position i, two_stream_window.py:252-290 -> fusion_emb [B,H]."""
    B = lang_emb.shape[0]
    lang_out = F.relu(_mlp(sd, f"fusion_head.lang_proj_heads.{i}", lang_emb, 2))
    vis_out = F.relu(_mlp(sd, f"fusion_head.vision_proj_heads.{i}", vision_emb.reshape(-1, vision_emb.shape[-1]), 3))
    vis_out = vis_out.view(B, T, H)
    if head_type == "mlp":
        x = torch.cat([vis_out, lang_out.unsqueeze(1)], dim=1).view(B, -1)
        return _mlp(sd, f"fusion_head.head.{i}", x, 3)
    if head_type == "cross_attn":
        return cross_attention(sd, "fusion_head.head", lang_out, vis_out)
    if head_type == "bilinear":           # :269-272
        p = f"fusion_head.bilinear_layers.{i}"
        x = F.bilinear(lang_out, vis_out.view(B, -1), sd[p + ".weight"], sd[p + ".bias"])
        h = f"fusion_head.head.{i}"
        x = _lin(sd, h + ".3", F.relu(_ln(sd, h + ".0", x)))
        return _lin(sd, h + ".7", F.relu(_ln(sd, h + ".4", x)))
    if head_type == "multiplication":     # :274-279
        e = f"fusion_head.lang_expand_layers.{i}"
        x = F.relu(_ln(sd, e + ".1", _lin(sd, e + ".0", lang_out)))
        x = F.relu(_ln(sd, e + ".5", _lin(sd, e + ".4", x)))
        return _mlp(sd, f"fusion_head.head.{i}", (vis_out * x.view(B, T, H)).view(B, -1), 3)
    if head_type == "self_attn":          # :281-283 with SelfAttention.forward :114-131 (4 heads, first token out)
        return self_attention_first(sd, "fusion_head.head", torch.cat([vis_out, lang_out.unsqueeze(1)], dim=1))
    raise RuntimeError(f"Unknown head_type {head_type}")


def window_stack(sd, x, num_layers=6, num_heads=16):
    """StackedVideoChapterAttention.forward (stacked_window_self_attention.py:203-223): x [B,W,H] -> logits, probs."""
    B, W, H = x.shape
    hd, mid = H // num_heads, W // 2
    for l in range(num_layers):
        p = f"window_attn.layers.{l}"
        n = _ln(sd, p + ".attention_norm", x)
        pos = ((torch.arange(W) - mid).float() / (mid + 1e-6)).unsqueeze(-1)
        n = n + _lin(sd, p + ".attention.position_encoding", pos)
        q, k, v = (_lin(sd, f"{p}.attention.{nm}", n).view(B, W, num_heads, hd).permute(0, 2, 1, 3)
                   for nm in ("query", "key", "value"))
        s = q @ k.transpose(-1, -2) / math.sqrt(hd) + sd[p + ".attention.window_pos_bias"][:, :, :, :W]
        ctx = (F.softmax(s, dim=-1) @ v).permute(0, 2, 1, 3).contiguous().view(B, W, H)
        x = x + _lin(sd, p + ".attention.out_proj", ctx)
        n = _ln(sd, p + ".ffn_norm", x)
        for j in range(4):
            n = _lin(sd, f"{p}.ffn.{3 * j}", n)
            if j < 3:
                n = F.gelu(n)
        x = x + n
    x = _ln(sd, "window_attn.final_layer_norm", x)[:, mid]
    logits = _mlp(sd, "window_attn.classifier", x, 5, act=F.gelu)
    return logits, F.softmax(logits, dim=-1)


def window_forward(sd, img_clips, text_ids, attention_masks, T, head_type="cross_attn", shift_div=8):
    """two_stream_window.TwoStream.forward, :392-445."""
    B, W, L = text_ids.shape
    fused = []
    for i in range(W):
        lang_emb = orc.bert_forward(sd, text_ids[:, i], attention_masks[:, i])
        x = img_clips[:, i].reshape(B * T, *img_clips.shape[3:]).contiguous()
        vis = orc.resnet50_tsm_forward(sd, x, T, shift_div).view(B, T, -1)
        fused.append(chapter_head(sd, lang_emb, vis, i, T, head_type))
    return window_stack(sd, torch.stack(fused, dim=1))


def _center_attention(sd, p, x, q, k, v, bias_row, num_heads=16, pre_norm=None, post_norm=None):
    """The centre clip queries its window: x [B,W,H] -> context [B,H].  Positions (t - W//2) / (W//2 + 1e-6) go through
    position_encoding = Sequential(Linear(1,H), LayerNorm(H)) and are added to x."""
    B, W, H = x.shape
    hd, mid = H // num_heads, W // 2
    if pre_norm:
        x = _ln(sd, pre_norm, x)
    pos = ((torch.arange(W) - mid).float() / (mid + 1e-6)).unsqueeze(-1)
    x = x + _ln(sd, p + ".position_encoding.1", _lin(sd, p + ".position_encoding.0", pos))
    if post_norm:
        x = _ln(sd, post_norm, x)
    qh = _lin(sd, q, x[:, mid:mid + 1]).view(B, 1, num_heads, hd).transpose(1, 2)
    kh = _lin(sd, k, x).view(B, W, num_heads, hd).transpose(1, 2)
    vh = _lin(sd, v, x).view(B, W, num_heads, hd).transpose(1, 2)
    att = F.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(hd) + bias_row, dim=-1)
    return (att @ vh).transpose(1, 2).contiguous().view(B, H)


def domain_specific_head(sd, lang_embs, vision_embs, T, H=128):
    """two_stream_domain_specific.ChapterHead.forward (:318-369): lang_embs [B,W,768], vision_embs [B,W,T,2048]."""
    B, W = lang_embs.shape[:2]
    lang, vis = [], []
    for i in range(W):
        lang.append(F.relu(_mlp(sd, f"fusion_head.lang_proj_heads.{i}", lang_embs[:, i], 2)))
        v = F.relu(_mlp(sd, f"fusion_head.vision_proj_heads.{i}", vision_embs[:, i].reshape(-1, vision_embs.shape[-1]), 3))
        vis.append(v.view(B, T, H).mean(dim=1))
    centre = []
    for name, x in (("lang_window_attn", torch.stack(lang, 1)), ("vision_window_attn", torch.stack(vis, 1))):
        p = f"fusion_head.{name}"
        bias = sd[p + ".window_pos_bias"][:, :, W // 2:W // 2 + 1, :W]          # the centre query's row
        ctx = _center_attention(sd, p, x, p + ".query_proj", p + ".key_proj", p + ".value_proj", bias, post_norm=p + ".norm")
        centre.append(_mlp(sd, p + ".out_proj", ctx, 4))
    logits = _mlp(sd, "fusion_head.classifier", torch.cat(centre, dim=1), 5)
    return logits, F.softmax(logits, dim=1)


def domain_specific_forward(sd, img_clips, text_ids, attention_masks, T, shift_div=8):
    """two_stream_domain_specific.TwoStream.forward, :446-482."""
    B, W, L = text_ids.shape
    lang, vis = [], []
    for i in range(W):
        lang.append(orc.bert_forward(sd, text_ids[:, i], attention_masks[:, i]))
        x = img_clips[:, i].reshape(B * T, *img_clips.shape[3:]).contiguous()
        vis.append(orc.resnet50_tsm_forward(sd, x, T, shift_div).view(B, T, -1))
    return domain_specific_head(sd, torch.stack(lang, 1), torch.stack(vis, 1), T)


def single_block_classifier(sd, x):
    """window_self_attention.VideoChapterClassifier.forward (:201-206) with VideoChapterBlock (:156-170) and
    VideoChapterWindowAttention (:80-121): x [B,W,H] -> logits, probs."""
    W, mid = x.shape[1], x.shape[1] // 2
    a = "window_block.attention"
    ctx = _center_attention(sd, a, x, a + ".query", a + ".key", a + ".value", sd[a + ".window_pos_bias"][:, :, :, :W],
                            pre_norm="window_block.attention_norm")
    y = _lin(sd, a + ".out_proj", ctx) + x[:, mid]
    n = _ln(sd, "window_block.ffn_norm", y)
    y = y + _lin(sd, "window_block.ffn.4", F.gelu(_lin(sd, "window_block.ffn.1", n)))
    logits = _lin(sd, "classifier.4", F.gelu(_lin(sd, "classifier.1", _ln(sd, "classifier.0", y))))
    return logits, F.softmax(logits, dim=-1)
