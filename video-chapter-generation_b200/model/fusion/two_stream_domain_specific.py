"""Mirror of the reference's model/fusion/two_stream_domain_specific.py — the third fusion variant: per-position
projection heads as in the window model, the T frame vectors mean-pooled, ONE window self-attention per modality
(WindowSelfAttention :9-135) of which only the centre clip's row is used (:354-356), concat, 5-layer classifier
(ChapterHead.forward :318-369; the CrossAttention member exists in the state dict but its call is commented out, :358-359).

Same classes, constructor arguments and state-dict keys; the modules hold parameters.  TwoStream.forward runs the
backbones of all B*(2w+1) clips in one engine pass (as two_stream_window does) and the head through vcg_op_mlp_chain /
vcg_op_center_attention.  No CPU / eager fallback.
"""
import math

import torch
from torch import nn

from model.fusion import two_stream_window as _tsw
from model.fusion._chain import MeanGroups, center_attention, run_chain
from model.fusion.stacked_window_self_attention import _no_forward


def _out_mlp(h):
    return nn.Sequential(nn.Linear(h, 2 * h), nn.LayerNorm(2 * h), nn.ReLU(), nn.Dropout(0.1),
                         nn.Linear(2 * h, 2 * h), nn.LayerNorm(2 * h), nn.ReLU(), nn.Dropout(0.1),
                         nn.Linear(2 * h, 2 * h), nn.LayerNorm(2 * h), nn.ReLU(), nn.Dropout(0.1),
                         nn.Linear(2 * h, h))


def _init_attention(mod, scale):
    for proj in (mod.query_proj, mod.key_proj, mod.value_proj):
        nn.init.xavier_uniform_(proj.weight, gain=scale)
        nn.init.zeros_(proj.bias)
    for m in mod.out_proj.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight, gain=scale)
            nn.init.zeros_(m.bias)


class WindowSelfAttention(nn.Module):
    def __init__(self, hidden_size, num_heads, window_size, dropout=0.1):
        super().__init__()
        if hidden_size % num_heads != 0:
            raise ValueError(f"The hidden size {hidden_size} is not a multiple of the number of attention "
                             f"heads {num_heads}.")
        self.hidden_size, self.num_heads, self.head_dim = hidden_size, num_heads, hidden_size // num_heads
        self.window_size = window_size
        self.query_proj = nn.Linear(hidden_size, hidden_size)
        self.key_proj = nn.Linear(hidden_size, hidden_size)
        self.value_proj = nn.Linear(hidden_size, hidden_size)
        self.out_proj = _out_mlp(hidden_size)
        self.norm = nn.LayerNorm(hidden_size)
        self.attention_dropout = nn.Dropout(dropout)
        self.output_dropout = nn.Dropout(dropout)
        self.window_pos_bias = nn.Parameter(torch.zeros(1, num_heads, 2 * window_size + 1, 2 * window_size + 1))
        self.position_encoding = nn.Sequential(nn.Linear(1, hidden_size), nn.LayerNorm(hidden_size), nn.Dropout(dropout))
        _init_attention(self, 1.0 / math.sqrt(self.head_dim))
        nn.init.normal_(self.window_pos_bias, mean=0.0, std=0.02)
        nn.init.xavier_uniform_(self.position_encoding[0].weight)
        nn.init.zeros_(self.position_encoding[0].bias)

    forward = _no_forward

    def center_row(self, x):
        """x [B,W,H] -> the centre clip's row of forward(x) [B,H] (the only row ChapterHead uses)."""
        W = x.shape[1]
        if W > self.window_pos_bias.shape[-1]:
            raise RuntimeError("window longer than window_pos_bias")
        ctx = center_attention(x, self.num_heads, self.position_encoding[0], self.position_encoding[1],
                               self.window_pos_bias, (W // 2) * self.window_pos_bias.shape[-1], self.query_proj,
                               self.key_proj, self.value_proj, post_norm=self.norm)
        return run_chain(self.out_proj, False, ctx)


class CrossAttention(nn.Module):
    """Present in the reference's ChapterHead (and its state dict); its forward is never called (:358-359)."""

    def __init__(self, hidden_size, num_heads, dropout=0.1):
        super().__init__()
        if hidden_size % num_heads != 0:
            raise ValueError(f"The hidden size {hidden_size} is not a multiple of the number of attention "
                             f"heads {num_heads}.")
        self.hidden_size, self.num_heads, self.head_dim = hidden_size, num_heads, hidden_size // num_heads
        self.query_proj = nn.Linear(hidden_size, hidden_size)
        self.key_proj = nn.Linear(hidden_size, hidden_size)
        self.value_proj = nn.Linear(hidden_size, hidden_size)
        self.out_proj = _out_mlp(hidden_size)
        self.vision_norm = nn.LayerNorm(hidden_size)
        self.lang_norm = nn.LayerNorm(hidden_size)
        self.attention_dropout = nn.Dropout(dropout)
        self.output_dropout = nn.Dropout(dropout)
        _init_attention(self, 1.0 / math.sqrt(self.head_dim))

    forward = _no_forward


class ChapterHead(nn.Module):
    def __init__(self, lang_emb_size, vision_emb_size, segment_size, hidden_size, window_size, output_size):
        super().__init__()
        self.lang_emb_size, self.vision_emb_size = lang_emb_size, vision_emb_size
        self.segment_size, self.hidden_size, self.window_size = segment_size, hidden_size, window_size
        self.num_clips = 2 * window_size + 1
        h = hidden_size
        self.lang_proj_heads = nn.ModuleList([
            nn.Sequential(nn.Linear(lang_emb_size, lang_emb_size // 2), nn.LayerNorm(lang_emb_size // 2), nn.ReLU(),
                          nn.Dropout(0.1), nn.Linear(lang_emb_size // 2, h)) for _ in range(self.num_clips)])
        self.vision_proj_heads = nn.ModuleList([_tsw._mlp3(vision_emb_size, 8 * h, 4 * h, h) for _ in range(self.num_clips)])
        self.lang_window_attn = WindowSelfAttention(h, 16, window_size, 0.1)
        self.vision_window_attn = WindowSelfAttention(h, 16, window_size, 0.1)
        self.cross_attn = CrossAttention(h, 16, 0.1)
        self.classifier = nn.Sequential(
            nn.Linear(2 * h, 2 * h), nn.LayerNorm(2 * h), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(2 * h, h), nn.LayerNorm(h), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h, h // 2), nn.LayerNorm(h // 2), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h // 2, h // 4), nn.LayerNorm(h // 4), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h // 4, output_size))

    forward = _no_forward


class TwoStream(_tsw.TwoStream):
    def __init__(self, lang_model, vision_model, lang_embed_size, vision_embed_size, segment_size, hidden_size,
                 window_size):
        super().__init__(lang_model, vision_model, lang_embed_size, vision_embed_size, segment_size, hidden_size,
                         window_size)
        del self.window_mlp          # not part of this variant's state dict

    def build_chapter_head(self, output_size, head_type=None):
        """head_type is accepted and ignored, as in the reference (:386-398)."""
        self.fusion_head = ChapterHead(self.lang_embed_size, self.vision_embed_size, self.segment_size,
                                       self.hidden_size, self.window_size, output_size)

    def _fuse_and_classify(self, vis_by_pos, lang_by_pos):
        fh = self.fusion_head
        W, T, H = 2 * self.window_size + 1, self.segment_size, self.hidden_size
        bs, dev = lang_by_pos[0].shape[0], lang_by_pos[0].device
        lang_t = torch.empty(bs, W, H, dtype=torch.float32, device=dev)
        vis_t = torch.empty(bs, W, H, dtype=torch.float32, device=dev)
        for i in range(W):
            lang_t[:, i] = run_chain(fh.lang_proj_heads[i], True, lang_by_pos[i].contiguous())
            frames = run_chain(fh.vision_proj_heads[i], True,
                               vis_by_pos[i].reshape(bs * T, self.vision_embed_size).contiguous())     # [bs*T,H]
            vis_t[:, i] = run_chain([MeanGroups(H)], False, frames.view(bs, T * H))                     # mean over frames
        lang_c = fh.lang_window_attn.center_row(lang_t)
        vis_c = fh.vision_window_attn.center_row(vis_t)
        logits = run_chain(fh.classifier, False, lang_c, vis_c)                                        # cat -> classifier
        probs = run_chain([nn.Softmax(dim=1)], False, logits)
        return logits, probs
