"""Mirror of the reference's model/fusion/window_self_attention.py — the single-block predecessor of the stacked window
attention: the centre clip queries its window once (VideoChapterWindowAttention :7-121), residual + FFN
(VideoChapterBlock :124-170), LayerNorm + 2-layer classifier (VideoChapterClassifier :173-206).

Same classes / constructor arguments / state-dict keys.  VideoChapterClassifier.forward(fusion_emb [B,W,128], clip_info)
-> (logits, probs) runs as one vcg_op_center_attention launch and one vcg_op_mlp_chain program (+ softmax)."""
import torch
import torch.nn as nn

from model.fusion._chain import AddSaved, Save, center_attention, run_chain
from model.fusion.stacked_window_self_attention import _no_forward


class VideoChapterWindowAttention(nn.Module):
    def __init__(self, hidden_size, num_attention_heads, window_size, dropout=0.1):
        super().__init__()
        if hidden_size % num_attention_heads != 0:
            raise ValueError(f"The hidden size {hidden_size} is not a multiple of the number of attention "
                             f"heads {num_attention_heads}.")
        self.num_attention_heads = num_attention_heads
        self.attention_head_size = hidden_size // num_attention_heads
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.window_size = window_size
        self.query = nn.Linear(hidden_size, self.all_head_size)
        self.key = nn.Linear(hidden_size, self.all_head_size)
        self.value = nn.Linear(hidden_size, self.all_head_size)
        self.out_proj = nn.Linear(hidden_size, hidden_size)
        self.attention_dropout = nn.Dropout(0.2)
        self.position_encoding = nn.Sequential(nn.Linear(1, hidden_size), nn.LayerNorm(hidden_size), nn.Dropout(dropout))
        self.window_pos_bias = nn.Parameter(torch.zeros(1, num_attention_heads, 1, 2 * window_size + 1))
        for lin in (self.query, self.key, self.value, self.out_proj, self.position_encoding[0]):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)
        nn.init.normal_(self.window_pos_bias, mean=0.0, std=0.02)

    forward = _no_forward


class VideoChapterBlock(nn.Module):
    def __init__(self, hidden_size, num_attention_heads, window_size, dropout=0.1):
        super().__init__()
        self.attention_norm = nn.LayerNorm(hidden_size)
        self.ffn_norm = nn.LayerNorm(hidden_size)
        self.attention = VideoChapterWindowAttention(hidden_size, num_attention_heads, window_size, dropout)
        self.ffn = nn.Sequential(nn.Dropout(0.1), nn.Linear(hidden_size, hidden_size * 4), nn.GELU(), nn.Dropout(0.25),
                                 nn.Linear(hidden_size * 4, hidden_size), nn.Dropout(0.15))
        for m in self.ffn:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    forward = _no_forward


class VideoChapterClassifier(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.window_block = VideoChapterBlock(config.hidden_size, config.num_attention_heads, config.window_size,
                                              config.attention_probs_dropout_prob)
        h = config.hidden_size
        self.classifier = nn.Sequential(nn.LayerNorm(h), nn.Linear(h, h // 2), nn.GELU(), nn.Dropout(0.1),
                                        nn.Linear(h // 2, 2))
        for m in self.classifier:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, fusion_emb, clip_info=None):
        """fusion_emb [B,W,128] fp32 CUDA -> (logits, probs) [B,2]."""
        if not fusion_emb.is_cuda:
            raise RuntimeError("VideoChapterClassifier.forward needs CUDA inputs: there is no CPU fallback")
        if self.training:
            raise RuntimeError("inference-only: call .eval() first")
        blk, att = self.window_block, self.window_block.attention
        with torch.no_grad():
            # attention(attention_norm(x)) + x[:, centre]
            y = center_attention(fusion_emb.float(), att.num_attention_heads, att.position_encoding[0],
                                 att.position_encoding[1], att.window_pos_bias, 0, att.query, att.key, att.value,
                                 pre_norm=blk.attention_norm, out_proj=att.out_proj, add_residual=True)
            # ffn(ffn_norm(y)) + y, then the classifier
            logits = run_chain([Save(), blk.ffn_norm] + list(blk.ffn) + [AddSaved()] + list(self.classifier), False, y)
            probs = run_chain([nn.Softmax(dim=-1)], False, logits)
        return logits, probs
