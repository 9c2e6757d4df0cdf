"""Mirror of the reference's model/fusion/stacked_window_self_attention.py: parameter holders with the reference's
module structure (hence its state-dict keys) for VideoChapterWindowAttention / VideoChapterBlock /
StackedVideoChapterAttention (:6-224).  The arithmetic runs in libvcg_b200.so (vcg_op_window_stack); see
two_stream_window.TwoStream.forward."""
import torch
import torch.nn as nn


def _no_forward(*a, **kw):
    raise RuntimeError("this module only holds parameters; the window model runs inside libvcg_b200.so "
                       "(call two_stream_window.TwoStream.forward)")


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
        self.attention_dropout = nn.Dropout(dropout)
        self.position_encoding = nn.Linear(1, hidden_size)
        self.window_pos_bias = nn.Parameter(torch.zeros(1, num_attention_heads, 1, 2 * window_size + 1))
        for lin in (self.query, self.key, self.value, self.out_proj, self.position_encoding):
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
        self.ffn = nn.Sequential(
            nn.Linear(hidden_size, hidden_size * 2), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(hidden_size * 2, hidden_size * 4), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(hidden_size * 4, hidden_size * 2), nn.GELU(), nn.Dropout(dropout),
            nn.Linear(hidden_size * 2, hidden_size), nn.Dropout(dropout))
        for m in self.ffn:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    forward = _no_forward


class StackedVideoChapterAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_layers = 6
        self.layers = nn.ModuleList([
            VideoChapterBlock(config.hidden_size, config.num_attention_heads, config.window_size,
                              config.attention_probs_dropout_prob) for _ in range(self.num_layers)])
        self.final_layer_norm = nn.LayerNorm(config.hidden_size)
        h = config.hidden_size
        self.classifier = nn.Sequential(
            nn.Linear(h, h), nn.LayerNorm(h), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(h, h), nn.LayerNorm(h), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(h, h // 2), nn.LayerNorm(h // 2), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(h // 2, h // 4), nn.LayerNorm(h // 4), nn.GELU(), nn.Dropout(0.1),
            nn.Linear(h // 4, 2))
        for m in self.classifier:
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    forward = _no_forward
