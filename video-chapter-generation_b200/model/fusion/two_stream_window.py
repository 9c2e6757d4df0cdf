"""Mirror of the reference's model/fusion/two_stream_window.py — the WINDOW ("update") chapter-boundary model that
test_video_segment_update.py:99-107 builds: every sample is a window of 2w+1 neighbouring clips; each clip goes through
BERT + ResNet-50-TSM and a per-position fusion head (ChapterHead), the 2w+1 fused vectors through six window-attention
blocks (StackedVideoChapterAttention) and the middle clip's vector through a classifier.

Same classes, constructor arguments and state-dict keys as the reference (:11-445); the modules only hold parameters.
``TwoStream.forward(img_clips, text_ids, attention_masks, clip_info)`` runs
  * the backbones of all B*(2w+1) clips in ONE pass of the tcgen05 engine (vcg_embed; the reference loops over the window
    positions, :404-434),
  * the per-position heads and the window stack through the fp32 operators of include/vcg.h (vcg_op_mlp_chain,
    vcg_op_cross_attention, vcg_op_window_stack).
Head types: "cross_attn" (the caller's default, test_video_segment_update.py:43), "mlp", and the reference's experimental
"bilinear" (one 3xTF32 tcgen05 GEMM over the [2h*h, T*h] view of nn.Bilinear's weight + vcg_op_bilinear_contract),
"multiplication" (VCG_MLP_MULHALVES step of the chain program) and "self_attn" (vcg_op_self_attention_first).
No CPU / eager fallback.
"""
import ctypes
import math
import os

import torch
from torch import nn

from model.fusion._chain import MulHalves, run_chain
from model.fusion.stacked_window_self_attention import StackedVideoChapterAttention, _no_forward
from ops.temporal_shift import TemporalShift


class CrossAttention(nn.Module):
    def __init__(self, hidden_size, num_heads, dropout=0.1):
        super().__init__()
        if hidden_size % num_heads != 0:
            raise ValueError(f"The hidden size {hidden_size} is not a multiple of the number of attention "
                             f"heads {num_heads}.")
        self.hidden_size, self.num_heads, self.head_dim = hidden_size, num_heads, hidden_size // num_heads
        self.query_proj = nn.Linear(hidden_size, hidden_size)
        self.key_proj = nn.Linear(hidden_size, hidden_size)
        self.value_proj = nn.Linear(hidden_size, hidden_size)
        self.out_proj = nn.Linear(hidden_size, hidden_size)
        self.lang_norm = nn.LayerNorm(hidden_size)
        self.vision_norm = nn.LayerNorm(hidden_size)
        self.attention_dropout = nn.Dropout(dropout)
        self.output_dropout = nn.Dropout(dropout)
        self.frame_pos_encoding = nn.Linear(1, hidden_size)
        scale = 1.0 / math.sqrt(self.head_dim)
        for proj in (self.query_proj, self.key_proj, self.value_proj, self.out_proj):
            nn.init.xavier_uniform_(proj.weight, gain=scale)
            nn.init.zeros_(proj.bias)
        nn.init.xavier_uniform_(self.frame_pos_encoding.weight)
        nn.init.zeros_(self.frame_pos_encoding.bias)

    forward = _no_forward


def _mlp3(d_in, d1, d2, d_out):
    return nn.Sequential(nn.Linear(d_in, d1), nn.LayerNorm(d1), nn.ReLU(), nn.Dropout(0.1),
                         nn.Linear(d1, d2), nn.LayerNorm(d2), nn.ReLU(), nn.Dropout(0.1), nn.Linear(d2, d_out))


class SelfAttention(nn.Module):
    """Parameter holder of the reference's SelfAttention (:91-131): key / query / value / proj."""

    def __init__(self, n_embd, n_head, output_size, attn_pdrop=0.1, resid_pdrop=0.1):
        super().__init__()
        assert n_embd % n_head == 0
        self.n_head, self.n_embd = n_head, n_embd
        self.key = nn.Linear(n_embd, n_embd)
        self.query = nn.Linear(n_embd, n_embd)
        self.value = nn.Linear(n_embd, n_embd)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.resid_drop = nn.Dropout(resid_pdrop)
        self.proj = nn.Linear(n_embd, output_size)

    forward = _no_forward


class ChapterHead(nn.Module):
    def __init__(self, lang_emb_size, vision_emb_size, segment_size, hidden_size, window_size, output_size,
                 head_type="mlp"):
        super().__init__()
        self.lang_emb_size, self.vision_emb_size = lang_emb_size, vision_emb_size
        self.segment_size, self.hidden_size = segment_size, hidden_size
        self.head_type, self.window_size = head_type, window_size
        self.num_clips = 2 * window_size + 1
        self.lang_proj_heads = nn.ModuleList([
            nn.Sequential(nn.Linear(lang_emb_size, lang_emb_size // 2), nn.LayerNorm(lang_emb_size // 2), nn.ReLU(),
                          nn.Dropout(0.1), nn.Linear(lang_emb_size // 2, hidden_size)) for _ in range(self.num_clips)])
        self.vision_proj_heads = nn.ModuleList([
            _mlp3(vision_emb_size, 8 * hidden_size, 4 * hidden_size, hidden_size) for _ in range(self.num_clips)])
        if head_type == "mlp":
            self.head = nn.ModuleList([
                _mlp3((segment_size + 1) * hidden_size, 8 * hidden_size, 4 * hidden_size, hidden_size)
                for _ in range(self.num_clips)])
        elif head_type == "cross_attn":
            self.head = CrossAttention(hidden_size, num_heads=16)
            self.output_proj = nn.Linear(hidden_size, output_size)
        elif head_type == "bilinear":
            h = hidden_size
            self.bilinear_layers = nn.ModuleList([nn.Bilinear(h, h * segment_size, 2 * h) for _ in range(self.num_clips)])
            self.head = nn.ModuleList([
                nn.Sequential(nn.LayerNorm(2 * h), nn.ReLU(), nn.Dropout(0.1), nn.Linear(2 * h, h), nn.LayerNorm(h),
                              nn.ReLU(), nn.Dropout(0.1), nn.Linear(h, h)) for _ in range(self.num_clips)])
        elif head_type == "multiplication":
            h = hidden_size
            self.lang_expand_layers = nn.ModuleList([
                nn.Sequential(nn.Linear(h, 8 * h), nn.LayerNorm(8 * h), nn.ReLU(), nn.Dropout(0.1),
                              nn.Linear(8 * h, h * segment_size), nn.LayerNorm(h * segment_size), nn.ReLU(),
                              nn.Dropout(0.1)) for _ in range(self.num_clips)])
            self.head = nn.ModuleList([_mlp3(h * segment_size, 8 * h, 4 * h, h) for _ in range(self.num_clips)])
        elif head_type == "self_attn":
            self.head = SelfAttention(hidden_size, 4, hidden_size)
        else:
            raise RuntimeError(f"Unknown head_type {head_type}")

    forward = _no_forward


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class TwoStream(nn.Module):
    def __init__(self, lang_model, vision_model, lang_embed_size, vision_embed_size, segment_size, hidden_size,
                 window_size):
        super().__init__()
        self.lang_model, self.vision_model = lang_model, vision_model
        self.segment_size, self.hidden_size, self.window_size = segment_size, hidden_size, window_size
        self.lang_embed_size, self.vision_embed_size = lang_embed_size, vision_embed_size
        if hidden_size != 128:
            raise RuntimeError("the window kernels are built for hidden_size 128 (every reference caller)")
        h = hidden_size
        self.window_mlp = nn.Sequential(        # present in the reference's state dict, unused by its forward (:436-441)
            nn.Linear(h * (2 * window_size + 1), h), nn.LayerNorm(h), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h, h // 2), nn.LayerNorm(h // 2), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h // 2, h // 4), nn.LayerNorm(h // 4), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h // 4, h // 8), nn.LayerNorm(h // 8), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h // 8, h // 16), nn.LayerNorm(h // 16), nn.ReLU(), nn.Dropout(0.1),
            nn.Linear(h // 16, 2))
        self.precision = os.environ.get("VCG_PRECISION", "bf16")
        self.vision_chunk = int(os.environ.get("VCG_VISION_CHUNK", "32"))
        self.max_tokens = 128
        self._engine = None
        self._engine_key = None

    def build_chapter_head(self, output_size, head_type="mlp"):
        self.fusion_head = ChapterHead(self.lang_embed_size, self.vision_embed_size, self.segment_size,
                                       self.hidden_size, self.window_size, output_size, head_type)
        cfg = type("Config", (), {"hidden_size": self.hidden_size, "num_attention_heads": 16,
                                  "attention_probs_dropout_prob": 0.1, "window_size": self.window_size})
        self.window_attn = StackedVideoChapterAttention(cfg)

    def configure_optimizers(self, train_config):
        raise NotImplementedError("training is out of scope (SURVEY.md section 2); this package is inference-only")

    # ------------------------------------------------------------------ plumbing
    def _weights_version(self):
        v = 0
        for t in list(self.parameters()) + list(self.buffers()):
            v += t._version + (t.data_ptr() % 1000003)
        return v

    def _get_engine(self, device, n_tokens):
        from vcg_b200.engine import Engine
        shift_div = 0
        for m in self.vision_model.modules():
            if isinstance(m, TemporalShift):
                shift_div = m.fold_div
                break
        max_tokens = max(self.max_tokens, n_tokens)
        key = (str(device), self.precision, self.vision_chunk, max_tokens, shift_div, self._weights_version())
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self.max_tokens = max_tokens
            eng = Engine(self.segment_size, "mlp", self.precision, True, max_tokens, self.vision_chunk,
                         self.hidden_size, shift_div, device=device, modality="embed")
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    def _run_chain(self, seq, final_relu, x0, x1=None):
        """nn.Sequential of Linear / LayerNorm / ReLU / GELU / Dropout over the rows of x0 (| x1).  Wide Linear layers
        over many rows (the 2048->1024->512 vision projections, B*T rows) go to the tcgen05 GEMM in its fp32 (3xTF32)
        mode; everything else runs as one vcg_op_mlp_chain program per stretch (model/fusion/_chain.py)."""
        return run_chain(seq, final_relu, x0, x1)

    # ------------------------------------------------------------------ forward
    def _chapter_head(self, i, lang_emb, vis_emb):
        """ChapterHead.forward for window position i (:248-289): lang_emb [bs,768], vis_emb [bs,T,2048] -> [bs,H]."""
        from vcg_b200 import binding as B
        lib = B.load_library()
        fh = self.fusion_head
        T, H = self.segment_size, self.hidden_size
        bs, dev = lang_emb.shape[0], lang_emb.device
        s = torch.cuda.current_stream().cuda_stream
        le = lang_emb.contiguous()
        ve = vis_emb.reshape(bs * T, self.vision_embed_size).contiguous()
        lang_out = self._run_chain(fh.lang_proj_heads[i], True, le)                 # relu(proj(lang)) [bs,H]
        vis_out = self._run_chain(fh.vision_proj_heads[i], True, ve)                # [bs*T,H]
        if fh.head_type == "mlp":      # cat([vision_out, lang_out]) -> head[i]
            f = self._run_chain(fh.head[i], False, vis_out.view(bs, T * H), lang_out)
        elif fh.head_type == "multiplication":     # head[i](vision_out * lang_expand[i](lang_out))   (:274-279)
            expanded = self._run_chain(fh.lang_expand_layers[i], False, lang_out)              # [bs, T*H]
            f = self._run_chain([MulHalves()] + list(fh.head[i]), False, vis_out.view(bs, T * H), expanded)
        elif fh.head_type == "bilinear":           # head[i](Bilinear(lang_out, vision_flat))          (:269-272)
            from vcg_b200 import ops
            bl = fh.bilinear_layers[i]
            w2 = bl.weight.detach().view(bl.out_features * bl.in1_features, bl.in2_features)   # [(o,i), j]
            y = ops.gemm(vis_out.view(bs, T * H), w2, None, None, B.ACT_NONE)                   # [bs, (o,i)]
            z = torch.empty(bs, bl.out_features, dtype=torch.float32, device=dev)
            B.check(lib.vcg_op_bilinear_contract(y.data_ptr(), lang_out.data_ptr(), bl.bias.data_ptr(), bs,
                                                 bl.in1_features, bl.out_features, z.data_ptr(), s))
            f = self._run_chain(fh.head[i], False, z)
        elif fh.head_type == "self_attn":          # SelfAttention over [frames..., lang], first token  (:281-283)
            sa = fh.head
            p = B.VcgSelfAttnParams(sa.n_head, *[t.data_ptr() for t in (
                sa.query.weight, sa.query.bias, sa.key.weight, sa.key.bias, sa.value.weight, sa.value.bias,
                sa.proj.weight, sa.proj.bias)])
            f = torch.empty(bs, H, dtype=torch.float32, device=dev)
            B.check(lib.vcg_op_self_attention_first(ctypes.byref(p), vis_out.data_ptr(), lang_out.data_ptr(), bs, T,
                                                    f.data_ptr(), s))
        else:                           # cross_attn: lang queries the T frame vectors
            ca = fh.head
            p = B.VcgCrossAttnParams(ca.num_heads, *[t.data_ptr() for t in (
                ca.lang_norm.weight, ca.lang_norm.bias, ca.vision_norm.weight, ca.vision_norm.bias,
                ca.frame_pos_encoding.weight, ca.frame_pos_encoding.bias,
                ca.query_proj.weight, ca.query_proj.bias, ca.key_proj.weight, ca.key_proj.bias,
                ca.value_proj.weight, ca.value_proj.bias, ca.out_proj.weight, ca.out_proj.bias)])
            f = torch.empty(bs, H, dtype=torch.float32, device=dev)
            B.check(lib.vcg_op_cross_attention(ctypes.byref(p), lang_out.data_ptr(), vis_out.data_ptr(), bs, T,
                                               f.data_ptr(), s))
        return f

    def _fuse_and_classify(self, vis_by_pos, lang_by_pos):
        """vis_by_pos[i] [bs,T,2048], lang_by_pos[i] [bs,768] (fp32 CUDA) for the 2w+1 window positions ->
        (logits, probs) [bs,2]: per-position ChapterHead, six window-attention blocks, classifier."""
        from vcg_b200 import binding as B
        lib = B.load_library()
        W, H = 2 * self.window_size + 1, self.hidden_size
        bs, dev = lang_by_pos[0].shape[0], lang_by_pos[0].device
        s = torch.cuda.current_stream().cuda_stream
        fused = torch.empty(bs, W, H, dtype=torch.float32, device=dev)
        for i in range(W):
            fused[:, i] = self._chapter_head(i, lang_by_pos[i], vis_by_pos[i])
        # six window-attention blocks + classifier on the middle clip
        wa = self.window_attn
        sp = B.VcgWindowStackParams()
        sp.num_layers, sp.pos_bias_stride = wa.num_layers, W
        for li, blk in enumerate(wa.layers):
            a, lins = blk.attention, [m for m in blk.ffn if isinstance(m, nn.Linear)]
            vals = [blk.attention_norm.weight, blk.attention_norm.bias, blk.ffn_norm.weight, blk.ffn_norm.bias,
                    a.position_encoding.weight, a.position_encoding.bias, a.window_pos_bias,
                    a.query.weight, a.query.bias, a.key.weight, a.key.bias, a.value.weight, a.value.bias,
                    a.out_proj.weight, a.out_proj.bias]
            for lin in lins:
                vals += [lin.weight, lin.bias]
            sp.layers[li] = B.VcgWindowLayer(*[t.data_ptr() for t in vals])
        sp.final_norm_w, sp.final_norm_b = wa.final_layer_norm.weight.data_ptr(), wa.final_layer_norm.bias.data_ptr()
        lins = [m for m in wa.classifier if isinstance(m, nn.Linear)]
        lns = [m for m in wa.classifier if isinstance(m, nn.LayerNorm)]
        for j, lin in enumerate(lins):
            sp.cls_w[j], sp.cls_b[j] = lin.weight.data_ptr(), lin.bias.data_ptr()
        for j, ln in enumerate(lns):
            sp.cls_norm_w[j], sp.cls_norm_b[j] = ln.weight.data_ptr(), ln.bias.data_ptr()
        logits = torch.empty(bs, 2, dtype=torch.float32, device=dev)
        probs = torch.empty(bs, 2, dtype=torch.float32, device=dev)
        B.check(lib.vcg_op_window_stack(ctypes.byref(sp), fused.data_ptr(), bs, W, logits.data_ptr(), probs.data_ptr(), s))
        return logits, probs

    def _check(self, text_ids):
        if not text_ids.is_cuda:
            raise RuntimeError("TwoStream.forward needs CUDA inputs: the B200 implementation has no CPU fallback")
        if self.training:
            raise RuntimeError("inference-only: call .eval() first")

    def forward(self, img_clips, text_ids, attention_masks, clip_info=None):
        """img_clips [B,W,T,3,224,224], text_ids / attention_masks [B,W,L] -> (binary_logits, binary_prob) [B,2]
        (reference :392-445; clip_info is accepted and, as in the reference, not used by the arithmetic)."""
        self._check(text_ids)
        bs, W, L = text_ids.shape
        T = self.segment_size
        if W != 2 * self.window_size + 1:
            raise RuntimeError(f"expected {2 * self.window_size + 1} clips per window, got {W}")
        eng = self._get_engine(text_ids.device, L)
        with torch.no_grad():
            # backbones of every clip of every window, window position major: row = i * bs + b
            img = img_clips.float().transpose(0, 1).reshape(W * bs, T, 3, 224, 224)
            ids = text_ids.transpose(0, 1).reshape(W * bs, L)
            mask = attention_masks.transpose(0, 1).reshape(W * bs, L)
            vis_emb, lang_emb = eng.embed(img, ids, mask)                      # [W*bs,T,2048], [W*bs,768]
            return self._fuse_and_classify([vis_emb[i * bs:(i + 1) * bs] for i in range(W)],
                                           [lang_emb[i * bs:(i + 1) * bs] for i in range(W)])

    def score_video(self, img_clips, text_ids, attention_masks, skip=None):
        """Every candidate clip of ONE video as a window target, with each clip's backbone embedding computed once.

        img_clips [N,T,3,224,224], text_ids / attention_masks [N,L]: the video's clips in order.  Target n sees the
        clips n + (i - w) * skip, i = 0..2w (skip = clip_frame_num // (2 * max_offset) = T // 4, the dataset's
        get_clip_info, infer_youtube_video_dataset.py:459-478); positions outside the video are the dataset's padding
        clip (zero frames, zero ids, zero mask, :488-499).  Equals forward() on the materialised windows; the reference
        (and forward()) run the backbones 2w+1 times per clip, this runs them once.  -> (logits, probs) [N,2]."""
        self._check(text_ids)
        N, L = text_ids.shape
        T, w = self.segment_size, self.window_size
        skip = max(1, T // 4) if skip is None else skip
        eng = self._get_engine(text_ids.device, L)
        with torch.no_grad():
            dev = text_ids.device
            img = torch.cat([img_clips.float(), torch.zeros(1, T, 3, 224, 224, device=dev)])       # + padding clip
            ids = torch.cat([text_ids.long(), torch.zeros(1, L, dtype=torch.long, device=dev)])
            mask = torch.cat([attention_masks.long(), torch.zeros(1, L, dtype=torch.long, device=dev)])
            vis_emb, lang_emb = eng.embed(img, ids, mask)                      # [N+1,T,2048], [N+1,768]
            return self._score_from_embeddings(vis_emb, lang_emb, N, skip)

    def score_video_u8(self, frames_u8, text_ids, attention_masks, first_start=0, clip_stride=4, clip_start=None,
                       skip=None):
        """score_video() straight from the video's decoded uint8 HWC frames [n,224,224,3] (CUDA): clip n = frames
        first_start + n*clip_stride .. +T-1 (or clip_start[n]); ToTensor + Normalize run on the device and, on the regular
        grid, the ResNet stem once per distinct frame.  The dataset's padding clip (all-zero NORMALISED frames, which no
        uint8 frame can express) is embedded separately through the fp32 entry point.  -> (logits, probs) [N,2]."""
        self._check(text_ids)
        N, L = text_ids.shape
        T = self.segment_size
        skip = max(1, T // 4) if skip is None else skip
        eng = self._get_engine(text_ids.device, L)
        with torch.no_grad():
            dev = text_ids.device
            vis, lang = eng.embed_u8(frames_u8, text_ids.long(), attention_masks.long(), clip_start, first_start, clip_stride)
            pad_vis, pad_lang = eng.embed(torch.zeros(1, T, 3, 224, 224, device=dev),
                                          torch.zeros(1, L, dtype=torch.long, device=dev),
                                          torch.zeros(1, L, dtype=torch.long, device=dev))
            return self._score_from_embeddings(torch.cat([vis, pad_vis]), torch.cat([lang, pad_lang]), N, skip)

    def _score_from_embeddings(self, vis_emb, lang_emb, N, skip):
        """Row N of the embeddings is the padding clip; target n sees the rows n + (i - w) * skip."""
        w = self.window_size
        n = torch.arange(N, device=lang_emb.device)
        vis_by_pos, lang_by_pos = [], []
        for i in range(2 * w + 1):
            src = n + (i - w) * skip
            src = torch.where((src >= 0) & (src < N), src, torch.full_like(src, N))     # N = the padding clip
            vis_by_pos.append(vis_emb[src])
            lang_by_pos.append(lang_emb[src])
        return self._fuse_and_classify(vis_by_pos, lang_by_pos)
