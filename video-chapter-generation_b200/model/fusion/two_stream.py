"""Mirror of the reference's model/fusion/two_stream.py — the two-stream chapter-boundary point model.

Drop-in for video_chapter_generation/model/fusion/two_stream.py: same classes (``SelfAttention``, ``ChapterHead``,
``TwoStream``), constructor arguments, ``build_chapter_head(output_size, head_type)``, checkpoint keys and
``forward(img_clip, text_ids, attention_mask, return_emb=False)`` results (reference :99-124, :172-194).

The modules below only hold parameters.  ``TwoStream.forward`` hands them to libvcg_b200.so (hand-written sm_100a
kernels behind the C ABI of include/vcg.h) on first use and whenever they change, then runs the whole forward —
BERT text stream, ResNet-50-TSM vision stream, ChapterHead, softmax — inside the library on the current CUDA
stream.  There is no eager/CPU fallback: without a Blackwell GPU or without the built library ``forward`` raises.

BatchNorm uses its running statistics (standard ``.eval()``, the semantics of test_whole_pipeline_per_video.py:105)
unless ``bn_batch_stats`` is set (attribute, or env VCG_BN_BATCH_STATS=1): then every BatchNorm2d normalises with the
statistics of the B*T frames of the call, which is what caller #1 computes after nulling the running statistics
(test_video_segment_point.py:116-122) — slower, layer by layer (vcg_b200/bn_batch.py), with the vision
stream in the fp32 (3xTF32) arithmetic unless ``bn_batch_precision = "bf16"``.  Extra knobs (attributes, not constructor arguments, so the reference call sites stay unchanged):
``precision`` ("bf16" default, or "fp32" = 3xTF32 verification mode; env VCG_PRECISION), ``vision_chunk`` (clips per
internal pass, env VCG_VISION_CHUNK) and ``max_tokens``.
"""
import os

import torch
from torch import nn

from ops.basic_ops import Identity
from ops.temporal_shift import TemporalShift


class SelfAttention(nn.Module):
    """Parameter layout of the reference's attn head (reference :8-29): key / query / value / proj."""

    def __init__(self, n_embd, n_head, output_size, attn_pdrop=0.1, resid_pdrop=0.1):
        super().__init__()
        assert n_embd % n_head == 0
        self.n_head = n_head
        self.n_embd = n_embd
        self.key = nn.Linear(n_embd, n_embd)
        self.query = nn.Linear(n_embd, n_embd)
        self.value = nn.Linear(n_embd, n_embd)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.resid_drop = nn.Dropout(resid_pdrop)
        self.proj = nn.Linear(n_embd, output_size)

    def forward(self, x, layer_past=None):
        raise RuntimeError("the head runs inside libvcg_b200.so; call TwoStream.forward")


class ChapterHead(nn.Module):
    def __init__(self, lang_emb_size, vision_emb_size, segment_size, hidden_size, output_size, head_type="mlp"):
        super().__init__()
        self.lang_emb_size = lang_emb_size
        self.vision_emb_size = vision_emb_size
        self.segment_size = segment_size
        self.hidden_size = hidden_size
        self.head_type = head_type
        self.lang_proj_head = nn.Linear(lang_emb_size, hidden_size, bias=False)
        self.vision_proj_head = nn.Linear(vision_emb_size, hidden_size, bias=False)
        if head_type == "mlp":
            self.head = nn.Linear((segment_size + 1) * hidden_size, output_size, bias=True)
        elif head_type == "attn":
            self.head = SelfAttention(hidden_size, 4, output_size)
        else:
            raise RuntimeError(f"Unknown head_type {head_type}")

    def forward(self, lang_emb, vision_emb):
        raise RuntimeError("the head runs inside libvcg_b200.so; call TwoStream.forward")


class TwoStream(nn.Module):
    def __init__(self, lang_model, vision_model, lang_embed_size, vision_embed_size, segment_size, hidden_size):
        super().__init__()
        self.lang_model = lang_model
        self.vision_model = vision_model
        self.segment_size = segment_size
        self.lang_embed_size = lang_embed_size
        self.vision_embed_size = vision_embed_size
        self.hidden_size = hidden_size
        # engine knobs
        self.precision = os.environ.get("VCG_PRECISION", "bf16")
        self.vision_chunk = int(os.environ.get("VCG_VISION_CHUNK", "32"))
        self.max_tokens = 128
        self.bn_batch_stats = os.environ.get("VCG_BN_BATCH_STATS", "0") == "1"
        # the mode exists to reproduce caller #1's numbers: its vision stream defaults to the fp32 arithmetic (1e-4); in
        # bf16 the per-layer rounding is not damped by folded running statistics and the logits land at ~3e-2
        self.bn_batch_precision = os.environ.get("VCG_BN_BATCH_PRECISION", "fp32")
        self._engine = None
        self._engine_key = None
        self._bn_vision = None
        self._bn_vision_key = None

    def build_chapter_head(self, output_size, head_type="mlp"):
        """head_type: mlp or attn (reference :118-124)."""
        if output_size != 2:
            raise RuntimeError("the boundary scorer has exactly two output classes")
        self.fusion_head = ChapterHead(self.lang_embed_size, self.vision_embed_size, self.segment_size,
                                       self.hidden_size, output_size, head_type)

    def configure_optimizers(self, train_config):
        raise NotImplementedError("training is out of scope (SURVEY.md section 2); this package is inference-only")

    # ------------------------------------------------------------------ engine plumbing
    def _vision_kind(self):
        """(has_backbone, shift_div): Identity -> precomputed embeddings; TemporalShift wrappers -> TSM."""
        vm = self.vision_model
        if isinstance(vm, Identity) or isinstance(vm, nn.Identity):
            return False, 8
        for m in vm.modules():
            if isinstance(m, TemporalShift):
                if m.n_segment != self.segment_size:
                    raise RuntimeError("TemporalShift n_segment differs from TwoStream segment_size")
                return True, m.fold_div
        return True, 0

    def _weights_version(self):
        v = 0
        for t in list(self.parameters()) + list(self.buffers()):
            v += t._version + (t.data_ptr() % 1000003)
        return v

    def _get_engine(self, device, n_tokens, precision=None):
        from vcg_b200.engine import Engine
        has_backbone, shift_div = self._vision_kind()
        max_tokens = max(self.max_tokens, n_tokens)
        precision = self.precision if precision is None else precision
        key = (str(device), precision, self.vision_chunk, max_tokens, has_backbone, shift_div,
               self.fusion_head.head_type, self._weights_version())
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self.max_tokens = max_tokens
            eng = Engine(self.segment_size, self.fusion_head.head_type, precision, has_backbone, max_tokens,
                         self.vision_chunk, self.hidden_size, shift_div, device=device)
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_key = eng, key
        return self._engine

    @property
    def engine(self):
        return self._engine

    def get_engine(self, device=None, n_tokens=None):
        """The vcg_b200.Engine holding this module's weights (built on first use): the way to the uint8 / host-buffer
        entry points (score_clips_u8_host ...) without a warm-up forward."""
        device = next(self.parameters()).device if device is None else device
        return self._get_engine(device, self.max_tokens if n_tokens is None else n_tokens)

    def _batch_stat_vision(self, device, shift_div):
        """Layer-by-layer vision stream with batch-statistics BatchNorm (vcg_b200.bn_batch), rebuilt when weights change."""
        from vcg_b200.bn_batch import BatchStatVision
        key = (str(device), self.bn_batch_precision, shift_div, self._weights_version())
        if self._bn_vision is None or key != self._bn_vision_key:
            self._bn_vision = BatchStatVision(self.state_dict(), self.segment_size, shift_div, self.bn_batch_precision,
                                              device)
            self._bn_vision_key = key
        return self._bn_vision

    # ------------------------------------------------------------------ forward
    def forward(self, img_clip, text_ids, attention_mask, return_emb=False):
        """-> (binary_logits [B,2], binary_prob [B,2]) (+ vision_emb [B,T,2048], lang_emb [B,768])."""
        if not text_ids.is_cuda:
            raise RuntimeError("TwoStream.forward needs CUDA inputs: the B200 implementation has no CPU fallback")
        if self.training:
            raise RuntimeError("TwoStream is inference-only here: call .eval() first")
        batch_stat = self.bn_batch_stats and self._vision_kind()[0]
        # the batch-statistics mode runs the WHOLE forward (text stream and head too) in bn_batch_precision
        eng = self._get_engine(text_ids.device, text_ids.shape[1], self.bn_batch_precision if batch_stat else None)
        with torch.no_grad():
            if eng.vision and batch_stat:
                emb = self._batch_stat_vision(text_ids.device, self._vision_kind()[1]).embed(img_clip)
                return eng.forward(None, text_ids, attention_mask, return_emb=return_emb, vision_emb=emb)
            if eng.vision:
                return eng.forward(img_clip, text_ids, attention_mask, return_emb=return_emb)
            # precomputed vision embeddings: [B,T,2048,1,1] (what rearrange + Identity + view yields)
            B = text_ids.shape[0]
            emb = img_clip.reshape(B, self.segment_size, -1)
            return eng.forward(None, text_ids, attention_mask, return_emb=return_emb, vision_emb=emb)
