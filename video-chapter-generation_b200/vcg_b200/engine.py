"""Engine: Python handle on a ``vcg_engine`` (include/vcg.h).

Takes a reference-schema state dict whose tensors live on the GPU, hands every entry to
``vcg_load_tensor`` and finalises (BatchNorm folding / repacking happen inside the library).  torch is used only to
own device memory and to name the current stream.
"""
import ctypes

import torch

from . import binding as _b

_HEADS = {"mlp": _b.HEAD_MLP, "attn": _b.HEAD_ATTN}
_PRECS = {"bf16": _b.PREC_BF16, "fp32": _b.PREC_FP32}
_MODALITIES = {"two_stream": _b.MODALITY_TWO_STREAM, "vision": _b.MODALITY_VISION, "text": _b.MODALITY_TEXT,
               "embed": _b.MODALITY_EMBED}


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Engine:
    def __init__(self, clip_frames, head_type="mlp", precision="bf16", vision=True, max_tokens=128, max_batch=32,
                 hidden_size=128, shift_div=8, device=None, modality="two_stream"):
        if head_type not in _HEADS:
            raise RuntimeError(f"Unknown head_type {head_type}")
        if precision not in _PRECS:
            raise RuntimeError(f"Unknown precision {precision}")
        if modality not in _MODALITIES:
            raise RuntimeError(f"Unknown modality {modality}")
        self.modality = modality
        if not torch.cuda.is_available():
            raise RuntimeError("vcg_b200 needs a CUDA (sm_100a) device: there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.clip_frames, self.head_type, self.precision = clip_frames, head_type, precision
        self.vision, self.max_tokens, self.max_batch = vision, max_tokens, max_batch
        self._lib = _b.load_library()
        cfg = _b.VcgConfig(clip_frames, max_tokens, hidden_size, _HEADS[head_type], _PRECS[precision],
                           _b.VISION_R50TSM if vision else _b.VISION_NONE, max_batch, shift_div, _MODALITIES[modality])
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _b.check(self._lib.vcg_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self._h = handle
        self._keep = []   # pinned staging tensors etc.
        self._frame_size = (224, 224)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vcg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict):
        """state_dict: reference key schema; CPU tensors are moved to the GPU one at a time."""
        with torch.cuda.device(self.device):
            for key, t in state_dict.items():
                if self.modality == "two_stream":
                    wanted = (key.startswith("lang_model.") or key.startswith("fusion_head.") or
                              (self.vision and key.startswith("vision_model.")))
                elif self.modality == "embed":   # backbones only (window model)
                    wanted = key.startswith("lang_model.") or key.startswith("vision_model.")
                else:   # Resnet50TSM / Resnet50 / BertHugface state dicts: base_model.* + head.*
                    wanted = key.startswith("base_model.") or key in ("head.weight", "head.bias")
                if not wanted:
                    continue
                if key.endswith("num_batches_tracked"):
                    continue
                t = t.detach()
                if t.dtype != torch.float32:
                    t = t.float()
                t = t.to(self.device, non_blocking=False).contiguous()
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                _b.check(self._lib.vcg_load_tensor(self._h, key.encode(), t.data_ptr(), shape, t.dim(),
                                                   _b.DTYPE_F32, _stream()))
                torch.cuda.current_stream().synchronize()   # `t` may be a temporary
            _b.check(self._lib.vcg_finalize(self._h, _stream()))

    # ------------------------------------------------------------------ scoring
    def _frames(self, frames_u8):
        """uint8 HWC frames [n,Hs,Ws,3]: frames of another size than 224 x 224 are resized (PIL BILINEAR, bit-identical)
        inside the pre-processing kernel (vcg_set_frame_size)."""
        assert frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous() and frames_u8.dim() == 4 and frames_u8.shape[3] == 3
        size = (int(frames_u8.shape[1]), int(frames_u8.shape[2]))
        if size != self._frame_size:
            with torch.cuda.device(self.device):
                _b.check(self._lib.vcg_set_frame_size(self._h, size[0], size[1], _stream()))
            self._frame_size = size

    def _text(self, text_ids, attention_mask):
        if not (text_ids.is_cuda and attention_mask.is_cuda):
            raise RuntimeError("vcg_b200: inputs must be CUDA tensors (no CPU fallback)")
        ids = text_ids.long().contiguous()
        mask = attention_mask.long().contiguous()
        B, L = ids.shape
        if L > self.max_tokens:
            raise RuntimeError(f"vcg_b200: {L} tokens exceed the engine's max_tokens={self.max_tokens}")
        return ids, mask, B, L

    def forward(self, img_clip, text_ids, attention_mask, return_emb=False, vision_emb=None):
        """TwoStream.forward: img_clip [B,T,3,224,224] fp32 (normalised) or vision_emb [B,T,2048] fp32."""
        ids, mask, B, L = self._text(text_ids, attention_mask)
        dev = ids.device
        logits = torch.empty(B, 2, dtype=torch.float32, device=dev)
        probs = torch.empty(B, 2, dtype=torch.float32, device=dev)
        ve = le = None
        img_ptr = emb_ptr = 0
        if vision_emb is not None:
            vision_emb = vision_emb.float().contiguous().view(B, self.clip_frames, 2048)
            emb_ptr = vision_emb.data_ptr()
        else:
            if not img_clip.is_cuda:
                raise RuntimeError("vcg_b200: inputs must be CUDA tensors (no CPU fallback)")
            img_clip = img_clip.float().contiguous()
            if tuple(img_clip.shape[1:]) != (self.clip_frames, 3, 224, 224):
                raise RuntimeError(f"vcg_b200: img_clip must be [B,{self.clip_frames},3,224,224], got {tuple(img_clip.shape)}")
            img_ptr = img_clip.data_ptr()
        if return_emb:
            ve = torch.empty(B, self.clip_frames, 2048, dtype=torch.float32, device=dev)
            le = torch.empty(B, 768, dtype=torch.float32, device=dev)
        if B == 0:      # nothing to score (empty tensors have no device pointer to hand over)
            return (logits, probs, ve, le) if return_emb else (logits, probs)
        with torch.cuda.device(dev):
            _b.check(self._lib.vcg_forward(self._h, img_ptr, emb_ptr, ids.data_ptr(), mask.data_ptr(), B, L,
                                           logits.data_ptr(), probs.data_ptr(), 0 if ve is None else ve.data_ptr(),
                                           0 if le is None else le.data_ptr(), _stream()))
        if return_emb:
            return logits, probs, ve, le
        return logits, probs

    def embed(self, img_clip, text_ids, attention_mask):
        """Backbone embeddings of B clips: (vision_emb [B,T,2048], lang_emb [B,768]) fp32 (modality "embed")."""
        ids, mask, B, L = self._text(text_ids, attention_mask)
        img_clip = img_clip.float().contiguous()
        if not img_clip.is_cuda or tuple(img_clip.shape) != (B, self.clip_frames, 3, 224, 224):
            raise RuntimeError(f"vcg_b200: img_clip must be a CUDA tensor [B,{self.clip_frames},3,224,224], got {tuple(img_clip.shape)}")
        dev = ids.device
        ve = torch.empty(B, self.clip_frames, 2048, dtype=torch.float32, device=dev)
        le = torch.empty(B, 768, dtype=torch.float32, device=dev)
        if B == 0:
            return ve, le
        with torch.cuda.device(dev):
            _b.check(self._lib.vcg_embed(self._h, img_clip.data_ptr(), ids.data_ptr(), mask.data_ptr(), B, L,
                                         ve.data_ptr(), le.data_ptr(), _stream()))
        return ve, le

    def embed_u8(self, frames_u8, text_ids, attention_mask, clip_start=None, first_start=0, clip_stride=4):
        """embed() from device-resident uint8 HWC frames [n,224,224,3]: clip b = frames clip_start[b].. (int32 CUDA), or
        the regular grid first_start + b*clip_stride when clip_start is None (stem once per distinct frame)."""
        ids, mask, B, L = self._text(text_ids, attention_mask)
        assert frames_u8.is_cuda
        self._frames(frames_u8)
        dev = ids.device
        if clip_start is not None:
            clip_start = clip_start.to(device=dev, dtype=torch.int32).contiguous()
            assert clip_start.numel() == B
        ve = torch.empty(B, self.clip_frames, 2048, dtype=torch.float32, device=dev)
        le = torch.empty(B, 768, dtype=torch.float32, device=dev)
        if B == 0:
            return ve, le
        with torch.cuda.device(dev):
            _b.check(self._lib.vcg_embed_u8(self._h, frames_u8.data_ptr(), frames_u8.shape[0],
                                            0 if clip_start is None else clip_start.data_ptr(), first_start, clip_stride,
                                            ids.data_ptr(), mask.data_ptr(), B, L, ve.data_ptr(), le.data_ptr(), _stream()))
        return ve, le

    def forward_vision(self, img_clip, return_emb=False):
        """Resnet50TSM.forward / Resnet50.forward: img_clip [B,T,3,224,224] fp32 -> (logits, probs)."""
        if not img_clip.is_cuda:
            raise RuntimeError("vcg_b200: inputs must be CUDA tensors (no CPU fallback)")
        img_clip = img_clip.float().contiguous()
        if tuple(img_clip.shape[1:]) != (self.clip_frames, 3, 224, 224):
            raise RuntimeError(f"vcg_b200: img_clip must be [B,{self.clip_frames},3,224,224], got {tuple(img_clip.shape)}")
        B, dev = img_clip.shape[0], img_clip.device
        logits = torch.empty(B, 2, dtype=torch.float32, device=dev)
        probs = torch.empty(B, 2, dtype=torch.float32, device=dev)
        ve = torch.empty(B, self.clip_frames, 2048, dtype=torch.float32, device=dev) if return_emb else None
        if B == 0:
            return (logits, probs, ve) if return_emb else (logits, probs)
        with torch.cuda.device(dev):
            _b.check(self._lib.vcg_forward_vision(self._h, img_clip.data_ptr(), B, logits.data_ptr(), probs.data_ptr(),
                                                  0 if ve is None else ve.data_ptr(), _stream()))
        return (logits, probs, ve) if return_emb else (logits, probs)

    def forward_text(self, text_ids, attention_mask, return_emb=False):
        """BertHugface.forward (pretrain_stage=False): pooler output -> Linear(768, 2) -> softmax."""
        ids, mask, B, L = self._text(text_ids, attention_mask)
        dev = ids.device
        logits = torch.empty(B, 2, dtype=torch.float32, device=dev)
        probs = torch.empty(B, 2, dtype=torch.float32, device=dev)
        le = torch.empty(B, 768, dtype=torch.float32, device=dev) if return_emb else None
        if B == 0:
            return (logits, probs, le) if return_emb else (logits, probs)
        with torch.cuda.device(dev):
            _b.check(self._lib.vcg_forward_text(self._h, ids.data_ptr(), mask.data_ptr(), B, L, logits.data_ptr(),
                                                probs.data_ptr(), 0 if le is None else le.data_ptr(), _stream()))
        return (logits, probs, le) if return_emb else (logits, probs)

    def score_clips_u8(self, frames_u8, clip_start, text_ids, attention_mask, out=None, clip_start_host=None):
        """Device-resident uint8 HWC frames [n,224,224,3] + int32 clip starts [B] -> (logits, probs).
        clip_start_host (the same starts as a CPU int32 tensor) lets the engine plan shared-stem vision passes over
        every regular run of overlapping clips (vcg_score_clips_u8_planned)."""
        ids, mask, B, L = self._text(text_ids, attention_mask)
        assert frames_u8.is_cuda
        self._frames(frames_u8)
        clip_start = clip_start.to(torch.int32).contiguous()
        assert clip_start.is_cuda and clip_start.numel() == B
        if clip_start_host is not None:
            assert (not clip_start_host.is_cuda and clip_start_host.dtype == torch.int32 and
                    clip_start_host.is_contiguous() and clip_start_host.numel() == B)
        dev = ids.device
        if out is None:
            logits = torch.empty(B, 2, dtype=torch.float32, device=dev)
            probs = torch.empty(B, 2, dtype=torch.float32, device=dev)
        else:
            logits, probs = out
        if B == 0:      # nothing to score (empty tensors have no device pointer to hand over)
            return logits, probs
        with torch.cuda.device(dev):
            if clip_start_host is not None:
                _b.check(self._lib.vcg_score_clips_u8_planned(self._h, frames_u8.data_ptr(), frames_u8.shape[0],
                                                              clip_start.data_ptr(), clip_start_host.data_ptr(),
                                                              ids.data_ptr(), mask.data_ptr(), B, L,
                                                              logits.data_ptr(), probs.data_ptr(), _stream()))
            else:
                _b.check(self._lib.vcg_score_clips_u8(self._h, frames_u8.data_ptr(), frames_u8.shape[0],
                                                      clip_start.data_ptr(), ids.data_ptr(), mask.data_ptr(), B, L,
                                                      logits.data_ptr(), probs.data_ptr(), _stream()))
        return logits, probs

    def score_video_u8(self, frames_u8, first_start, clip_stride, text_ids, attention_mask, out=None):
        """Clips on a regular grid (clip b = frames first_start + b*clip_stride ..): shares the stem between
        overlapping clips.  Device-resident uint8 HWC frames -> (logits, probs)."""
        ids, mask, B, L = self._text(text_ids, attention_mask)
        assert frames_u8.is_cuda
        self._frames(frames_u8)
        dev = ids.device
        if out is None:
            logits = torch.empty(B, 2, dtype=torch.float32, device=dev)
            probs = torch.empty(B, 2, dtype=torch.float32, device=dev)
        else:
            logits, probs = out
        if B == 0:
            return logits, probs
        with torch.cuda.device(dev):
            _b.check(self._lib.vcg_score_video_u8(self._h, frames_u8.data_ptr(), frames_u8.shape[0], first_start,
                                                  clip_stride, ids.data_ptr(), mask.data_ptr(), B, L,
                                                  logits.data_ptr(), probs.data_ptr(), _stream()))
        return logits, probs

    def score_clips_u8_host(self, frames_u8, clip_start, text_ids, attention_mask, out=None):
        """HOST tensors in (pinned for full speed), host tensors out; H2D/D2H copies happen inside the call."""
        for t in (frames_u8, clip_start, text_ids, attention_mask):
            assert not t.is_cuda and t.is_contiguous()
        assert clip_start.dtype == torch.int32
        self._frames(frames_u8)
        assert text_ids.dtype == torch.int64 and attention_mask.dtype == torch.int64
        B, L = text_ids.shape
        if out is None:
            logits = torch.empty(B, 2, dtype=torch.float32).pin_memory()
            probs = torch.empty(B, 2, dtype=torch.float32).pin_memory()
        else:
            logits, probs = out
        if B == 0:
            return logits, probs
        with torch.cuda.device(self.device):
            _b.check(self._lib.vcg_score_clips_u8_host(self._h, frames_u8.data_ptr(), frames_u8.shape[0],
                                                       clip_start.data_ptr(), text_ids.data_ptr(),
                                                       attention_mask.data_ptr(), B, L, logits.data_ptr(),
                                                       probs.data_ptr(), _stream()))
        return logits, probs

    def forward_host(self, vision_emb, text_ids, attention_mask, img_clip=None, out=None):
        """HOST tensors in, host tensors out (TwoStream.forward through vcg_forward_host)."""
        src = img_clip if img_clip is not None else vision_emb
        for t in (src, text_ids, attention_mask):
            assert not t.is_cuda and t.is_contiguous()
        assert src.dtype == torch.float32 and text_ids.dtype == torch.int64 and attention_mask.dtype == torch.int64
        B, L = text_ids.shape
        if out is None:
            logits = torch.empty(B, 2, dtype=torch.float32).pin_memory()
            probs = torch.empty(B, 2, dtype=torch.float32).pin_memory()
        else:
            logits, probs = out
        if B == 0:
            return logits, probs
        with torch.cuda.device(self.device):
            _b.check(self._lib.vcg_forward_host(self._h, 0 if img_clip is None else img_clip.data_ptr(),
                                                0 if img_clip is not None else vision_emb.data_ptr(),
                                                text_ids.data_ptr(), attention_mask.data_ptr(), B, L,
                                                logits.data_ptr(), probs.data_ptr(), _stream()))
        return logits, probs

    def profile_begin(self):
        _b.check(self._lib.vcg_profile_begin(self._h))

    def profile_end(self):
        """-> list of dicts {kernel, layer, launches, ms, flops, bytes}, one per '<kernel>|<layer>' name."""
        n = ctypes.c_int32(0)
        buf = (_b.VcgProfileEntry * 256)()
        with torch.cuda.device(self.device):
            _b.check(self._lib.vcg_profile_end(self._h, _stream(), buf, 256, ctypes.byref(n)))
        out = []
        for i in range(min(n.value, 256)):
            kernel, _, layer = buf[i].name.decode().partition("|")
            out.append({"kernel": kernel, "layer": layer, "launches": int(buf[i].launches), "ms": buf[i].ms,
                        "flops": buf[i].flops, "bytes": buf[i].bytes})
        return out

    def debug_checksums(self):
        """Checksums of the last vision pass's kernel outputs (engine created under VCG_DEBUG_CHECKSUM=1)."""
        buf = (ctypes.c_uint64 * 256)()
        n = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            _b.check(self._lib.vcg_debug_checksums(self._h, buf, 256, ctypes.byref(n), _stream()))
        return [int(buf[i]) for i in range(n.value)]

    @property
    def launch_count(self):
        return int(self._lib.vcg_launch_count(self._h))
