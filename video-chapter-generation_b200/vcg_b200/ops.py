"""Torch-tensor front ends of the stand-alone C-ABI operators (vcg_op_*).

torch is used for device memory and the current stream only; all arithmetic happens in libvcg_b200.so.
Activation tensors are NHWC (channels-last, contiguous), bf16 (precision 0) or fp32 (precision 1).
"""
import torch

from . import binding as _b

STEM_HP, STEM_WP = 230, 240


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _prec(t):
    if t.dtype == torch.bfloat16:
        return _b.PREC_BF16
    if t.dtype == torch.float32:
        return _b.PREC_FP32
    raise TypeError(f"unsupported activation dtype {t.dtype}")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vcg_b200 operators need CUDA tensors: there is no CPU fallback")


def gemm(a, w, bias=None, residual=None, act=_b.ACT_NONE, out=None):
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias (+ residual)).  a may be a strided row view (stride(1) == 1)."""
    _need_cuda(a, w, bias, residual)
    M, K = a.shape
    N = w.shape[0]
    assert a.stride(1) == 1 and w.is_contiguous() and w.shape[1] == K and w.dtype == a.dtype
    if out is None:
        out = torch.empty(M, N, dtype=a.dtype, device=a.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_gemm(a.data_ptr(), a.stride(0), w.data_ptr(), _ptr(bias), _ptr(residual),
                             0 if residual is None else residual.stride(0), out.data_ptr(), out.stride(0), M, N, K,
                             act, _prec(a), _stream()))
    return out


def conv2d_nhwc(x, w, bias=None, residual=None, stride=1, act=_b.ACT_NONE, tsm_in=None, tsm_out=None, tsm_fold=0,
                clip_frames=1):
    """x [n,H,W,Cin], w [Cout,k,k,Cin] -> [n,H/stride,W/stride,Cout]."""
    _need_cuda(x, w, bias, residual, tsm_in, tsm_out)
    n, H, W, Cin = x.shape
    Cout, k = w.shape[0], w.shape[1]
    assert x.is_contiguous() and w.is_contiguous() and w.dtype == x.dtype
    out = torch.empty(n, H // stride, W // stride, Cout, dtype=x.dtype, device=x.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_conv2d_nhwc(x.data_ptr(), n, H, W, Cin, w.data_ptr(), _ptr(bias), _ptr(residual),
                                    out.data_ptr(), Cout, k, stride, act, _prec(x), _ptr(tsm_in),
                                    0 if tsm_in is None else tsm_in.shape[-1], _ptr(tsm_out), tsm_fold, clip_frames,
                                    _stream()))
    return out


def bottleneck_tail(x, w2, bias2, w3, bias3, residual=None, stride=1, tsm_out=None, tsm_fold=0, clip_frames=1, variant=0):
    """Fused conv2 (3x3) + conv3 (1x1, residual, ReLU) of a bottleneck, bf16: x [n,H,W,P], w2 [P,3,3,P], w3 [4P,P]
    -> [n,H/stride,W/stride,4P].  variant 0 = conv23_kernel, 1 = conv23h_kernel (P = 64, stride 1)."""
    _need_cuda(x, w2, bias2, w3, bias3, residual, tsm_out)
    n, H, W, P = x.shape
    assert x.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16 and w3.dtype == torch.bfloat16
    assert x.is_contiguous() and w2.is_contiguous() and w3.is_contiguous()
    out = torch.empty(n, H // stride, W // stride, 4 * P, dtype=x.dtype, device=x.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_bottleneck_tail(x.data_ptr(), n, H, W, P, stride, w2.data_ptr(), bias2.data_ptr(), w3.data_ptr(),
                                        bias3.data_ptr(), _ptr(residual), out.data_ptr(), _ptr(tsm_out), tsm_fold,
                                        clip_frames, variant, _stream()))
    return out


def preprocess_u8(frames, frame_index=None, dtype=torch.bfloat16):
    """uint8 [n,224,224,3] -> normalised, zero-padded stem input (fp32: NHWC4 [n,230,240,4]; bf16: the same bytes
    count with row pairs interleaved per pixel, i.e. [n,115,240,2,4] viewed as [n,230,240,4])."""
    _need_cuda(frames, frame_index)
    n = frames.shape[0] if frame_index is None else frame_index.shape[0]
    out = torch.zeros(n, STEM_HP, STEM_WP, 4, dtype=dtype, device=frames.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_preprocess_u8(frames.data_ptr(), _ptr(frame_index), n, out.data_ptr(), _prec(out), _stream()))
    return out


def resize_u8(frames, want_u8=True, stem_dtype=None):
    """uint8 [n,Hs,Ws,3] -> bilinear resize to 224 x 224 (PIL / torchvision BILINEAR, bit-identical): returns
    (uint8 [n,224,224,3] or None, normalised padded stem input in ``stem_dtype`` or None)."""
    _need_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.dim() == 4 and frames.shape[3] == 3
    n, Hs, Ws = frames.shape[:3]
    out_u8 = torch.empty(n, 224, 224, 3, dtype=torch.uint8, device=frames.device) if want_u8 else None
    out_st = torch.zeros(n, STEM_HP, STEM_WP, 4, dtype=stem_dtype, device=frames.device) if stem_dtype is not None else None
    lib = _b.load_library()
    _b.check(lib.vcg_op_resize_u8(frames.data_ptr(), n, Hs, Ws, _ptr(out_u8), _ptr(out_st),
                                  _b.PREC_FP32 if stem_dtype == torch.float32 else _b.PREC_BF16, _stream()))
    return out_u8, out_st


def nchw_to_stem(img, dtype=torch.bfloat16):
    """fp32 [n,3,224,224] (normalised) -> zero-padded NHWC4 [n,230,240,4]."""
    _need_cuda(img)
    n = img.shape[0]
    out = torch.zeros(n, STEM_HP, STEM_WP, 4, dtype=dtype, device=img.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_nchw_to_stem(img.data_ptr(), n, out.data_ptr(), _prec(out), _stream()))
    return out


def stem_conv(x_padded, w_packed, bias):
    """padded NHWC4 -> relu(conv7x7/2 + bias) as NHWC [n,112,112,64]; w_packed bf16 [64,4,8,2,4] / fp32 [64,7,8,4]."""
    _need_cuda(x_padded, w_packed, bias)
    n = x_padded.shape[0]
    out = torch.empty(n, 112, 112, 64, dtype=x_padded.dtype, device=x_padded.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_stem_conv(x_padded.data_ptr(), n, w_packed.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                  _prec(x_padded), _stream()))
    return out


def maxpool_tsm(x, clip_frames, shift_div=8):
    """[n,112,112,64] -> (pooled [n,56,56,64], temporally shifted copy [n,56,56,64]; None when shift_div == 0)."""
    _need_cuda(x)
    n = x.shape[0]
    out = torch.empty(n, 56, 56, 64, dtype=x.dtype, device=x.device)
    shifted = torch.zeros_like(out) if shift_div > 0 else None
    lib = _b.load_library()
    _b.check(lib.vcg_op_maxpool_tsm(x.data_ptr(), n, out.data_ptr(), _ptr(shifted), clip_frames, shift_div,
                                    _prec(x), _stream()))
    return out, shifted


def stem_conv_act(x_padded, w_packed, bias=None, act=_b.ACT_NONE):
    """stem_conv with an explicit activation and an optional bias (batch-statistics BatchNorm mode wants the raw conv)."""
    _need_cuda(x_padded, w_packed, bias)
    n = x_padded.shape[0]
    out = torch.empty(n, 112, 112, 64, dtype=x_padded.dtype, device=x_padded.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_stem_conv_act(x_padded.data_ptr(), n, w_packed.data_ptr(), _ptr(bias), out.data_ptr(), act,
                                      _prec(x_padded), _stream()))
    return out


def bn_batch_stats(x, eps=1e-5):
    """Per-channel (mean, 1/sqrt(biased var + eps)) over every row of the NHWC activation x [..., C]: the statistics
    F.batch_norm uses when a BatchNorm2d has no running statistics (test_video_segment_point.py:116-122)."""
    _need_cuda(x)
    assert x.is_contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    lib = _b.load_library()
    n_partial = lib.vcg_op_bn_partials(rows, C)
    if n_partial <= 0:
        raise RuntimeError(f"vcg_b200: batch statistics need C = 64 * 2^k <= 2048 and at least one row (C={C}, rows={rows})")
    partial = torch.empty(n_partial, 2, C, dtype=torch.float64, device=x.device)
    mean = torch.empty(C, dtype=torch.float32, device=x.device)
    rstd = torch.empty(C, dtype=torch.float32, device=x.device)
    _b.check(lib.vcg_op_bn_batch_stats(x.data_ptr(), rows, C, eps, partial.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                       _prec(x), _stream()))
    return mean, rstd


def bn_apply(x, mean, rstd, gamma, beta, residual=None, relu=True):
    """relu?((x - mean) * rstd * gamma + beta (+ residual)) over the NHWC activation x [..., C]."""
    _need_cuda(x, mean, rstd, gamma, beta, residual)
    assert x.is_contiguous() and (residual is None or (residual.is_contiguous() and residual.shape == x.shape
                                                       and residual.dtype == x.dtype))
    C = x.shape[-1]
    out = torch.empty_like(x)
    lib = _b.load_library()
    _b.check(lib.vcg_op_bn_apply(x.data_ptr(), x.numel() // C, C, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                 beta.data_ptr(), _ptr(residual), 1 if relu else 0, out.data_ptr(), _prec(x), _stream()))
    return out


def tsm_shift(x, clip_frames, fold):
    """TemporalShift.shift (ops/temporal_shift.py:34-51) of the NHWC activation x [n,H,W,C], n = clips * clip_frames."""
    _need_cuda(x)
    assert x.is_contiguous() and x.dim() == 4
    n, H, W, C = x.shape
    out = torch.empty_like(x)
    lib = _b.load_library()
    _b.check(lib.vcg_op_tsm_shift(x.data_ptr(), n, H * W, C, clip_frames, fold, out.data_ptr(), _prec(x), _stream()))
    return out


def avgpool(x):
    """AdaptiveAvgPool2d(1): NHWC [n,H,W,C] -> fp32 [n,C]."""
    _need_cuda(x)
    assert x.is_contiguous() and x.dim() == 4
    n, H, W, C = x.shape
    out = torch.empty(n, C, dtype=torch.float32, device=x.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_avgpool(x.data_ptr(), n, H * W, C, out.data_ptr(), _prec(x), _stream()))
    return out


def bert_attention(qkv, attention_mask, B, L):
    """qkv [B*L, 2304], attention_mask int64 [B,L] -> ctx [B*L, 768]."""
    _need_cuda(qkv, attention_mask)
    ctx = torch.empty(B * L, 768, dtype=qkv.dtype, device=qkv.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_bert_attention(qkv.data_ptr(), attention_mask.data_ptr(), ctx.data_ptr(), B, L, _prec(qkv),
                                       _stream()))
    return ctx


def bert_attention_packed(qkv, cu, key_ok, max_len):
    """Token-packed layout: qkv [rows, 2304] bf16, cu int32 [B+1], key_ok uint8 [rows] -> ctx [rows, 768] (rows past
    cu[B] are left untouched)."""
    _need_cuda(qkv, cu, key_ok)
    assert qkv.dtype == torch.bfloat16 and cu.dtype == torch.int32 and key_ok.dtype == torch.uint8
    ctx = torch.zeros(qkv.shape[0], 768, dtype=qkv.dtype, device=qkv.device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_bert_attention_packed(qkv.data_ptr(), cu.data_ptr(), key_ok.data_ptr(), ctx.data_ptr(),
                                              cu.shape[0] - 1, max_len, qkv.shape[0], _stream()))
    return ctx


def layernorm(x, gamma, beta, eps=1e-12):
    _need_cuda(x, gamma, beta)
    y = torch.empty_like(x)
    lib = _b.load_library()
    _b.check(lib.vcg_op_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), x.shape[0],
                                  x.shape[1], eps, _prec(x), _stream()))
    return y
