"""Shared plumbing of the window-model variants: nn.Sequential stretches -> vcg_op_mlp_chain programs (one CTA per row)
or, for wide Linear layers over many rows, the tcgen05 GEMM in its fp32 (3xTF32) mode; and the centre-query window
attention call.  Only pointers and sizes cross the C ABI."""
import ctypes

import torch
from torch import nn


class MulHalves(nn.Module):
    """Chain marker: a row of 2n values becomes row[:n] * row[n:] (VCG_MLP_MULHALVES)."""


class MeanGroups(nn.Module):
    """Chain marker: a row of g*width values becomes the mean of its g groups (VCG_MLP_MEANGROUPS)."""

    def __init__(self, width):
        super().__init__()
        self.width = width


class Save(nn.Module):
    """Chain marker: remember the row (VCG_MLP_SAVE) ..."""


class AddSaved(nn.Module):
    """... and add it back (VCG_MLP_ADDSAVED): a residual connection inside one program."""


def mlp_op(m):
    from vcg_b200 import binding as B
    if isinstance(m, nn.Linear):
        return B.VcgMlpOp(B.MLP_LINEAR, m.in_features, m.out_features, 0.0, m.weight.data_ptr(),
                          m.bias.data_ptr() if m.bias is not None else None)
    if isinstance(m, nn.LayerNorm):
        return B.VcgMlpOp(B.MLP_LAYERNORM, 0, 0, m.eps, m.weight.data_ptr(), m.bias.data_ptr())
    simple = {nn.ReLU: B.MLP_RELU, nn.GELU: B.MLP_GELU, MulHalves: B.MLP_MULHALVES, nn.Softmax: B.MLP_SOFTMAX,
              Save: B.MLP_SAVE, AddSaved: B.MLP_ADDSAVED}
    for cls, code in simple.items():
        if isinstance(m, cls):
            return B.VcgMlpOp(code, 0, 0, 0.0, None, None)
    if isinstance(m, MeanGroups):
        return B.VcgMlpOp(B.MLP_MEANGROUPS, 0, m.width, 0.0, None, None)
    raise RuntimeError(f"unsupported module in an MLP chain: {type(m).__name__}")


def _out_dim(dim, m):
    if isinstance(m, nn.Linear):
        return m.out_features
    if isinstance(m, MulHalves):
        return dim // 2
    if isinstance(m, MeanGroups):
        return m.width
    return dim


def run_chain(seq, final_relu, x0, x1=None):
    """Modules of ``seq`` (Dropout skipped) over the rows of x0 (| x1 concatenated) -> [rows, out] fp32."""
    from vcg_b200 import binding as B
    from vcg_b200 import ops
    lib = B.load_library()
    s = torch.cuda.current_stream().cuda_stream
    mods = [m for m in seq if not isinstance(m, nn.Dropout)] + ([nn.ReLU()] if final_relu else [])
    pending, cur, cur1 = [], x0, x1

    def flush():
        nonlocal pending, cur, cur1
        if not pending:
            return
        out_dim = cur.shape[1] + (0 if cur1 is None else cur1.shape[1])
        for m in pending:
            out_dim = _out_dim(out_dim, m)
        ops_arr = (B.VcgMlpOp * len(pending))(*[mlp_op(m) for m in pending])
        out = torch.empty(cur.shape[0], out_dim, dtype=torch.float32, device=cur.device)
        B.check(lib.vcg_op_mlp_chain(cur.data_ptr(), cur.shape[1], cur.stride(0), 0 if cur1 is None else cur1.data_ptr(),
                                     0 if cur1 is None else cur1.shape[1], 0 if cur1 is None else cur1.stride(0),
                                     cur.shape[0], ops_arr, len(pending), out.data_ptr(), out.stride(0), s))
        pending, cur, cur1 = [], out, None

    saving = False
    for m in mods:
        if isinstance(m, MulHalves):      # its own program, so that a wide Linear behind it can take the GEMM
            pending.append(m)
            flush()
            continue
        saving = (saving or isinstance(m, Save)) and not isinstance(m, AddSaved)
        big = (isinstance(m, nn.Linear) and cur1 is None and not saving and x0.shape[0] >= 128 and m.in_features >= 512
               and m.out_features >= 512 and m.in_features % 32 == 0 and m.out_features % 64 == 0)
        if big:
            flush()
            cur = ops.gemm(cur.contiguous(), m.weight.detach(), m.bias.detach(), None, B.ACT_NONE)
        else:
            pending.append(m)
    flush()
    return cur


def center_attention(x, num_heads, pos_linear, pos_norm, pos_bias, bias_offset, query, key, value, pre_norm=None,
                     post_norm=None, out_proj=None, add_residual=False):
    """x [B,W,128] fp32 CUDA -> [B,128]: vcg_op_center_attention (the centre clip queries its window)."""
    from vcg_b200 import binding as B
    lib = B.load_library()
    Bn, W, H = x.shape
    x = x.contiguous()

    def wb(m):
        return (None, None) if m is None else (m.weight.data_ptr(), m.bias.data_ptr())

    p = B.VcgCenterAttnParams(num_heads, pos_bias.stride(1), bias_offset, int(add_residual), *wb(pre_norm), *wb(post_norm),
                              *wb(pos_linear), *wb(pos_norm), pos_bias.data_ptr(), *wb(query), *wb(key), *wb(value),
                              *wb(out_proj))
    out = torch.empty(Bn, H, dtype=torch.float32, device=x.device)
    B.check(lib.vcg_op_center_attention(ctypes.byref(p), x.data_ptr(), Bn, W, out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
    return out
