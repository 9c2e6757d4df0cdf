"""GPU parity of the batch-statistics BatchNorm mode (reference caller #1, test_video_segment_point.py:116-122):
the stand-alone operators against torch fp32 restatements, and the whole two-stream forward through the mirror's
TwoStream (bn_batch_stats = True) against the goldens the UNMODIFIED reference produced with its BatchNorm2d layers
treated as caller #1 treats them (oracle/make_golden_bn_batch.py).  Tolerances as everywhere: 1e-4 fp32, 2e-2 bf16."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from test_bn_batch import golden_case, rel
from test_parity_gpu import TOL, build_model

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,rows", [(64, 12544 * 3 + 0), (256, 3137), (512, 785), (2048, 49 * 5)])
def test_bn_stats_and_apply(dtype, C, rows):
    from vcg_b200 import ops
    torch.manual_seed(C + rows)
    x = (torch.randn(rows, C, device="cuda") * (1 + torch.rand(C, device="cuda")) + torch.randn(C, device="cuda")).to(dtype)
    mean, rstd = ops.bn_batch_stats(x, 1e-5)
    xd = x.double()
    m_ref = xd.mean(0)
    v_ref = xd.var(0, unbiased=False)
    assert float((mean.double() - m_ref).abs().max()) <= 1e-6 * max(1.0, float(m_ref.abs().max()))
    assert rel(rstd, 1.0 / torch.sqrt(v_ref + 1e-5)) <= 1e-6
    m2, r2 = ops.bn_batch_stats(x, 1e-5)
    assert torch.equal(mean, m2) and torch.equal(rstd, r2)          # fixed reduction order
    gamma, beta = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
    res = torch.randn(rows, C, device="cuda").to(dtype)
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    for residual, relu in ((None, True), (res, True), (None, False), (res, False)):
        y = ops.bn_apply(x, mean, rstd, gamma, beta, residual, relu)
        ref = (x.float() - mean) * rstd * gamma + beta
        if residual is not None:
            ref = ref + residual.float()
        if relu:
            ref = F.relu(ref)
        assert y.dtype == dtype and rel(y, ref) <= tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T,shift_div,C,hw", [(8, 8, 64, 6), (16, 8, 256, 5), (4, 4, 512, 3), (1, 8, 64, 4)])
def test_tsm_shift_and_avgpool(dtype, T, shift_div, C, hw):
    from vcg_b200 import ops
    n = 2 * T
    x = torch.randn(n, hw, hw, C, device="cuda").to(dtype)
    fold = C // shift_div
    out = ops.tsm_shift(x, T, fold)
    v = x.view(2, T, hw, hw, C)
    ref = torch.zeros_like(v)
    ref[:, :-1, ..., :fold] = v[:, 1:, ..., :fold]
    ref[:, 1:, ..., fold:2 * fold] = v[:, :-1, ..., fold:2 * fold]
    ref[..., 2 * fold:] = v[..., 2 * fold:]
    assert torch.equal(out, ref.view(n, hw, hw, C))
    pooled = ops.avgpool(x)
    assert pooled.dtype == torch.float32 and rel(pooled, x.float().mean(dim=(1, 2))) <= 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_raw_stem_conv(dtype):
    from vcg_b200 import ops
    from vcg_b200.bn_batch import pack_stem_weight
    torch.manual_seed(5)
    img = torch.randn(3, 3, 224, 224, device="cuda")
    w = torch.randn(64, 3, 7, 7, device="cuda") / 147 ** 0.5
    y = ops.stem_conv_act(ops.nchw_to_stem(img, dtype), pack_stem_weight(w, dtype))
    ref = F.conv2d(img.to(dtype).float(), w.to(dtype).float(), stride=2, padding=3).permute(0, 2, 3, 1)
    assert float(y.float().min()) < 0                                 # no ReLU, no bias
    assert rel(y, ref) <= (2e-5 if dtype == torch.float32 else 1e-2)


# (engine precision = text stream + head, precision of the batch-statistics vision stream, logits tol, vision_emb tol).
# The vision stream of this mode defaults to the fp32 arithmetic; bf16 is an opt-in whose per-layer rounding is not damped
# by folded running statistics: a CPU simulation of bf16 rounding after every conv and BatchNorm gives 4.2e-2 / 4.6e-2 on
# vision_emb and 1.8e-2 / 2.7e-2 on the logits of the two goldens, hence the looser bound for that combination.
MODES = [("fp32", "fp32", 1e-4, 1e-4), ("bf16", "fp32", 2e-2, 1e-4), ("bf16", "bf16", 5e-2, 6e-2)]


@pytest.mark.parametrize("precision,vision_precision,tol,vis_tol", MODES)
@pytest.mark.parametrize("name,head", [("attn_T8_L32_B2", "attn"), ("mlp_T16_L100_B3", "mlp")])
def test_two_stream_batch_stat_mode_matches_reference(precision, vision_precision, tol, vis_tol, name, head):
    g, T, L, B, ids, mask, img = golden_case(name)
    model, _ = build_model(T, head, precision)
    assert model.bn_batch_precision == "fp32"                          # the default of the mode
    model.bn_batch_stats, model.bn_batch_precision = True, vision_precision
    logits, probs, vis, lang = model(img.cuda(), ids.cuda(), mask.cuda(), return_emb=True)
    assert rel(vis, g["vision_emb"]) <= vis_tol
    assert rel(logits, g["logits"]) <= tol and rel(probs, g["probs"]) <= tol
    assert logits.topk(1, 1, True, True)[1].view(-1).tolist() == g["labels"].tolist()
    # the clips of a call are coupled, exactly as in the reference: clip 0 alone gives the reference's "alone" logits
    alone = model(img[:1].cuda(), ids[:1].cuda(), mask[:1].cuda())[0]
    assert rel(alone, g["logits_clip0_alone"]) <= tol
    # and the default (standard eval, running statistics) is a different function
    model.bn_batch_stats = False
    assert rel(model(img.cuda(), ids.cuda(), mask.cuda())[0], g["logits"]) > 0.5
