"""Fused bottleneck tail (conv2 3x3 + conv3 1x1 + residual, csrc/conv23.cuh and csrc/conv23h.cuh) against a plain PyTorch
fp32 reference of the same op on the same bf16-rounded inputs: both kernel variants, ragged frame counts (tiles past the
end of the frame stack), the half-empty last tile row of the 8 x 16 halo tiling, and the TSM scatter
(ops/temporal_shift.py:34-51: channels [0,f) of frame t go to frame t-1, [f,2f) to frame t+1)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(x, w2, b2, w3, b3, res, stride):
    xf = x.float().permute(0, 3, 1, 2)
    mid = torch.relu(torch.nn.functional.conv2d(xf, w2.float().permute(0, 3, 1, 2), b2, stride=stride, padding=1))
    mid = mid.to(torch.bfloat16).float()                      # the kernels hand conv2's output to conv3 in bf16
    out = torch.nn.functional.conv2d(mid, w3.float()[:, :, None, None], b3)
    if res is not None:
        out = out + res.float().permute(0, 3, 1, 2)
    return torch.relu(out).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("variant,P,H,stride,n", [(0, 64, 56, 1, 6), (1, 64, 56, 1, 6), (1, 64, 56, 1, 1), (1, 64, 56, 1, 37),
                                                   (1, 64, 24, 1, 5), (0, 128, 28, 1, 8), (0, 128, 56, 2, 4), (1, 128, 28, 1, 9), (1, 128, 28, 1, 1),
                                                   (1, 128, 40, 1, 3),
                                                   # P = 128 without a TSM scatter (n % 4 != 0 below): conv23t_kernel, the
                                                   # conv3 operand in tensor memory; stride 1 and 2, ragged frame counts
                                                   (0, 128, 28, 1, 9), (0, 128, 56, 2, 5), (0, 128, 28, 1, 1), (0, 128, 28, 1, 37)])
def test_bottleneck_tail_matches_torch(variant, P, H, stride, n):
    from vcg_b200 import ops
    g = torch.Generator().manual_seed(7 + n + P)
    dev = "cuda"
    x = torch.randn(n, H, H, P, generator=g).to(dev).to(torch.bfloat16)
    w2 = (torch.randn(P, 3, 3, P, generator=g) / (9 * P) ** 0.5).to(dev).to(torch.bfloat16)
    w3 = (torch.randn(4 * P, P, generator=g) / P ** 0.5).to(dev).to(torch.bfloat16)
    b2 = torch.randn(P, generator=g).to(dev) * 0.1
    b3 = torch.randn(4 * P, generator=g).to(dev) * 0.1
    Ho = H // stride
    res = torch.randn(n, Ho, Ho, 4 * P, generator=g).to(dev).to(torch.bfloat16)
    T = 4 if (n % 4 == 0 and not (variant == 1 and P == 128)) else 1   # the P = 128 halo kernel has no TSM scatter
    fold = 4 * P // 8
    tsm = torch.zeros(n, Ho, Ho, 2 * fold, device=dev, dtype=torch.bfloat16) if T > 1 else None
    out = ops.bottleneck_tail(x, w2, b2, w3, b3, res, stride, tsm_out=tsm, tsm_fold=fold if T > 1 else 0, clip_frames=T,
                              variant=variant)
    torch.cuda.synchronize()
    want = _reference(x, w2, b2, w3, b3, res, stride)
    err = float((out.float() - want).abs().max() / want.abs().max())
    assert err <= 1e-2, err                                   # bf16 output rounding (2^-9) of O(1) values
    if tsm is not None:
        o = out.view(n // T, T, Ho, Ho, 4 * P)
        want_tsm = torch.zeros(n // T, T, Ho, Ho, 2 * fold, device=dev, dtype=torch.bfloat16)
        want_tsm[:, :-1, :, :, :fold] = o[:, 1:, :, :, :fold]
        want_tsm[:, 1:, :, :, fold:] = o[:, :-1, :, :, fold:2 * fold]
        assert torch.equal(tsm.view_as(want_tsm), want_tsm)


def test_halo_variant_equals_per_tap_variant():
    """Same inputs through conv23_kernel and conv23h_kernel: the same fp32 tensor-core sums; the halo variant adds the
    residual inside the accumulator (before the bias) instead of after it, so the fp32 values may differ in the last
    bit and the bf16 results by at most one bf16 ulp (2^-8 relative), and only rarely."""
    from vcg_b200 import ops
    g = torch.Generator().manual_seed(11)
    dev = "cuda"
    n, H, P = 16, 56, 64
    x = torch.randn(n, H, H, P, generator=g).to(dev).to(torch.bfloat16)
    w2 = (torch.randn(P, 3, 3, P, generator=g) / 24).to(dev).to(torch.bfloat16)
    w3 = (torch.randn(4 * P, P, generator=g) / 8).to(dev).to(torch.bfloat16)
    b2 = torch.randn(P, generator=g).to(dev) * 0.1
    b3 = torch.randn(4 * P, generator=g).to(dev) * 0.1
    res = torch.randn(n, H, H, 4 * P, generator=g).to(dev).to(torch.bfloat16)
    a = ops.bottleneck_tail(x, w2, b2, w3, b3, res, 1, variant=0)
    b = ops.bottleneck_tail(x, w2, b2, w3, b3, res, 1, variant=1)
    torch.cuda.synchronize()
    d = (a.float() - b.float()).abs()
    assert float((d / a.float().abs().clamp_min(1e-3)).max()) <= 2 ** -7
    assert float((d > 0).float().mean()) < 0.02
