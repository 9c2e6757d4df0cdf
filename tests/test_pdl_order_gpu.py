"""Ordering of the programmatic-dependent-launch chain under load (csrc/launch.cuh).

With the stream busy, the kernels of a forward are queued back to back and their CTAs become resident long before their
predecessors finish; anything a kernel reads ahead of its griddepcontrol.wait then comes from the PREVIOUS pass (or from a
fresh engine's zeroed counters).  The tcgen05 attention kernel once read its item count that way (the compiler had hoisted
the __ldg above the wait): invisible with identical inputs on an idle GPU, wrong on a busy one.  These tests make the GPU
busy first and compare with the same forward on an idle GPU, bit for bit."""
import pytest
import torch

pytestmark = pytest.mark.gpu

T, L = 8, 32


def _engine():
    from vcg_b200 import synthetic as W
    from vcg_b200.engine import Engine
    sd = {k: v.cuda() for k, v in W.make_state_dict(T, "attn", seed=123).items()}
    eng = Engine(T, "attn", "bf16", vision=True, max_tokens=128, max_batch=32)
    eng.load_state_dict(sd)
    return eng


def _busy(a, n=60):
    x = a
    for _ in range(n):
        x = (x @ a) * 1e-2
    return x


def _batch(B, seed, lo, hi):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 30000, (B, L), generator=g)
    lens = torch.randint(lo, hi + 1, (B,), generator=g)
    mask = (torch.arange(L)[None, :] < lens[:, None]).long()
    ids[:, 0] = 101
    ids = ids * mask
    vis = torch.randn(B, T, 2048, generator=g)
    return ids.cuda(), mask.cuda(), vis.cuda()


def test_first_forward_of_a_fresh_engine_on_a_busy_gpu():
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    ids, mask, vis = _batch(2, 1, 10, 32)
    for _ in range(3):                       # later engines recycle the device memory of earlier ones
        eng = _engine()
        torch.cuda.synchronize()
        keep = _busy(a)
        first = [t.clone() for t in eng.forward(None, ids, mask, True, vis)]
        torch.cuda.synchronize()
        quiet = eng.forward(None, ids, mask, True, vis)
        for x, y in zip(first, quiet):
            assert torch.equal(x, y)
        eng.close()
        del keep


def test_consecutive_passes_with_different_length_mixes_on_a_busy_gpu():
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    eng = _engine()
    # few long clips (12 attention items each) then many short ones (3 each) and back: the item count changes every pass
    batches = [_batch(4, 2, 28, 32), _batch(32, 3, 4, 12), _batch(3, 4, 20, 32),
               _batch(29, 5, 2, 32), _batch(1, 6, 32, 32)]
    quiet = []
    for ids, mask, vis in batches:
        quiet.append([t.clone() for t in eng.forward(None, ids, mask, True, vis)])
        torch.cuda.synchronize()
    for rep in range(3):
        keep = _busy(a)
        outs = [[t.clone() for t in eng.forward(None, ids, mask, True, vis)] for ids, mask, vis in batches]
        torch.cuda.synchronize()
        for got, ref in zip(outs, quiet):
            for x, y in zip(got, ref):
                assert torch.equal(x, y)
        del keep
    eng.close()
