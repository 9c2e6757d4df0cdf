"""Per-operator determinism stress: the same launch repeated, outputs compared bit for bit with the first run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
import torch
from vcg_b200 import ops, binding as B
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev, NF, T = "cuda", 256, 16
g = torch.Generator(device="cpu").manual_seed(1)
def rnd(*s, scale=1.0): return (torch.randn(*s, generator=g) * scale).to(dev).to(torch.bfloat16)
def conv_case(H, Cin, Cout, k, stride, res, tsm_out, tsm_in, act=B.ACT_RELU):
    x = rnd(NF, H, H, Cin); w = rnd(Cout, k, k, Cin, scale=(k * k * Cin) ** -0.5); b = torch.randn(Cout, generator=g).to(dev)
    Ho = H // stride
    r = rnd(NF, Ho, Ho, Cout) if res else None
    fold = Cout // 8
    ti = rnd(NF, H, H, Cin // 4) if tsm_in else None
    def run():
        to = torch.zeros(NF, Ho, Ho, 2 * fold, device=dev, dtype=torch.bfloat16) if tsm_out else None
        out = ops.conv2d_nhwc(x, w, b, r, stride, act, tsm_in=ti, tsm_out=to, tsm_fold=fold if tsm_out else 0, clip_frames=T)
        return (out, to) if tsm_out else (out,)
    return run
def tail_case(P, H, stride, variant):
    x = rnd(NF, H, H, P); w2 = rnd(P, 3, 3, P, scale=(9 * P) ** -0.5); w3 = rnd(4 * P, P, scale=P ** -0.5)
    b2 = torch.randn(P, generator=g).to(dev) * 0.1; b3 = torch.randn(4 * P, generator=g).to(dev) * 0.1
    r = rnd(NF, H // stride, H // stride, 4 * P)
    return lambda: (ops.bottleneck_tail(x, w2, b2, w3, b3, r, stride, variant=variant),)
def stem_case():
    fr = torch.randint(0, 256, (NF, 224, 224, 3), generator=g, dtype=torch.uint8).to(dev)
    w = rnd(64, 4, 8, 2, 4, scale=0.1); b = torch.randn(64, generator=g).to(dev)
    def run():
        xp = ops.preprocess_u8(fr)
        y = ops.stem_conv(xp, w, b)
        p, s = ops.maxpool_tsm(y, T)
        return (xp, y, p, s)
    return run
CASES = {
    "stem+maxpool": stem_case(),
    "conv1_l1 (tsm_in)": conv_case(56, 256, 64, 1, 1, False, False, True),
    "conv1_l1.0": conv_case(56, 64, 64, 1, 1, False, False, False),
    "ds_l1": conv_case(56, 64, 256, 1, 1, False, False, False, act=B.ACT_NONE),
    "conv2_l1": conv_case(56, 64, 64, 3, 1, False, False, False),
    "conv3_l1 (res, tsm_out)": conv_case(56, 64, 256, 1, 1, True, True, False),
    "conv1_l2 (tsm_in)": conv_case(28, 512, 128, 1, 1, False, False, False),
    "conv2_l2": conv_case(28, 128, 128, 3, 1, False, False, False),
    "conv3_l2 (res)": conv_case(28, 128, 512, 1, 1, True, False, False),
    "ds_l2 (s2)": conv_case(56, 256, 512, 1, 2, False, False, False, act=B.ACT_NONE),
    "conv2_l3 (pair)": conv_case(14, 256, 256, 3, 1, False, False, False),
    "conv3_l3 (res)": conv_case(14, 256, 1024, 1, 1, True, False, False),
    "conv1_l3 (pair)": conv_case(14, 1024, 256, 1, 1, False, False, False),
    "conv2_l4 (pair)": conv_case(7, 512, 512, 3, 1, False, False, False),
    "conv3_l4 (res)": conv_case(7, 512, 2048, 1, 1, True, False, False),
    "tail conv23h P64": tail_case(64, 56, 1, 1),
    "tail conv23t P128": tail_case(128, 28, 1, 0),
    "tail conv23t P128 s2": tail_case(128, 56, 2, 0),
}
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
for name, run in CASES.items():
    if only and not any(o in name for o in only): continue
    ref = [t.clone() for t in run()]
    torch.cuda.synchronize()
    bad = 0
    for rep in range(reps):
        out = run()
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(ref, out)):
            if not torch.equal(a, b):
                bad += 1
                d = (a.float() - b.float()).abs()
                nz = torch.nonzero(d.reshape(d.shape[0], -1).max(1).values > 0).flatten().tolist()
                print(f"  {name}: rep {rep} output {i} differs: images {nz[:12]} max {float(d.max()):.3e} n_elems {int((d > 0).sum())}", flush=True)
                break
    print(f"{name}: {bad} of {reps} runs differ", flush=True)
