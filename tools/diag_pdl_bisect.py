"""Bisects the first-forward-on-a-busy-GPU problem to a kernel boundary: fresh bf16 engine, GPU busy with matmuls, first
forward with launches [from, to) of the text chain sent without the PDL attribute; compares lang_emb with a quiet rerun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200")); sys.path.insert(0, ROOT)
import torch
from vcg_b200 import synthetic as W
from vcg_b200 import binding as _b
from vcg_b200.engine import Engine
T, L, B = 8, 32, 2
sd = {k: v.cuda() for k, v in W.make_state_dict(T, "attn", seed=123).items()}
ids, mask = W.make_text(B, L, seed=123)
ids, mask = ids.cuda(), mask.cuda()
vis = torch.randn(B, T, 2048, device="cuda")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
lib = _b.load_library()
def busy(n=60):
    x = a
    for _ in range(n):
        x = (x @ a) * 1e-2
    return x
def trial(frm, to):
    eng = Engine(T, "attn", "bf16", vision=True, max_tokens=128, max_batch=32)
    eng.load_state_dict(sd)
    torch.cuda.synchronize()
    keep = busy()
    lib.vcg_debug_pdl_window(frm, to)
    first = eng.forward(None, ids, mask, True, vis)[3].clone()
    n = lib.vcg_debug_pdl_window(0, 0)
    torch.cuda.synchronize()
    second = eng.forward(None, ids, mask, True, vis)[3]
    bad = not torch.equal(first, second)
    eng.close()
    return bad
for i in range(2):   # engines to recycle memory from
    trial(0, 0)
print("no window:", "BAD" if trial(0, 0) else "good", flush=True)
print("all off [0,200):", "BAD" if trial(0, 200) else "good", flush=True)
windows = [tuple(int(x) for x in w.split(":")) for w in sys.argv[1:]]
for frm, to in windows:
    print(f"PDL off for launches [{frm},{to}):", "BAD" if trial(frm, to) else "good", flush=True)
