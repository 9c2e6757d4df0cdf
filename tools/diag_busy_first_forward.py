"""Fresh bf16 engines whose FIRST forward is queued behind other GPU work (torch matmuls, or the operator launches of the
batch-statistics BatchNorm mode): lang_emb / logits against the reference golden.  This is the flow that exposed the attention
kernel reading its item count ahead of the dependent-launch wait; run with VCG_PDL=0 for the contrast."""
import sys, torch
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/video-chapter-generation_b200')
from test_bn_batch import golden_case, rel
from test_parity_gpu import build_model
from vcg_b200.bn_batch import BatchStatVision
g,T,L,B,ids,mask,img = golden_case("attn_T8_L32_B2")
ids, mask, img = ids.cuda(), mask.cuda(), img.cuda()
gv = torch.tensor(g["vision_emb"]).cuda()
def rep(tag, o):
    l = o[3]
    print(f"{tag}: lang {rel(l, g['lang_emb']):.2e} logits {rel(o[0], g['logits']):.2e}", flush=True)
def fresh():
    model,_ = build_model(T, "attn", "bf16")
    return model
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
def busy(n=60):
    x = a
    for _ in range(n):
        x = (x @ a) * 1e-2
    return x
for i in range(2):
    m = fresh(); eng = m.get_engine(ids.device, L); rep(f"warm-up engine {i} (idle GPU)", eng.forward(None, ids, mask, True, gv))
busy(5); torch.cuda.synchronize()
m = fresh(); eng = m.get_engine(ids.device, L); keep = busy()
rep("E1 fresh engine, GPU busy with torch matmuls, first forward", eng.forward(None, ids, mask, True, gv))
rep("E1b second forward (idle)", eng.forward(None, ids, mask, True, gv))
m = fresh(); eng = m.get_engine(ids.device, L); keep = busy()
rep("E2 fresh engine, GPU busy, first forward FROM FRAMES", eng.forward(img, ids, mask, True))
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device)
emb = bsv.embed(img); rep("E3 engine,bsv,embed,fwd (the failing flow)", eng.forward(None, ids, mask, True, emb))
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device)
emb = bsv.embed(img); torch.cuda.synchronize(); rep("E4 same with a synchronize before the forward", eng.forward(None, ids, mask, True, emb))
