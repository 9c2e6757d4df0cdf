import sys, torch
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/video-chapter-generation_b200')
from test_bn_batch import golden_case, rel
from test_parity_gpu import build_model
g,T,L,B,ids,mask,img = golden_case("attn_T8_L32_B2")
ids, mask, img = ids.cuda(), mask.cuda(), img.cuda()
gv = torch.tensor(g["vision_emb"]).cuda()
def lang_err(t): return "%.2e" % rel(t, g["lang_emb"])
print("A: with-backbone bf16 engines, vision_emb given, no batch-stat operators", flush=True)
for i in range(3):
    model,_ = build_model(T, "attn", "bf16")
    eng = model.get_engine(ids.device, L)
    o = eng.forward(None, ids, mask, return_emb=True, vision_emb=gv)
    o2 = eng.forward(img, ids, mask, return_emb=True)
    o3 = eng.forward(None, ids, mask, return_emb=True, vision_emb=gv)
    print("  engine", i, "emb-in lang", lang_err(o[3]), "| frames lang", lang_err(o2[3]), "| emb-in again", lang_err(o3[3]), flush=True)
print("B: batch-stat mode bf16, fresh models; two calls each, then a standard call", flush=True)
for i in range(3):
    model,_ = build_model(T, "attn", "bf16")
    model.bn_batch_stats, model.bn_batch_precision = True, "bf16"
    o = model(img, ids, mask, return_emb=True)
    o2 = model(img, ids, mask, return_emb=True)
    model.bn_batch_stats = False
    o3 = model(img, ids, mask, return_emb=True)
    print("  model", i, "lang", lang_err(o[3]), lang_err(o2[3]), "| standard", lang_err(o3[3]), "| logits", "%.2e %.2e" % (rel(o[0], g["logits"]), rel(o2[0], g["logits"])), flush=True)
print("C: batch-stat vision ops first (fp32 ops), then a bf16 engine emb-in call", flush=True)
for i in range(2):
    model,_ = build_model(T, "attn", "bf16")
    from vcg_b200.bn_batch import BatchStatVision
    v = BatchStatVision(model.state_dict(), T, 8, "fp32", ids.device).embed(img)
    torch.cuda.synchronize()
    eng = model.get_engine(ids.device, L)
    o = eng.forward(None, ids, mask, return_emb=True, vision_emb=v)
    print("  model", i, "vis %.2e" % rel(v, g["vision_emb"]), "lang", lang_err(o[3]), flush=True)
