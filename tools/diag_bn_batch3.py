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
    print(f"{tag}: lang {rel(l, g['lang_emb']):.2e} finite {bool(torch.isfinite(l).all())} absmax {float(l.abs().max()):.3f} logits {rel(o[0], g['logits']):.2e}", flush=True)
def fresh():
    model,_ = build_model(T, "attn", "bf16")
    return model
sync = torch.cuda.synchronize
# V3: engine, bsv, embed, forward — no sync
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device)
emb = bsv.embed(img); rep("V3 engine,bsv,embed,fwd (no sync)", eng.forward(None, ids, mask, True, emb))
# V2: sync between embed and forward
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device)
emb = bsv.embed(img); sync(); rep("V2 ... sync before fwd", eng.forward(None, ids, mask, True, emb))
# V4: bsv before the engine
m = fresh(); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device); eng = m.get_engine(ids.device, L)
emb = bsv.embed(img); rep("V4 bsv,engine,embed,fwd", eng.forward(None, ids, mask, True, emb))
# V5: sync after bsv init only
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device); sync()
emb = bsv.embed(img); rep("V5 engine,bsv,sync,embed,fwd", eng.forward(None, ids, mask, True, emb))
# V6: engine warmed first
m = fresh(); eng = m.get_engine(ids.device, L); eng.forward(None, ids, mask, True, gv)
bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device); emb = bsv.embed(img)
rep("V6 engine warmed, bsv, embed, fwd", eng.forward(None, ids, mask, True, emb))
# V7: bsv init only, golden embeddings
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device)
rep("V7 engine,bsv init only,fwd(golden vis)", eng.forward(None, ids, mask, True, gv))
# V8: embed twice before the first forward
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "bf16", ids.device)
emb = bsv.embed(img); emb = bsv.embed(img); rep("V8 engine,bsv,embed x2,fwd", eng.forward(None, ids, mask, True, emb))
# V9: fp32 operators, bf16 engine, no sync
m = fresh(); eng = m.get_engine(ids.device, L); bsv = BatchStatVision(m.state_dict(), T, 8, "fp32", ids.device)
emb = bsv.embed(img); rep("V9 fp32 operators, bf16 engine, no sync", eng.forward(None, ids, mask, True, emb))
