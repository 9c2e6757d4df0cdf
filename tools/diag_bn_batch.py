import sys, torch
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/video-chapter-generation_b200')
from test_bn_batch import golden_case, rel
from test_parity_gpu import build_model
for name, head in (("attn_T8_L32_B2","attn"),("mlp_T16_L100_B3","mlp")):
    g,T,L,B,ids,mask,img = golden_case(name)
    for eng_prec in ("fp32","bf16"):
      for mode_prec in ("fp32","bf16"):
        model,_ = build_model(T, head, eng_prec)
        model.bn_batch_stats, model.bn_batch_precision = True, mode_prec
        lo, pr, vis, lang = model(img.cuda(), ids.cuda(), mask.cuda(), return_emb=True)
        print(name, "engine", eng_prec, "mode", mode_prec, "vis %.2e lang %.2e logits %.2e (abs %.2e) probs %.2e labels %s" % (
            rel(vis,g["vision_emb"]), rel(lang,g["lang_emb"]), rel(lo,g["logits"]), float((lo.cpu()-torch.tensor(g["logits"])).abs().max()),
            rel(pr,g["probs"]), lo.topk(1,1)[1].view(-1).tolist()==g["labels"].tolist()), flush=True)
