"""Library-kernel bar on the same B200: the reference forward (oracle restatement = the reference's own op sequence)
run by PyTorch/cuDNN/cuBLAS on the GPU, fp32 and autocast-bf16.  Not part of the product path."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import two_stream_oracle as orc, weights as W

T, L = 16, 100
torch.backends.cudnn.benchmark = True
dev = "cuda"
res = {}
with torch.no_grad():
    sd = {k: v.to(dev) for k, v in W.make_state_dict(T, "mlp", seed=123).items()}
    for name, B, vision in (("config1_precomputed_B256", 256, False), ("pipeline_B32", 32, True)):
        ids, mask = W.make_text(B, L, seed=3)
        ids, mask = ids.to(dev), mask.to(dev)
        emb = torch.rand(B, T, 2048, device=dev) * 2
        img = torch.randn(B, T, 3, 224, 224, device=dev) if vision else None
        if img is not None:
            img = img.contiguous(memory_format=torch.channels_last_3d) if False else img
        for mode in ("fp32", "bf16"):
            def run():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                    # monkey-free: the oracle builds masks on CPU by default; move them
                    return orc.two_stream_forward(sd, img, ids, mask, T, 128, "mlp", 8, vision_emb=None if vision else emb)
            try:
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                n = 5
                t0 = time.perf_counter()
                for _ in range(n):
                    run()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / n
                res[f"{name}_{mode}"] = {"clips_per_s": B / dt, "ms": dt * 1e3}
            except Exception as ex:
                res[f"{name}_{mode}"] = {"error": str(ex)[:200]}
            print(name, mode, res[f"{name}_{mode}"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "torch_gpu_baseline.json"), "w"), indent=1)
