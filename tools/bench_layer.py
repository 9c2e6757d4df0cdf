"""Times single conv/GEMM layers of the scorer on a B200 (CUDA events, L2-cold by rotating buffers).
Usage: python tools/bench_layer.py <case> [iters]
Cases: conv3_l1 conv1_l1 conv2_l1 conv3_l2 conv3_l3 conv3_l4 ffn_in ffn_out attn_out qkv stem all
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
import torch
from vcg_b200 import ops, binding as B

dev = "cuda"
NF = int(os.environ.get("NF", "512"))   # frames (32 clips x 16)
T = 16
GEMM_M = int(os.environ.get("GEMM_M", "25600"))   # rows of the BERT GEMMs (256 clips x 100 tokens; ~14080 when packed)

def conv_case(H, Cin, Cout, k, stride, res, tsm_out, tsm_in, act=B.ACT_RELU, nbuf=3):
    xs = [torch.randn(NF, H, H, Cin, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    w = (torch.randn(Cout, k, k, Cin, device=dev) / (k * k * Cin) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=dev)
    Ho = H // stride
    rs = [torch.randn(NF, Ho, Ho, Cout, device=dev).to(torch.bfloat16) for _ in range(nbuf)] if res else [None] * nbuf
    fold = Cout // 8
    to = torch.zeros(NF, Ho, Ho, 2 * fold, device=dev, dtype=torch.bfloat16) if tsm_out else None
    ti = torch.randn(NF, H, H, Cin // 4, device=dev).to(torch.bfloat16) if tsm_in else None
    flops = 2.0 * NF * Ho * Ho * Cout * k * k * Cin
    byts = 2.0 * NF * (H * H * Cin + Ho * Ho * Cout * (2 if res else 1) + (Ho * Ho * 2 * fold if tsm_out else 0))
    def run(i):
        ops.conv2d_nhwc(xs[i % nbuf], w, b, rs[i % nbuf], stride, act, tsm_in=ti, tsm_out=to, tsm_fold=fold if tsm_out else 0, clip_frames=T)
    return run, flops, byts

def gemm_case(M, N, K, act, res, nbuf=3):
    as_ = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    rs = [torch.randn(M, N, device=dev).to(torch.bfloat16) for _ in range(nbuf)] if res else [None] * nbuf
    outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    def run(i):
        ops.gemm(as_[i % nbuf], w, b, rs[i % nbuf], act, out=outs[i % nbuf])
    return run, 2.0 * M * N * K, 2.0 * (M * K + M * N * (2 if res else 1))

CASES = {
    "conv3_l1": lambda: conv_case(56, 64, 256, 1, 1, True, True, False),
    "conv1_l1": lambda: conv_case(56, 256, 64, 1, 1, False, False, True),
    "conv2_l1": lambda: conv_case(56, 64, 64, 3, 1, False, False, False),
    "conv3_l2": lambda: conv_case(28, 128, 512, 1, 1, True, True, False),
    "conv1_l2": lambda: conv_case(28, 512, 128, 1, 1, False, False, True),
    "conv2_l2": lambda: conv_case(28, 128, 128, 3, 1, False, False, False),
    "conv3_l3": lambda: conv_case(14, 256, 1024, 1, 1, True, True, False),
    "conv3_l3_nt": lambda: conv_case(14, 256, 1024, 1, 1, True, False, False),
    "conv3_l4_nt": lambda: conv_case(7, 512, 2048, 1, 1, True, False, False),
    "conv1_l3": lambda: conv_case(14, 1024, 256, 1, 1, False, False, True),
    "conv2_l3": lambda: conv_case(14, 256, 256, 3, 1, False, False, False),
    "conv3_l4": lambda: conv_case(7, 512, 2048, 1, 1, True, True, False),
    "conv1_l4": lambda: conv_case(7, 2048, 512, 1, 1, False, False, True),
    "conv2_l4": lambda: conv_case(7, 512, 512, 3, 1, False, False, False),
    "ds_l2": lambda: conv_case(56, 256, 512, 1, 2, False, False, False, act=B.ACT_NONE),
    "qkv": lambda: gemm_case(GEMM_M, 2304, 768, B.ACT_NONE, False),
    "attn_out": lambda: gemm_case(GEMM_M, 768, 768, B.ACT_NONE, True),
    "attn_out_nores": lambda: gemm_case(GEMM_M, 768, 768, B.ACT_NONE, False),
    "ffn_in": lambda: gemm_case(GEMM_M, 3072, 768, B.ACT_GELU, False),
    "ffn_out": lambda: gemm_case(GEMM_M, 768, 3072, B.ACT_NONE, True),
    "ffn_in_noact": lambda: gemm_case(GEMM_M, 3072, 768, B.ACT_NONE, False),
    "ffn_in_relu": lambda: gemm_case(GEMM_M, 3072, 768, B.ACT_RELU, False),
}

def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    names = list(CASES) if which == "all" else which.split(",")
    for name in names:
        run, flops, byts = CASES[name]()
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print(f"{name:10s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s  {byts/ms/1e6:8.1f} GB/s (algorithmic)", flush=True)

main()
