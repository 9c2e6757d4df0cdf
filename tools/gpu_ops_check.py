"""Op-level bring-up check on a B200: every C-ABI operator against a plain torch fp32 reference.
Usage (on the GPU box): python tools/gpu_ops_check.py [bf16|fp32|all]
"""
import os, sys, json, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
import torch
import torch.nn.functional as F
from vcg_b200 import ops, binding as B

torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False          # the torch reference must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
results = []

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()

def report(name, err, tol):
    ok = err <= tol
    results.append({"name": name, "err": err, "tol": tol, "ok": ok})
    print(f"{'OK  ' if ok else 'FAIL'} {name}: rel err {err:.3e} (tol {tol:.1e})", flush=True)

def run(name, fn):
    try:
        fn()
        torch.cuda.synchronize()
    except Exception as ex:
        traceback.print_exc()
        results.append({"name": name, "err": None, "ok": False, "exc": str(ex)})
        print(f"FAIL {name}: {ex}", flush=True)

def t_gemm(dtype, M, N, K, act=B.ACT_NONE, bias=True, res=False, tol=None):
    def f():
        a = torch.randn(M, K, device=dev).to(dtype)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dtype)
        b = torch.randn(N, device=dev) if bias else None
        r = torch.randn(M, N, device=dev).to(dtype) if res else None
        out = ops.gemm(a, w, b, r, act)
        ref = a.float() @ w.float().t()
        if bias: ref = ref + b
        if res: ref = ref + r.float()
        if act == B.ACT_RELU: ref = F.relu(ref)
        if act == B.ACT_GELU: ref = F.gelu(ref)
        if act == B.ACT_TANH: ref = torch.tanh(ref)
        report(f"gemm {dtype} M{M} N{N} K{K} act{act} bias{bias} res{res}", rel(out, ref), tol or (1e-2 if dtype == torch.bfloat16 else 2e-5))
    run(f"gemm {dtype} {M}x{N}x{K}", f)

def t_conv(dtype, n, H, Cin, Cout, k, stride, act=B.ACT_RELU, res=False, tol=None):
    def f():
        x = torch.randn(n, H, H, Cin, device=dev).to(dtype)
        w = (torch.randn(Cout, k, k, Cin, device=dev) / (k * k * Cin) ** 0.5).to(dtype)
        b = torch.randn(Cout, device=dev)
        r = torch.randn(n, H // stride, H // stride, Cout, device=dev).to(dtype) if res else None
        out = ops.conv2d_nhwc(x, w, b, r, stride, act)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b, stride=stride, padding=k // 2)
        ref = ref.permute(0, 2, 3, 1)
        if res: ref = ref + r.float()
        if act == B.ACT_RELU: ref = F.relu(ref)
        report(f"conv {dtype} n{n} H{H} Cin{Cin} Cout{Cout} k{k} s{stride} res{res}", rel(out, ref), tol or (1e-2 if dtype == torch.bfloat16 else 2e-5))
    run(f"conv {dtype} n{n} H{H} {Cin}->{Cout} k{k} s{stride}", f)

def t_tsm(dtype, B_, T, H, C, planes):
    """1x1 conv with the temporal shift folded into the operand load + epilogue scatter for the next block."""
    def f():
        n = B_ * T
        fold = C // 8
        x = torch.randn(n, H, H, C, device=dev).to(dtype)
        # reference shift (ops/temporal_shift.py:34-51) on NHWC
        xv = x.view(B_, T, H, H, C)
        sh = torch.zeros_like(xv)
        sh[:, :-1, ..., :fold] = xv[:, 1:, ..., :fold]
        sh[:, 1:, ..., fold:2 * fold] = xv[:, :-1, ..., fold:2 * fold]
        sh[..., 2 * fold:] = xv[..., 2 * fold:]
        sh = sh.view(n, H, H, C)
        tsm_in = sh[..., :2 * fold].contiguous()
        w = (torch.randn(planes, 1, 1, C, device=dev) / C ** 0.5).to(dtype)
        b = torch.randn(planes, device=dev)
        out = ops.conv2d_nhwc(x, w, b, None, 1, B.ACT_RELU, tsm_in=tsm_in)
        ref = F.relu(F.conv2d(sh.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b)).permute(0, 2, 3, 1)
        report(f"tsm-load conv1x1 {dtype} B{B_} T{T} H{H} C{C}->{planes}", rel(out, ref), 1e-2 if dtype == torch.bfloat16 else 2e-5)
        # epilogue scatter: conv producing C channels writes the shifted buffer for the next block
        w2 = (torch.randn(C, 1, 1, planes, device=dev) / planes ** 0.5).to(dtype)
        b2 = torch.randn(C, device=dev)
        tsm_out = torch.zeros(n, H, H, 2 * fold, device=dev, dtype=dtype)
        y = ops.conv2d_nhwc(out, w2, b2, x, 1, B.ACT_RELU, tsm_out=tsm_out, tsm_fold=fold, clip_frames=T)
        yv = y.view(B_, T, H, H, C)
        exp = torch.zeros(B_, T, H, H, 2 * fold, device=dev, dtype=dtype)
        exp[:, :-1, ..., :fold] = yv[:, 1:, ..., :fold]
        exp[:, 1:, ..., fold:] = yv[:, :-1, ..., fold:2 * fold]
        report(f"tsm-scatter {dtype} B{B_} T{T} H{H} C{C}", (tsm_out.float() - exp.view(n, H, H, 2 * fold).float()).abs().max().item(), 0.0)
    run(f"tsm {dtype}", f)

def stem_to_nhwc4(xp, dtype):
    """stem input buffer -> plain padded NHWC4 view [n,230,240,4] (bf16 stores row pairs interleaved per pixel)."""
    if dtype == torch.bfloat16:
        n = xp.shape[0]
        return xp.view(n, 115, 240, 2, 4).permute(0, 1, 3, 2, 4).reshape(n, 230, 240, 4)
    return xp

def t_stem(dtype, n):
    def f():
        frames = torch.randint(0, 256, (n, 224, 224, 3), device=dev, dtype=torch.uint8)
        xp_raw = ops.preprocess_u8(frames, None, dtype)
        xp = stem_to_nhwc4(xp_raw, dtype)
        mean = torch.tensor([0.485, 0.456, 0.406], device=dev); std = torch.tensor([0.229, 0.224, 0.225], device=dev)
        img = ((frames.float() / 255.0) - mean) / std            # NHWC fp32
        interior = xp[:, 3:227, 3:227, :3].float()
        report(f"preprocess_u8 {dtype}", rel(interior, img), 4e-3 if dtype == torch.bfloat16 else 1e-6)
        border = xp.clone(); border[:, 3:227, 3:227, :] = 0
        report(f"preprocess border zero {dtype}", border.float().abs().max().item() + xp[..., 3].float().abs().max().item(), 0.0)
        xp2 = ops.nchw_to_stem(img.permute(0, 3, 1, 2).contiguous(), dtype)
        report(f"nchw_to_stem == preprocess {dtype}", (xp2.float() - xp_raw.float()).abs().max().item(), 4e-2 if dtype == torch.bfloat16 else 1e-6)
        w = torch.randn(64, 3, 7, 7, device=dev) / 147 ** 0.5
        b = torch.randn(64, device=dev)
        if dtype == torch.bfloat16:      # [64][4 row pairs][8 px][2 rows][4 ch]
            w8 = torch.zeros(64, 8, 8, 4, device=dev)          # [co][kh][kw][c], kh = 7 / kw = 7 / c = 3 zero
            w8[:, :7, :7, :3] = w.permute(0, 2, 3, 1)
            wp = w8.view(64, 4, 2, 8, 4).permute(0, 1, 3, 2, 4).contiguous().to(dtype)
            wq = w8[:, :7, :7, :3].to(dtype).float()
        else:                            # [64][7][8 px][4 ch]
            w7 = torch.zeros(64, 7, 8, 4, device=dev)
            w7[:, :, :7, :3] = w.permute(0, 2, 3, 1)
            wp = w7.contiguous()
            wq = w7[:, :, :7, :3]
        out = ops.stem_conv(xp_raw, wp, b)
        xin = xp[:, 3:227, 3:227, :3].float().permute(0, 3, 1, 2)
        ref = F.relu(F.conv2d(xin, wq.permute(0, 3, 1, 2), b, stride=2, padding=3)).permute(0, 2, 3, 1)
        report(f"stem conv {dtype} n{n}", rel(out, ref), 1e-2 if dtype == torch.bfloat16 else 2e-5)
        T = 4 if n % 4 == 0 else 1
        pooled, shifted = ops.maxpool_tsm(out, T, 8)
        refp = F.max_pool2d(out.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        report(f"maxpool {dtype}", (pooled.float() - refp).abs().max().item(), 0.0)
        pv = refp.view(n // T, T, 56, 56, 64)
        sh = torch.zeros_like(pv)
        sh[:, :-1, ..., :8] = pv[:, 1:, ..., :8]
        sh[:, 1:, ..., 8:16] = pv[:, :-1, ..., 8:16]
        sh[..., 16:] = pv[..., 16:]
        report(f"maxpool shifted {dtype}", (shifted.float() - sh.view(n, 56, 56, 64)).abs().max().item(), 0.0)
    run(f"stem {dtype}", f)

def t_attn(dtype, B_, L):
    def f():
        qkv = torch.randn(B_ * L, 2304, device=dev).to(dtype)
        lens = torch.randint(1, L + 1, (B_,), device=dev)
        mask = (torch.arange(L, device=dev)[None, :] < lens[:, None]).long()
        ctx = ops.bert_attention(qkv, mask, B_, L)
        q, k, v = [t.float().view(B_, L, 12, 64).transpose(1, 2) for t in qkv.split(768, dim=1)]
        add = torch.zeros(B_, 1, 1, L, device=dev).masked_fill(mask[:, None, None, :] == 0, float("-inf"))
        att = torch.softmax(q @ k.transpose(-1, -2) / 8.0 + add, dim=-1)
        ref = (att @ v).transpose(1, 2).reshape(B_ * L, 768)
        report(f"attention {dtype} B{B_} L{L}", rel(ctx, ref), 1.5e-2 if dtype == torch.bfloat16 else 2e-5)
    run(f"attention {dtype} B{B_} L{L}", f)

def t_ln(dtype, rows):
    def f():
        x = torch.randn(rows, 768, device=dev).to(dtype) * 3 + 1
        g = torch.randn(768, device=dev); b = torch.randn(768, device=dev)
        y = ops.layernorm(x, g, b, 1e-12)
        ref = F.layer_norm(x.float(), (768,), g, b, 1e-12)
        report(f"layernorm {dtype} rows{rows}", rel(y, ref), 1e-2 if dtype == torch.bfloat16 else 1e-5)
    run(f"layernorm {dtype}", f)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dtypes = {"bf16": [torch.bfloat16], "fp32": [torch.float32], "all": [torch.bfloat16, torch.float32]}[which]
print(B.load_library().vcg_version().decode(), torch.cuda.get_device_name(0), flush=True)
for dt in dtypes:
    t_gemm(dt, 128, 64, 64, bias=False)
    t_gemm(dt, 128, 128, 64)
    t_gemm(dt, 128, 256, 128)
    t_gemm(dt, 1000, 768, 768, res=True)
    t_gemm(dt, 3200, 2304, 768)
    t_gemm(dt, 777, 3072, 768, act=B.ACT_GELU)
    t_gemm(dt, 25600, 768, 3072, res=True)
    t_gemm(dt, 8, 768, 768, act=B.ACT_TANH)
    t_conv(dt, 4, 56, 64, 64, 1, 1)
    t_conv(dt, 4, 56, 64, 256, 1, 1, res=True)
    t_conv(dt, 4, 56, 64, 64, 3, 1)
    t_conv(dt, 16, 56, 128, 128, 3, 2)
    t_conv(dt, 16, 56, 256, 512, 1, 2, act=B.ACT_NONE)
    t_conv(dt, 16, 28, 128, 128, 3, 1)
    t_conv(dt, 32, 28, 256, 256, 3, 2)
    t_conv(dt, 32, 14, 256, 256, 3, 1)
    t_conv(dt, 32, 14, 512, 512, 3, 2)
    t_conv(dt, 32, 7, 512, 512, 3, 1)
    t_conv(dt, 128, 7, 512, 2048, 1, 1, res=True)
    t_conv(dt, 128, 7, 512, 512, 3, 1)
    t_tsm(dt, 2, 4, 28, 512, 128)
    t_tsm(dt, 2, 8, 7, 2048, 512)
    t_stem(dt, 4)
    t_attn(dt, 3, 100)
    t_attn(dt, 2, 512)
    t_attn(dt, 4, 17)
    t_ln(dt, 1003)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(results, open(os.path.join(ROOT, "gpurun_out", f"ops_check_{which}.json"), "w"), indent=1)
bad = [r for r in results if not r["ok"]]
print(f"{len(results) - len(bad)}/{len(results)} passed")
sys.exit(1 if bad else 0)
