"""Times the fused resize + pre-processing kernel (for ncu / GB/s).  Usage: python tools/prof_resize.py [Hs Ws n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
import torch
from vcg_b200 import ops
Hs, Ws, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (360, 640, 256)
frames = torch.randint(0, 256, (n, Hs, Ws, 3), dtype=torch.uint8, device="cuda")
small = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
from vcg_b200 import binding as _b
lib = _b.load_library()
stem = torch.zeros(n, 230, 240, 4, dtype=torch.bfloat16, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for name, fn, byts in (("resize_preprocess %dx%d" % (Hs, Ws), lambda: _b.check(lib.vcg_op_resize_u8(frames.data_ptr(), n, Hs, Ws, 0, stem.data_ptr(), _b.PREC_BF16, st)), n * (Hs * Ws * 3 + 224 * 224 * 3 * 2)),
                       ("preprocess 224x224", lambda: _b.check(lib.vcg_op_preprocess_u8(small.data_ptr(), 0, n, stem.data_ptr(), _b.PREC_BF16, st)), n * (224 * 224 * 3 * 3))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {n} frames {ms*1e3:.1f} us/call   {byts/ms/1e6:.0f} GB/s algorithmic")
