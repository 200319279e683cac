"""Per-layer timing of the fused depth-wise + 1x1 kernel (mafb200_dwconv_conv1x1) at the shapes of the N variant
(bs 32): CUDA events, 20 launches after 3 warm-ups; inputs (>= 100 MB) exceed L2 only at the 160 / 80-pixel levels.
Used with MAFB200_LIB=<ablation build> (-DMAFB200_DWPW_ABL=n: results wrong, timing only) to see which phase of the
kernel the time belongs to."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maf_yolo_b200 import ops

dev = torch.device("cuda")
SHAPES = [("L2", 72, 24, 3, 160, 160, "silu"), ("L4", 144, 48, 5, 80, 80, "silu"), ("L20", 192, 64, 5, 80, 80, "silu"),
          ("L31", 64, 64, 5, 80, 80, "none")]
n = int(os.environ.get("BS", "32"))


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / reps


out = []
for name, c, cout, k, h, w, act1 in SHAPES:
    g = torch.Generator().manual_seed(1)
    wd = torch.randn(c, 1, k, k, generator=g) / k; bd = torch.randn(c, generator=g) * 0.5
    w2 = torch.randn(cout, c, generator=g) / c ** 0.5; b2 = torch.randn(cout, generator=g)
    src = ops.NHWC(torch.randn((n, h, w, (c + 15) // 16 * 16), device=dev).half(), 0, c)
    dst = ops.NHWC.empty(n, h, w, cout, dev, ld=(cout + 15) // 16 * 16)
    pdw = ops.pack_dw(wd, bd, dev)
    pw2 = ops.pack_conv1x1(w2, b2, [c], dev)
    t = timeit(lambda: ops.dwconv_conv1x1(src, *pdw, k, act1, *pw2, "silu", dst))
    out.append(f"{name} {t:.1f}")
print(os.environ.get("MAFB200_LIB", "default"), " | ".join(out), "us")

