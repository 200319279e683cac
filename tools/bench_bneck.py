"""Micro-benchmark of the fused DepthBottleneckUni kernel (K4) against the two-kernel form it replaces, at the shapes
of the N variant (bs 32), plus a clock64 timeline of CTA 0 (which role waits for which)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maf_yolo_b200 import ops, _lib

dev = torch.device("cuda")
SHAPES = [("L2", 24, 3, 160, 160), ("L4", 48, 5, 80, 80), ("L20", 64, 5, 80, 80)]
n = 32


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


for name, c_, k, h, w in SHAPES:
    mid = 3 * c_
    g = torch.Generator().manual_seed(1)
    w1 = torch.randn(mid, c_, generator=g) / c_ ** 0.5; b1 = torch.randn(mid, generator=g) * 0.5
    wd = torch.randn(mid, 1, k, k, generator=g) / k; bd = torch.randn(mid, generator=g) * 0.5
    w2 = torch.randn(c_, mid, generator=g) / mid ** 0.5; b2 = torch.randn(c_, generator=g)
    ld = (3 * c_ + 15) // 16 * 16
    cat = ops.NHWC(torch.randn((n, h, w, ld), device=dev).half(), 0, 3 * c_)
    packed = ops.pack_bottleneck(w1, b1, wd, bd, w2, b2, device=dev)
    t1 = ops.NHWC.empty(n, h, w, mid, dev, ld=(mid + 15) // 16 * 16)
    pw1 = ops.pack_conv1x1(w1, b1, [c_], dev)
    pdw = ops.pack_dw(wd, bd, dev)
    pw2 = ops.pack_conv1x1(w2, b2, [mid], dev)
    src, dst = cat.slice(c_, c_), cat.slice(2 * c_, c_)
    fused = timeit(lambda: ops.bottleneck(src, packed, dst))

    def two():
        ops.conv1x1([src], *pw1, "silu", t1)
        ops.dwconv_conv1x1(t1, *pdw, k, "silu", *pw2, "silu", dst)

    unfused = timeit(two)
    steps = n * ((h + 9) // 10) * ((w + 19) // 20) * ((mid + 63) // 64) / 148
    print(f"{name}: c_={c_} mid={mid} k={k} {h}x{w} bs{n}: fused {fused:.1f} us, two kernels {unfused:.1f} us; "
          f"{steps:.1f} steps/SM -> {fused * 1.9e3 / steps:.0f} clk/step (at 1.9 GHz)")
    if name == "L20" and os.environ.get("MAFB200_BNECK_SHAPE") == "1":
        tr = torch.zeros(3 * 64 * 4, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().mafb200_bottleneck_trace(tr.data_ptr()))
        ops.bottleneck(src, packed, dst)
        torch.cuda.synchronize()
        _lib.check(_lib.lib().mafb200_bottleneck_trace(None))
        t = tr.view(3, 64, 4).cpu()
        t0 = int(t[t > 0].min())
        print("step | MMA: mma1 issued, mma2 issued | EPI: top, acc1 ready, T1 free, done | TAP: top, T1 ready, taps done, A2 written   (cycles since first stamp)")
        for s in range(14):
            row = [int(v) - t0 if v > 0 else -1 for v in t[:, s, :].flatten().tolist()]
            print(f"{s:3d} | {row[0]:7d} {row[1]:7d} | {row[4]:7d} {row[5]:7d} {row[6]:7d} {row[7]:7d} | {row[8]:7d} {row[9]:7d} {row[10]:7d} {row[11]:7d}")
