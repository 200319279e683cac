"""GPU parity tests of every kernel behind the C ABI, one op at a time.

Floating-point kernels are compared with a plain PyTorch fp32 CPU computation of the same op on
the same fp16-rounded inputs/weights (tolerance: fp16 output rounding, stated per test); the NMS
kernel is compared bit-exactly with the oracle (oracle/nms.py), which is itself pinned
bit-exactly against the reference's non_max_suppression.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL = 2e-3  # fp16 output rounding is 2^-11 = 4.9e-4 relative; accumulation order adds a little
ATOL = 2e-3


def _act_ref(y, act):
    return {"none": lambda t: t, "silu": F.silu, "relu": F.relu, "sigmoid": torch.sigmoid}[act](y)


def _close(got, ref, what, rtol=RTOL, atol=ATOL):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance, max abs err "
                           f"{err.max().item():.4e} at {np.unravel_index(int(err.argmax()), err.shape)}, "
                           f"ref there {ref.flatten()[err.argmax()].item():.4e}")


# every distinct (sources -> cout @ H) family of the N/S/M graphs, small batch
CONV1X1_CASES = [
    # (src_channels, cout, h, w, n, act)
    ([48], 48, 20, 20, 2, "silu"),
    ([24], 72, 40, 40, 1, "silu"),       # K < 64 (zero-filled K tail), N not multiple of 16
    ([72], 24, 40, 40, 1, "silu"),       # K = 64 + 8 tail
    ([144], 96, 20, 12, 3, "silu"),      # M not a multiple of 128
    ([288], 192, 20, 20, 2, "silu"),
    ([192], 576, 20, 20, 1, "silu"),     # 3 N tiles
    ([576], 384, 20, 20, 1, "silu"),     # 9 K blocks > 4 stages, 2 N tiles
    ([768], 384, 20, 20, 2, "silu"),     # 12 K blocks
    ([96, 384], 192, 20, 20, 2, "silu"),             # L11 concat
    ([64, 192, 192], 128, 40, 40, 1, "silu"),        # L15 concat
    ([128, 128, 128, 192], 128, 40, 40, 2, "silu"),  # L25 concat (4 sources)
    ([24, 24, 24], 48, 40, 40, 1, "silu"),           # RepHDW conv2 at c_=24 (three 24-ch slices)
    ([128], 80, 40, 40, 2, "none"),      # cls_pred
    ([128], 68, 40, 40, 2, "none"),      # reg_pred (cout % 8 == 4)
    ([192], 80, 20, 20, 2, "sigmoid"),
    ([1536], 768, 20, 20, 1, "silu"),    # M variant, 24 K blocks, 3 N tiles of 256
    ([64], 64, 7, 9, 1, "relu"),         # tiny M < 128
]


@pytest.mark.parametrize("src_channels,cout,h,w,n,act", CONV1X1_CASES)
def test_conv1x1(cuda_device, src_channels, cout, h, w, n, act):
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(1234 + cout + sum(src_channels))
    ktot = sum(src_channels)
    xs = [torch.randn(n, c, h, w, generator=g).half().float() for c in src_channels]
    wgt = (torch.randn(cout, ktot, generator=g) / ktot ** 0.5).half().float()
    bias = torch.randn(cout, generator=g)
    ref = _act_ref(F.conv2d(torch.cat(xs, 1), wgt[:, :, None, None], bias), act)

    # sources are channel slices of ONE wider buffer when they all have equal spatial size:
    # exercises c_stride > c and non-zero channel offsets (zero-copy concat views)
    wide = ops.NHWC.from_nchw(torch.cat(xs + [torch.randn(n, 8, h, w, generator=g)], 1).to(cuda_device))
    srcs, off = [], 0
    for c in src_channels:
        srcs.append(wide.slice(off, c))
        off += c
    wp, bp = ops.pack_conv1x1(wgt, bias, src_channels, cuda_device)
    dst_buf = ops.NHWC.empty(n, h, w, cout + 16, cuda_device)
    dst_buf.buf.fill_(-77.0)
    dst = dst_buf.slice(8, cout)
    ops.conv1x1(srcs, wp, bp, act, dst)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, f"conv1x1 {src_channels}->{cout}")
    # neighbours of the destination slice must be untouched
    assert (dst_buf.buf[..., :8] == -77.0).all() and (dst_buf.buf[..., 8 + cout:] == -77.0).all()


def test_conv1x1_fused_upsample(cuda_device):
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(7)
    n, cin, cout, h, w = 2, 96, 192, 20, 20
    x = torch.randn(n, cin, h, w, generator=g).half().float()
    wgt = (torch.randn(cout, cin, generator=g) / cin ** 0.5).half().float()
    bias = torch.randn(cout, generator=g)
    ref = F.silu(F.conv2d(x, wgt[:, :, None, None], bias))
    src = ops.NHWC.from_nchw(x.to(cuda_device))
    wp, bp = ops.pack_conv1x1(wgt, bias, [cin], cuda_device)
    dst = ops.NHWC.empty(n, h, w, cout, cuda_device)
    up = ops.NHWC.empty(n, 2 * h, 2 * w, cout, cuda_device)
    ops.conv1x1([src], wp, bp, "silu", dst, up)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, "conv1x1+up: main output")
    assert torch.equal(up.to_nchw(), F.interpolate(dst.to_nchw(), scale_factor=2, mode="nearest")), "upsampled copy"


CONV3X3_CASES = [
    # (cin, cout, h, w, n, act)
    (24, 48, 64, 64, 1, "relu"),      # L1-like: cin < 64
    (48, 48, 40, 40, 2, "relu"),
    (96, 64, 40, 40, 1, "silu"),      # cin = 64 + 32 tail
    (128, 128, 40, 40, 2, "silu"),
    (192, 96, 40, 40, 1, "silu"),
    (192, 192, 20, 20, 3, "relu"),    # tiles straddle images (100 px / image)
    (384, 256, 20, 20, 1, "silu"),    # M variant, 54 K blocks
    (32, 64, 12, 20, 1, "relu"),      # non-square, M = 60 < 128
]


@pytest.mark.parametrize("cin,cout,h,w,n,act", CONV3X3_CASES)
def test_conv3x3s2(cuda_device, cin, cout, h, w, n, act):
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(99 + cin + cout)
    x = torch.randn(n, cin, h, w, generator=g).half().float()
    wgt = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).half().float()
    bias = torch.randn(cout, generator=g)
    ref = _act_ref(F.conv2d(x, wgt, bias, stride=2, padding=1), act)
    src = ops.NHWC.from_nchw(x.to(cuda_device))
    wp, bp = ops.pack_conv3x3(wgt, bias, cuda_device)
    dst = ops.NHWC.empty(n, h // 2, w // 2, cout, cuda_device)
    ops.conv3x3s2(src, wp, bp, act, dst)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, f"conv3x3s2 {cin}->{cout} @{h}x{w}")


@pytest.mark.parametrize("cin,cout,h,w,n,act", [(24, 48, 40, 64, 2, "relu"), (32, 64, 24, 20, 1, "silu"),
                                                (24, 48, 320, 320, 1, "relu"), (8, 16, 6, 10, 3, "none")])
def test_conv3x3s2_pair(cuda_device, cin, cout, h, w, n, act):
    """Pixel-pair im2col form of the 3x3 s2 conv (narrow inputs, channel stride 32): same result as the 9-tap form;
    the padding channels of the source hold finite junk that must not leak (they meet zero weights)."""
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(91 + cin + cout)
    x = torch.randn(n, cin, h, w, generator=g).half().float()
    wgt = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).half().float()
    bias = torch.randn(cout, generator=g)
    ref = _act_ref(F.conv2d(x, wgt, bias, stride=2, padding=1), act)
    src = ops.NHWC.from_nchw(x.to(cuda_device), ld=32)
    if cin < 32:
        src.buf[..., cin:] = 123.0  # junk in the padding channels
    wp, bp = ops.pack_conv3x3_pair(wgt, bias, 32, cuda_device)
    dst = ops.NHWC.empty(n, h // 2, w // 2, cout, cuda_device)
    ops.conv3x3s2_pair(src, wp, bp, act, dst)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, f"conv3x3s2_pair {cin}->{cout} @{h}x{w}")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.uint8])
@pytest.mark.parametrize("cout", [24, 32, 48])
def test_stem_conv(cuda_device, dtype, cout):
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(5 + cout)
    n, h, w = 2, 64, 96
    if dtype == torch.uint8:
        x = torch.randint(0, 256, (n, 3, h, w), generator=g, dtype=torch.uint8)
        xf = x.float() / 255
    else:
        x = torch.rand(n, 3, h, w, generator=g).to(dtype)
        xf = x.float()
    wgt = torch.randn(cout, 3, 3, 3, generator=g) / 27 ** 0.5
    bias = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(xf, wgt, bias, stride=2, padding=1))
    wp, bp = ops.pack_stem(wgt, bias, cuda_device)
    dst = ops.NHWC.empty(n, h // 2, w // 2, cout, cuda_device)
    ops.stem_conv3x3s2(x.to(cuda_device), wp, bp, "relu", dst)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, f"stem conv {dtype} cout={cout}")
    if dtype == torch.uint8:
        # raw uint8 input == the reference's `imgs.float() / 255` fed as fp32, bit for bit (the /255 folded into the kernel
        # is a multiplication whose fp16 rounding equals the division's for all 256 values)
        dst2 = ops.NHWC.empty(n, h // 2, w // 2, cout, cuda_device)
        ops.stem_conv3x3s2(xf.to(cuda_device), wp, bp, "relu", dst2)
        torch.cuda.synchronize()
        assert torch.equal(dst.to_nchw(), dst2.to_nchw())


DW_CASES = [(72, 3, 40, 40, 2, "silu"), (144, 5, 40, 40, 1, "silu"), (288, 7, 40, 40, 1, "silu"),
            (576, 9, 20, 20, 2, "silu"), (128, 5, 24, 20, 1, "none"), (192, 9, 20, 20, 1, "none"),
            (64, 7, 10, 15, 1, "silu"), (96, 3, 7, 13, 1, "relu")]


@pytest.mark.parametrize("c,k,h,w,n,act", DW_CASES)
def test_dwconv(cuda_device, c, k, h, w, n, act):
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(31 + c + k)
    x = torch.randn(n, c, h, w, generator=g).half().float()
    wgt = torch.randn(c, 1, k, k, generator=g) / k
    bias = torch.randn(c, generator=g)
    ref = _act_ref(F.conv2d(x, wgt, bias, padding=k // 2, groups=c), act)
    wide = ops.NHWC.from_nchw(torch.cat([torch.zeros(n, 8, h, w), x], 1).to(cuda_device))
    src = wide.slice(8, c)
    wp, bp = ops.pack_dw(wgt, bias, cuda_device)
    dst = ops.NHWC.empty(n, h, w, c, cuda_device)
    ops.dwconv(src, wp, bp, k, act, dst)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, f"dwconv c={c} k={k}")


@pytest.mark.parametrize("c,cout,k,h,w,n,act1,act2", [(192, 64, 5, 40, 40, 2, "silu", "silu"), (128, 128, 5, 24, 30, 1, "none", "silu"),
                                                      (72, 24, 3, 40, 40, 1, "silu", "silu"), (64, 32, 3, 7, 9, 2, "relu", "none"),
                                                      (144, 48, 5, 20, 20, 2, "silu", "silu"), (192, 64, 5, 80, 80, 1, "silu", "silu"),
                                                      # k = 3 with 2, 3 and 4 channel blocks (the halo buffer and both W2 slots are reused)
                                                      (144, 24, 3, 33, 41, 2, "silu", "silu"), (256, 32, 3, 20, 20, 1, "silu", "none"),
                                                      (96, 32, 3, 160, 160, 1, "silu", "silu"), (72, 64, 3, 20, 20, 1, "silu", "silu"),
                                                      # 16- and 8-channel tail blocks (fetched by cp.async) on ragged maps
                                                      (80, 32, 3, 23, 37, 2, "silu", "silu"), (136, 40, 5, 17, 26, 1, "none", "silu")])
def test_dwconv_conv1x1_fused(cuda_device, c, cout, k, h, w, n, act1, act2):
    """Fused depth-wise -> 1x1 kernel against the two reference ops in fp32 (the intermediate is rounded to fp16 once,
    exactly as the two separate kernels do)."""
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(57 + c + k)
    x = torch.randn(n, c, h, w, generator=g).half().float()
    dw_w = torch.randn(c, 1, k, k, generator=g) / k
    dw_b = torch.randn(c, generator=g) * 0.5
    pw_w = (torch.randn(cout, c, generator=g) / c ** 0.5).half().float()
    pw_b = torch.randn(cout, generator=g)
    mid = _act_ref(F.conv2d(x, dw_w, dw_b, padding=k // 2, groups=c), act1).half().float()
    ref = _act_ref(F.conv2d(mid, pw_w[:, :, None, None], pw_b), act2)
    src = ops.NHWC.from_nchw(x.to(cuda_device))
    dwp, dbp = ops.pack_dw(dw_w, dw_b, cuda_device)
    pwp, pbp = ops.pack_conv1x1(pw_w, pw_b, [c], cuda_device)
    dst = ops.NHWC.empty(n, h, w, cout, cuda_device, ld=(cout + 15) // 16 * 16)
    ops.dwconv_conv1x1(src, dwp, dbp, k, act1, pwp, pbp, act2, dst)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, f"dwconv_conv1x1 c={c}->{cout} k={k}", rtol=3e-3, atol=3e-3)


# every (c_, 3 c_, k) of the N / S / M graphs the fused kernel takes (c_ <= 64, k <= 5), full tiles and ragged edges,
# one to three 64-channel blocks (incl. the 8- and 16-channel tails of mid = 72 / 144), many tiles per CTA (persistence)
@pytest.mark.parametrize("c_,k,h,w,n", [(24, 3, 40, 44, 2), (48, 5, 24, 40, 2), (64, 5, 20, 30, 2), (32, 3, 16, 20, 1),
                                        (64, 5, 80, 80, 2), (48, 3, 13, 27, 3), (24, 3, 160, 160, 2), (64, 3, 7, 9, 1),
                                        (16, 5, 10, 20, 1), (64, 5, 160, 160, 4)])
def test_bottleneck_fused(cuda_device, c_, k, h, w, n):
    """K4 (mafb200_bottleneck): the whole DepthBottleneckUni against its three reference ops in fp32, the two
    intermediates rounded to fp16 once each (as the kernel keeps them in shared memory).  Source and destination are
    channel slices of one wider buffer, as in RepHDW's concat; the other channels must stay untouched."""
    from maf_yolo_b200 import ops

    mid = 3 * c_
    g = torch.Generator().manual_seed(131 + c_ + k + h)
    x = torch.randn(n, c_, h, w, generator=g).half().float()
    w1 = (torch.randn(mid, c_, generator=g) / c_ ** 0.5).half().float()
    b1 = torch.randn(mid, generator=g) * 0.5
    wd = torch.randn(mid, 1, k, k, generator=g) / k
    bd = torch.randn(mid, generator=g) * 0.5
    w2 = (torch.randn(c_, mid, generator=g) / mid ** 0.5).half().float()
    b2 = torch.randn(c_, generator=g)
    assert ops.bottleneck_supported(c_, mid, c_, k)
    t1 = F.silu(F.conv2d(x, w1[:, :, None, None], b1)).half().float()
    t2 = F.silu(F.conv2d(t1, wd, bd, padding=k // 2, groups=mid)).half().float()
    ref = F.silu(F.conv2d(t2, w2[:, :, None, None], b2))
    ld = (3 * c_ + 15) // 16 * 16
    cat = ops.NHWC(torch.full((n, h, w, ld), 7.0, dtype=torch.float16, device=cuda_device), 0, 3 * c_)
    cat.buf[..., c_:2 * c_] = x.permute(0, 2, 3, 1).to(cuda_device).half()
    packed = ops.pack_bottleneck(w1, b1, wd, bd, w2, b2, device=cuda_device)
    ops.bottleneck(cat.slice(c_, c_), packed, cat.slice(2 * c_, c_))
    torch.cuda.synchronize()
    _close(cat.slice(2 * c_, c_).to_nchw(), ref, f"bottleneck c_={c_} k={k} {h}x{w}", rtol=4e-3, atol=4e-3)
    assert (cat.buf[..., :c_] == 7.0).all() and (cat.buf[..., 3 * c_:] == 7.0).all(), "bottleneck wrote outside its slice"
    assert torch.equal(cat.buf[..., c_:2 * c_].cpu(), x.permute(0, 2, 3, 1).half()), "bottleneck modified its input"
    # twice in a row on the same stream (barrier phases / TMEM re-allocation across launches) and bit-identical
    first = cat.slice(2 * c_, c_).to_nchw().clone()
    ops.bottleneck(cat.slice(c_, c_), packed, cat.slice(2 * c_, c_))
    torch.cuda.synchronize()
    assert torch.equal(cat.slice(2 * c_, c_).to_nchw(), first)


@pytest.mark.parametrize("c,cout,h,w,n,act,wide", [(48, 48, 40, 40, 2, "silu", True), (96, 96, 20, 24, 3, "silu", True),
                                                   (256, 128, 16, 16, 2, "silu", False), (48, 24, 10, 14, 1, "none", False),
                                                   (64, 48, 160, 160, 2, "silu", True), (192, 96, 40, 40, 1, "relu", True)])
def test_maxpool2x2_conv1x1_fused(cuda_device, c, cout, h, w, n, act, wide):
    """Fused 2x2 max pool -> 1x1 kernel against the two reference ops in fp32 (max of fp16 values is exact).  `wide`:
    the output is the first half of a 2*cout-wide buffer (MPRep's concat); the other half must stay untouched."""
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(91 + c + cout)
    x = torch.randn(n, c, h, w, generator=g).half().float()
    pw_w = (torch.randn(cout, c, generator=g) / c ** 0.5).half().float()
    pw_b = torch.randn(cout, generator=g)
    ref = _act_ref(F.conv2d(F.max_pool2d(x, 2, 2), pw_w[:, :, None, None], pw_b), act)
    src = ops.NHWC.from_nchw(x.to(cuda_device))
    wp, bp = ops.pack_conv1x1(pw_w, pw_b, [c], cuda_device)
    full = ops.NHWC.empty(n, h // 2, w // 2, 2 * cout if wide else cout, cuda_device,
                          ld=((2 * cout if wide else cout) + 15) // 16 * 16)
    full.buf.fill_(7.0)
    dst = full.slice(0, cout)
    ops.maxpool2x2_conv1x1(src, wp, bp, act, dst)
    torch.cuda.synchronize()
    _close(dst.to_nchw(), ref, f"maxpool2x2_conv1x1 c={c}->{cout}", rtol=3e-3, atol=3e-3)
    assert (full.buf[..., cout:] == 7.0).all(), "maxpool2x2_conv1x1 wrote outside its channels"


def test_pool_upsample_layout(cuda_device):
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(3)
    n, c, h, w = 2, 48, 20, 24
    x = torch.randn(n, c, h, w, generator=g).half().float()
    src = ops.NHWC.from_nchw(x.to(cuda_device))
    # 2x2 max pool: exact
    dst = ops.NHWC.empty(n, h // 2, w // 2, c, cuda_device)
    ops.maxpool2x2(src, dst)
    assert torch.equal(dst.to_nchw().cpu(), F.max_pool2d(x, 2, 2))
    # SPPF chain: exact, written into channel slices of one concat buffer
    cat = ops.NHWC.empty(n, h, w, 4 * c, cuda_device)
    ops.sppf_pool(src, cat.slice(c, c), cat.slice(2 * c, c), cat.slice(3 * c, c))
    y1 = F.max_pool2d(x, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    y3 = F.max_pool2d(y2, 5, 1, 2)
    assert torch.equal(cat.slice(c, c).to_nchw().cpu(), y1)
    assert torch.equal(cat.slice(2 * c, c).to_nchw().cpu(), y2)
    assert torch.equal(cat.slice(3 * c, c).to_nchw().cpu(), y3)
    # nearest upsample: exact
    up = ops.NHWC.empty(n, 2 * h, 2 * w, c, cuda_device)
    ops.upsample2x(src, up)
    assert torch.equal(up.to_nchw().cpu(), F.interpolate(x, scale_factor=2, mode="nearest"))
    # layout converters round-trip
    xin = torch.randn(n, 37, h, w, generator=g)
    d2 = ops.NHWC.empty(n, h, w, 37, cuda_device)
    ops.nchw_to_nhwc(xin.to(cuda_device), d2)
    assert torch.equal(d2.to_nchw().cpu(), xin.half().float())
    back = torch.empty(n, 37, h, w, device=cuda_device)
    ops.nhwc_to_nchw(d2, back)
    torch.cuda.synchronize()
    assert torch.equal(back.cpu(), xin.half().float())


def _decode_ref(cls_logits, regs, strides, reg_max=16):
    """Plain torch fp32 restatement of Detect_yaml eval branch (yolo.py:355-396) + the head sigmoid."""
    outs_c, outs_r, pts, sts = [], [], [], []
    for cl, rg, s in zip(cls_logits, regs, strides):
        b, _, h, w = cl.shape
        l = h * w
        r = rg.reshape(b, 4, reg_max + 1, l).permute(0, 2, 1, 3)
        r = (F.softmax(r, dim=1) * torch.arange(reg_max + 1.0).view(1, -1, 1, 1)).sum(1)
        outs_c.append(torch.sigmoid(cl).reshape(b, -1, l))
        outs_r.append(r.reshape(b, 4, l))
        sy, sx = torch.meshgrid(torch.arange(h) + 0.5, torch.arange(w) + 0.5, indexing="ij")
        pts.append(torch.stack([sx, sy], -1).reshape(-1, 2))
        sts.append(torch.full((l, 1), float(s)))
    c = torch.cat(outs_c, -1).permute(0, 2, 1)
    r = torch.cat(outs_r, -1).permute(0, 2, 1)
    pts, sts = torch.cat(pts), torch.cat(sts)
    x1y1, x2y2 = pts - r[..., :2], pts + r[..., 2:]
    box = torch.cat([(x1y1 + x2y2) / 2, x2y2 - x1y1], -1) * sts
    return torch.cat([box, torch.ones(c.shape[0], c.shape[1], 1), c], -1)


def test_head_decode(cuda_device):
    from maf_yolo_b200 import ops

    g = torch.Generator().manual_seed(17)
    n, nc = 2, 80
    sizes, strides = [(16, 24), (8, 12), (5, 6)], [8.0, 16.0, 32.0]
    cls = [(torch.randn(n, nc, h, w, generator=g) * 2 - 3).half().float() for h, w in sizes]
    reg = [(torch.randn(n, 68, h, w, generator=g) * 2).half().float() for h, w in sizes]
    ref = _decode_ref(cls, reg, strides)
    pred = torch.empty(ref.shape, device=cuda_device)
    ops.head_decode([ops.NHWC.from_nchw(t.to(cuda_device)) for t in cls],
                    [ops.NHWC.from_nchw(t.to(cuda_device)) for t in reg], strides, 16, pred)
    torch.cuda.synchronize()
    # fp32 in / fp32 out: boxes within 1e-3 relative to max(1,|ref|), scores within 1e-5
    _close(pred[..., :4], ref[..., :4], "decode boxes", rtol=1e-5, atol=1e-3)
    assert (pred[..., 4] == 1).all()
    _close(pred[..., 5:], ref[..., 5:], "decode scores", rtol=0, atol=1e-6)


@pytest.mark.parametrize("c_in,sizes", [(128, [(16, 24), (8, 12), (5, 6)]), (192, [(40, 40), (20, 20), (10, 10)])])
def test_head_pred_epilogues(cuda_device, c_in, sizes):
    """K7 (mafb200_head_pred): cls_pred / reg_pred 1x1 conv + sigmoid / DFL decode in the GEMM epilogue against fp32
    torch (conv on the fp16-rounded operands, then the reference decode), and its detect mode — boxes + candidate keys
    -> mafb200_nms_select — against mafb200_nms on the pred tensor the same kernels wrote: bit for bit."""
    from maf_yolo_b200 import nn as mnn, ops

    g = torch.Generator().manual_seed(23)
    n, nc, strides = 2, 80, [8.0, 16.0, 32.0]
    feats = [(torch.randn(n, c_in, h, w, generator=g)).half().float() for h, w in sizes]
    wc = [(torch.randn(nc, c_in, generator=g) * 0.25).half().float() for _ in sizes]
    bc = [torch.randn(nc, generator=g) * 0.5 - 3.0 for _ in sizes]
    wr = [(torch.randn(68, c_in, generator=g) * 0.2).half().float() for _ in sizes]
    br = [torch.randn(68, generator=g) for _ in sizes]
    cls = [torch.einsum("nchw,oc->nohw", f, w_) + b_.view(1, -1, 1, 1) for f, w_, b_ in zip(feats, wc, bc)]
    reg = [torch.einsum("nchw,oc->nohw", f, w_) + b_.view(1, -1, 1, 1) for f, w_, b_ in zip(feats, wr, br)]
    ref = _decode_ref(cls, reg, strides)
    total = sum(h * w for h, w in sizes)
    pred = torch.full((n, total, 5 + nc), float("nan"), device=cuda_device)
    boxes = torch.full((n, total, 4), float("nan"), device=cuda_device)
    ws = torch.empty((ops.nms_workspace_bytes(n, total, nc) + 7) // 8, dtype=torch.int64, device=cuda_device)
    packed, off = [], 0
    for (h, w), f, w_c, b_c, w_r, b_r, st in zip(sizes, feats, wc, bc, wr, br, strides):
        src = ops.NHWC.from_nchw(f.to(cuda_device))
        pc = ops.pack_conv1x1(w_c, b_c, [c_in], device=cuda_device)
        pr = ops.pack_head_reg(w_r, b_r, device=cuda_device)
        packed.append((src, pc, pr, off, st))
        ops.head_pred(src, *pc, "cls", off, total, st, nc, pred=pred)
        ops.head_pred(src, *pr, "reg", off, total, st, nc, pred=pred)
        off += h * w
    torch.cuda.synchronize()
    assert not torch.isnan(pred).any(), "every element of pred must be written by the six launches"
    _close(pred[..., :4], ref[..., :4], "K7 boxes", rtol=1e-5, atol=2e-3)
    assert (pred[..., 4] == 1).all()
    _close(pred[..., 5:], ref[..., 5:], "K7 scores", rtol=0, atol=2e-5)
    for kw in (dict(conf_thres=0.03, multi_label=True), dict(conf_thres=0.05, multi_label=False),
               dict(conf_thres=0.03, multi_label=True, classes=[1, 5, 7, 40])):
        want_d, want_c = mnn.non_max_suppression_padded(pred, iou_thres=0.65, **kw)
        cfg = ops.detect_cfg_host(kw["conf_thres"], kw["multi_label"], nc, kw.get("classes")).to(cuda_device)
        ops.detect_reset(ws, n)
        for src, pc, pr, off, st in packed:
            ops.head_pred(src, *pc, "cls", off, total, st, nc, detect_cfg=cfg, workspace=ws)
            ops.head_pred(src, *pr, "reg", off, total, st, nc, boxes=boxes)
        det = torch.empty_like(want_d)
        cnt = torch.empty_like(want_c)
        ops.nms_select(boxes, nc, 0.65, False, 300, 30000, det, cnt, ws)
        torch.cuda.synchronize()
        assert int(want_c.sum()) > 0 and torch.equal(cnt, want_c) and torch.equal(det, want_d), kw
        # the same selection written as rows of the multi-GPU gather buffer (detections + count bits per row)
        rows = torch.full((n, 300 * 6 + 2), float("nan"), device=cuda_device)
        ops.nms_select_packed(boxes, nc, 0.65, False, 300, 30000, rows, ws)
        torch.cuda.synchronize()
        assert torch.equal(rows[:, :1800].reshape(n, 300, 6), want_d)
        assert torch.equal(rows.view(torch.int32)[:, 1800], want_c)
    assert torch.equal(boxes, pred[..., :4])


from tests._synthetic import synthetic_pred as _synthetic_pred  # noqa: E402


NMS_CASES = [
    dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300),                      # eval settings
    dict(conf_thres=0.4, iou_thres=0.45, multi_label=False, max_det=1000),                     # tools/infer.py defaults
    dict(conf_thres=0.01, iou_thres=0.45, multi_label=False, max_det=1000),
    dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, agnostic=True, classes=[1, 5, 7]),
    dict(conf_thres=0.001, iou_thres=0.65, multi_label=True, max_det=300),                     # > 8192 candidates: global sort path
    dict(conf_thres=0.001, iou_thres=0.65, multi_label=True, max_det=300, max_nms=3000),       # max_nms cut
    dict(conf_thres=0.9999, iou_thres=0.5, multi_label=True),                                  # no candidates at all
]


@pytest.mark.parametrize("kw", NMS_CASES)
def test_nms_bitexact(cuda_device, kw):
    from maf_yolo_b200 import nn as mnn
    from oracle import nms as onms

    pred = _synthetic_pred(3, 8400, 80, seed=1)
    ref = onms.non_max_suppression(pred.numpy(), **kw)
    kw2 = dict(kw)
    max_nms = kw2.pop("max_nms", None)
    got = mnn.non_max_suppression(pred.to(cuda_device), **kw2, **({"max_nms": max_nms} if max_nms else {}))
    assert len(got) == len(ref)
    for i, (gt, rf) in enumerate(zip(got, ref)):
        assert gt.shape == rf.shape, (i, gt.shape, rf.shape)
        assert np.array_equal(gt.cpu().numpy(), rf), f"image {i}: NMS output differs from the oracle"


def test_nms_many_candidates_prefix_select(cuda_device):
    """> 8192 candidates per image: the radix pre-selection of the best <= 8192 keys must give the same result as
    sorting everything — when the prefix suffices (max_det reached), when it does not (the best 8192 candidates
    collapse to a handful of survivors, so the kernel falls back to the full sort), with > max_nms candidates,
    and with massive exact score ties (selection impossible -> full sort)."""
    from maf_yolo_b200 import nn as mnn
    from oracle import nms as onms

    g = torch.Generator().manual_seed(21)
    A, nc = 8400, 80
    # (a) 40k candidates, random boxes: the prefix reaches max_det
    pa = _synthetic_pred(1, A, nc, seed=3)
    pa[..., 5:] = (pa[..., 5:] * 40).clamp(max=0.999)
    # (b) the 9000 best-scoring candidates are the same box (one survivor), the rest are spread out
    pb = _synthetic_pred(1, A, nc, seed=4, dup=False)
    pb[..., 5:] = pb[..., 5:] * 0.0
    pb[0, :, 5] = 0.05 + 0.1 * torch.rand(A, generator=g)             # everybody: class 0, low scores
    pb[0, :6000, :4] = torch.tensor([320.0, 320.0, 100.0, 100.0])     # 6000 anchors share one box ...
    pb[0, :6000, 5] = 0.5 + 0.4 * torch.rand(6000, generator=g)       # ... with the best class-0 scores
    pb[0, :6000, 6] = 0.5 + 0.4 * torch.rand(6000, generator=g)       # ... and class-1 scores: 12000 top keys, 2 survivors
    # (c) every candidate has exactly the same score
    pc = _synthetic_pred(1, A, nc, seed=5, dup=False)
    pc[..., 5:] = 0.0
    pc[0, :, 5:8] = 0.25
    for name, pred, kws in (("prefix", pa, [dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300),
                                            dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300, max_nms=20000),
                                            dict(conf_thres=0.001, iou_thres=0.3, multi_label=True, max_det=2000)]),
                            ("fallback", pb, [dict(conf_thres=0.03, iou_thres=0.5, multi_label=True, max_det=300)]),
                            ("ties", pc, [dict(conf_thres=0.03, iou_thres=0.5, multi_label=True, max_det=300)])):
        for kw in kws:
            n_cand = int(((pred[..., 5:] * pred[..., 4:5]) > kw["conf_thres"]).sum())
            assert n_cand > 8192, (name, n_cand)
            ref = onms.non_max_suppression(pred.numpy(), **kw)
            got = mnn.non_max_suppression(pred.to(cuda_device), **kw)
            assert np.array_equal(got[0].cpu().numpy(), ref[0]), f"{name} {kw}: differs from the oracle ({n_cand} candidates)"


@pytest.mark.parametrize("kw", [dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300),
                                dict(conf_thres=0.3, iou_thres=0.45, multi_label=False, max_det=1000),
                                dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, agnostic=True, classes=[1, 5, 7])])
def test_decode_detect_equals_decode_then_nms(cuda_device, kw):
    """mafb200_head_decode_detect + mafb200_nms_select (prediction tensor never written) give the same detections,
    bit for bit, as mafb200_head_decode followed by mafb200_nms — and as the oracle NMS on the decoded tensor."""
    from maf_yolo_b200 import nn as mnn, ops
    from oracle import nms as onms

    g = torch.Generator().manual_seed(23)
    n, nc = 3, 80
    sizes, strides = [(16, 24), (8, 12), (5, 6)], [8.0, 16.0, 32.0]
    cls = [(torch.randn(n, nc, h, w, generator=g) * 2.5 - 4).half().float() for h, w in sizes]
    reg = [(torch.randn(n, 68, h, w, generator=g) * 2).half().float() for h, w in sizes]
    cls_t = [ops.NHWC.from_nchw(t.to(cuda_device)) for t in cls]
    reg_t = [ops.NHWC.from_nchw(t.to(cuda_device)) for t in reg]
    a = sum(h * w for h, w in sizes)
    pred = torch.empty((n, a, 5 + nc), device=cuda_device)
    ops.head_decode(cls_t, reg_t, strides, 16, pred)
    kw2 = dict(kw)
    want_det, want_cnt = mnn.non_max_suppression_padded(pred, **kw2)
    ref = onms.non_max_suppression(pred.cpu().numpy(), **kw)
    boxes = torch.full((n, a, 4), float("nan"), device=cuda_device)
    ws = torch.empty((ops.nms_workspace_bytes(n, a, nc) + 7) // 8, dtype=torch.int64, device=cuda_device)
    filt = None
    if kw.get("classes") is not None:
        filt = torch.zeros(nc, dtype=torch.uint8)
        filt[kw["classes"]] = 1
        filt = filt.to(cuda_device)
    ops.head_decode_detect(cls_t, reg_t, strides, 16, boxes, kw["conf_thres"], kw["multi_label"], filt, ws)
    det = torch.full((n, kw.get("max_det", 300), 6), -1.0, device=cuda_device)
    cnt = torch.full((n,), -1, dtype=torch.int32, device=cuda_device)
    ops.nms_select(boxes, nc, kw["iou_thres"], kw.get("agnostic", False), kw.get("max_det", 300), 30000, det, cnt, ws)
    torch.cuda.synchronize()
    cand = pred[..., 5:].max(-1).values > kw["conf_thres"]  # boxes are only defined (and only read) for candidate rows
    assert torch.equal(boxes[cand], pred[..., :4][cand]) and int(cand.sum()) > 0
    assert torch.equal(cnt, want_cnt) and torch.equal(det, want_det)
    for i, r in enumerate(ref):
        assert np.array_equal(det[i, :int(cnt[i])].cpu().numpy(), r)
    assert int(cnt.sum()) > 0


def test_nms_obj_and_small(cuda_device):
    """objectness != 1, a batch where one image is empty, nc == 1 (multi_label is forced off)."""
    from maf_yolo_b200 import nn as mnn
    from oracle import nms as onms

    pred = _synthetic_pred(2, 500, 6, seed=9, dup=False)
    pred[..., 4] = torch.rand(2, 500, generator=torch.Generator().manual_seed(2))
    pred[..., 5:] = pred[..., 5:] * 50
    pred[1, :, 4] = 0.0  # second image: nothing passes
    for kw in (dict(conf_thres=0.05, iou_thres=0.5, multi_label=True), dict(conf_thres=0.05, iou_thres=0.5)):
        ref = onms.non_max_suppression(pred.numpy(), **kw)
        got = mnn.non_max_suppression(pred.to(cuda_device), **kw)
        for gt, rf in zip(got, ref):
            assert np.array_equal(gt.cpu().numpy(), rf)
    p1 = _synthetic_pred(1, 300, 1, seed=4, dup=False)
    p1[..., 5:] = p1[..., 5:] * 200
    ref = onms.non_max_suppression(p1.numpy(), conf_thres=0.05, iou_thres=0.5, multi_label=True)
    got = mnn.non_max_suppression(p1.to(cuda_device), conf_thres=0.05, iou_thres=0.5, multi_label=True)
    assert np.array_equal(got[0].cpu().numpy(), ref[0])


def test_error_codes(cuda_device):
    """The C ABI reports bad arguments through return codes + last_error, never by crashing."""
    from maf_yolo_b200 import _lib, ops

    a = ops.NHWC.empty(1, 8, 8, 16, cuda_device)
    b = ops.NHWC.empty(1, 4, 4, 16, cuda_device)
    w = torch.zeros(16, 64, dtype=torch.float16, device=cuda_device)
    bias = torch.zeros(16, device=cuda_device)
    with pytest.raises(_lib.MafError) as e:
        ops.conv1x1([a], w, bias, "silu", b)  # spatial mismatch
    assert e.value.code == -1 and "differ" in str(e.value)
    with pytest.raises(_lib.MafError):
        ops.dwconv(a, w, bias, 4, "silu", a)  # unsupported kernel size
    with pytest.raises(_lib.MafError):
        ops.conv3x3s2(a, w, bias, "silu", a)  # wrong output size
