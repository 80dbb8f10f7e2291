"""GPU parity of resblock2d_fused_kernel (csrc/resblock2d_fused.cu): both convolutions of a `res_block`
(overlap_detector_temp.py:253-280: BN -> ELU -> Conv2D(C, 3) -> BN -> ELU -> Conv2D(C, (4, 1)) [+ x]) in ONE launch with the
intermediate in shared memory, through the C-ABI's mmla_debug_resblock2d, against
  * the two conv_slab_kernel launches it replaces (mmla_debug_conv2d, kernel 2): same TF32 operands, same K order, same fp32
    epilogue expressions => BIT-IDENTICAL, and
  * a torch float64 pair of convolutions with explicit Keras 'same' padding (k = 4: one row before, two after): TF32 operands
    twice, so the bar is 5e-3 of the block's largest output.
Every block shape of the overlap classifier plus ragged geometries (an image smaller than one tile, one column, H = 2).
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ELU = 2


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _slab_conv(torch, lib, x, w, bias, bn, res):
    B, H, W, Cin = x.shape
    kh, kw, _, N = w.shape
    y = torch.empty(B, H, W, N, device="cuda", dtype=torch.float32)
    wh = np.ascontiguousarray(w.reshape(kh * kw * Cin, N), dtype=np.float32)
    rc = lib.mmla_debug_conv2d(_ptr(x), wh.ctypes.data_as(C.c_void_p), _ptr(bias), _ptr(bn[0]), _ptr(bn[1]), ELU, _ptr(res), _ptr(y),
                               B, H, W, Cin, N, kh, kw, 2, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.mmla_last_error().decode()
    return y


def _fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res, hpool=0):
    B, H, W, Cin = x.shape
    Cc = w1.shape[3]
    y = torch.full((B, H // 2 if hpool else H, W, Cc), float("nan"), device="cuda", dtype=torch.float32)
    w1h = np.ascontiguousarray(w1.reshape(9 * Cin, Cc), dtype=np.float32)
    w2h = np.ascontiguousarray(w2.reshape(4 * Cc, Cc), dtype=np.float32)
    rc = lib.mmla_debug_resblock2d(_ptr(x), w1h.ctypes.data_as(C.c_void_p), _ptr(b1), _ptr(bn1[0]), _ptr(bn1[1]),
                                   w2h.ctypes.data_as(C.c_void_p), _ptr(b2), _ptr(bn2[0]), _ptr(bn2[1]), _ptr(res), _ptr(y),
                                   B, H, W, Cin, Cc, hpool, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.mmla_last_error().decode()
    return y


def _torch_ref(torch, x, w1, b1, bn1, w2, b2, bn2, res):
    F = torch.nn.functional

    def conv(a, w, bias, bn):
        kh, kw = w.shape[0], w.shape[1]
        a = a * bn[0].double() + bn[1].double()
        a = torch.where(a > 0, a, torch.expm1(a)).permute(0, 3, 1, 2)
        pt, pl = (kh - 1) // 2, (kw - 1) // 2
        a = F.pad(a, (pl, kw - 1 - pl, pt, kh - 1 - pt))
        y = F.conv2d(a, torch.as_tensor(w, device=a.device).double().permute(3, 2, 0, 1)) + bias.double()[None, :, None, None]
        return y.permute(0, 2, 3, 1)

    y = conv(conv(x.double(), w1, b1, bn1), w2, b2, bn2)
    if res is not None:
        y = y + res.double()
    return y.float()


BLOCKS = [  # (H, W, Cin, C, residual, B)
    (128, 151, 16, 32, False, 2),      # block 1 (pooled: no residual here, MaxPool + shortcut follow)
    (64, 76, 32, 32, True, 3),         # blocks 2, 3
    (64, 76, 32, 64, False, 2),        # block 4
    (32, 38, 64, 64, True, 5),         # blocks 5, 6
    (32, 38, 64, 128, False, 3),       # block 7
    (16, 19, 128, 128, True, 7),       # blocks 8, 9
    (5, 7, 32, 32, True, 3),           # less than one tile per image
    (9, 130, 16, 64, False, 2),
    (37, 1, 64, 32, True, 4),          # one column
    (2, 3, 128, 128, True, 1),
    (128, 151, 16, 32, True, 1),       # Cin != C with a residual of C channels (generic entry)
    (2, 3, 64, 64, False, 2),          # smallest pooled block (H = 2): row-pooled epilogue on a one-tile image
    (6, 5, 16, 32, False, 3),          # tiny block-1 shape: persistent kernel with fewer work items than SMs
    (33, 9, 32, 32, False, 2),         # odd height: no row pooling, 36-row columns
]


@pytest.mark.parametrize("H,W,Cin,Cc,with_res,B", BLOCKS)
def test_fused_block_matches_two_slab_convs_bitwise(cuda, monkeypatch, H, W, Cin, Cc, with_res, B):
    torch = cuda
    monkeypatch.setenv("MMLA_NET_PERSIST", "0")          # resblock2d_fused_kernel; the persistent kernel is compared below
    from mmla_audio_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(H * 1000 + W * 10 + Cin + Cc)
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w1 = (torch.randn(3, 3, Cin, Cc, generator=g) * (2.0 / (9 * Cin)) ** 0.5).numpy()
    w2 = (torch.randn(4, 1, Cc, Cc, generator=g) * (2.0 / (4 * Cc)) ** 0.5).numpy()
    b1 = (torch.randn(Cc, generator=g) * 0.1).cuda()
    b2 = (torch.randn(Cc, generator=g) * 0.1).cuda()
    bn1 = ((torch.rand(Cin, generator=g) + 0.5).cuda(), (torch.randn(Cin, generator=g) * 0.3).cuda())
    bn2 = ((torch.rand(Cc, generator=g) + 0.5).cuda(), (torch.randn(Cc, generator=g) * 0.3).cuda())
    res = torch.randn(B, H, W, Cc, generator=g).cuda() if with_res else None
    fused = _fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res)
    u = _slab_conv(torch, lib, x, w1, b1, bn1, None)
    two = _slab_conv(torch, lib, u, w2, b2, bn2, res)
    ref = _torch_ref(torch, x, w1, b1, bn1, w2, b2, bn2, res)
    assert not torch.isnan(fused).any(), "an output pixel was never written"
    scale = ref.abs().max().item()
    d_ref = (fused - ref).abs().max().item()
    n_diff = int((fused != two).sum().item())
    print(f"fused vs fp64 convs {d_ref / scale:.2e} of max |y| = {scale:.2f}; elements differing from the two slab launches: {n_diff}")
    assert n_diff == 0
    assert d_ref <= 5e-3 * scale
    hp = None
    if not with_res and H % 2 == 0:
        # pooled blocks: the maximum over the row pairs (2i, 2i + 1) taken in the kernel's epilogue (HPOOL: column pitch H + 4,
        # 128 T - 4 outputs per CTA) against the same maximum of the full-resolution output
        hp = _fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, None, hpool=1)
        assert torch.equal(hp, torch.maximum(fused[:, 0::2], fused[:, 1::2]))
    if Cc == 32 and Cin in (16, 32):
        # resblock2d_persist_kernel (persistent, warp-specialised, weights resident): the same bits, also with a forced
        # one-tile configuration (more work items per CTA, every buffer parity exercised)
        monkeypatch.setenv("MMLA_NET_PERSIST", "2")
        for tiles in (None, "1"):
            if tiles:
                monkeypatch.setenv("MMLA_RB_TILES", tiles)
            ps = _fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res)
            assert torch.equal(ps, fused), f"persistent kernel differs (tiles {tiles})"
            if hp is not None:
                assert torch.equal(_fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, None, hpool=1), hp)
        monkeypatch.delenv("MMLA_RB_TILES", raising=False)


@pytest.mark.parametrize("H,W,Cin,Cc,with_res,B", BLOCKS)
def test_fp16_operand_block_matches_fp64_convs(cuda, monkeypatch, H, W, Cin, Cc, with_res, B):
    """The fp16-operand form of the kernel (MMLA_PRECISION_F16; here through MMLA_RB_F16=1 of the debug entry): fp16 keeps the 11
    significant bits TF32 keeps, so the bar against the float64 pair of convolutions is the TF32 one (5e-3 of max |y|) and the
    two tensor-core forms agree to 3e-3; the row-pooled variant must equal the maximum over row pairs of the full output."""
    torch = cuda
    from mmla_audio_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(H * 1000 + W * 10 + Cin + Cc + 7)
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w1 = (torch.randn(3, 3, Cin, Cc, generator=g) * (2.0 / (9 * Cin)) ** 0.5).numpy()
    w2 = (torch.randn(4, 1, Cc, Cc, generator=g) * (2.0 / (4 * Cc)) ** 0.5).numpy()
    b1 = (torch.randn(Cc, generator=g) * 0.1).cuda()
    b2 = (torch.randn(Cc, generator=g) * 0.1).cuda()
    bn1 = ((torch.rand(Cin, generator=g) + 0.5).cuda(), (torch.randn(Cin, generator=g) * 0.3).cuda())
    bn2 = ((torch.rand(Cc, generator=g) + 0.5).cuda(), (torch.randn(Cc, generator=g) * 0.3).cuda())
    res = torch.randn(B, H, W, Cc, generator=g).cuda() if with_res else None
    monkeypatch.setenv("MMLA_NET_PERSIST", "0")
    tf32 = _fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res)
    monkeypatch.setenv("MMLA_RB_F16", "1")
    f16 = _fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res)
    ref = _torch_ref(torch, x, w1, b1, bn1, w2, b2, bn2, res)
    assert not torch.isnan(f16).any(), "an output pixel was never written"
    scale = ref.abs().max().item()
    d_ref, d_tf = (f16 - ref).abs().max().item(), (f16 - tf32).abs().max().item()
    print(f"fp16 operands vs fp64 convs {d_ref / scale:.2e}, vs the TF32 form {d_tf / scale:.2e} of max |y| = {scale:.2f}")
    assert d_ref <= 5e-3 * scale and d_tf <= 3e-3 * scale
    if not with_res and H % 2 == 0:
        hp = _fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, None, hpool=1)
        assert torch.equal(hp, torch.maximum(f16[:, 0::2], f16[:, 1::2]))
    for tiles in ("1", "2"):                      # other tile counts per CTA: same arithmetic per element
        monkeypatch.setenv("MMLA_RB_TILES", tiles)
        assert torch.equal(_fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res), f16), f"tiles {tiles}"


def test_fp16_operands_saturate_instead_of_overflowing(cuda, monkeypatch):
    """Activations beyond the fp16 range (> 65504 after BN + ELU) are clamped to the largest finite half: the output stays
    finite (an unclamped conversion would give inf, and inf * 0-weight NaN)."""
    torch = cuda
    from mmla_audio_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(5)
    B, H, W, Cin, Cc = 1, 8, 8, 32, 32
    x = torch.randn(B, H, W, Cin, generator=g).cuda() * 1e6
    w1 = (torch.randn(3, 3, Cin, Cc, generator=g) * 1e-3).numpy()
    w2 = (torch.randn(4, 1, Cc, Cc, generator=g) * 1e-3).numpy()
    z = torch.zeros(Cc).cuda()
    one_in, one = (torch.ones(Cin).cuda(), torch.zeros(Cin).cuda()), (torch.ones(Cc).cuda(), torch.zeros(Cc).cuda())
    monkeypatch.setenv("MMLA_NET_PERSIST", "0")
    monkeypatch.setenv("MMLA_RB_F16", "1")
    y = _fused(torch, lib, x, w1, z, one_in, w2, z, one, None)
    assert torch.isfinite(y).all()


def test_fp16_mode_activated_operand_handoff_is_bitwise_neutral(cuda, monkeypatch):
    """Whole overlap net, precision "fp16": by default the kernel that produces a block's input (pool_shortcut_kernel or the
    previous conv-pair kernel's epilogue) also writes ELU(BN1(x)) as fp16, and the conv-pair kernel's fill is a plain
    asynchronous copy; MMLA_NET_F16_ACT=0 makes every fill convert from the fp32 tensor instead.  Same expression, same rounding:
    identical probabilities.  Against the TF32 mode the probabilities agree to 2e-3."""
    from mmla_audio_b200 import _lib, models, weights as W
    torch = cuda
    model = models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision="fp16")
    x8 = torch.randint(0, 256, (6, 128, 151, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).cuda()
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MMLA_NET_F16_ACT", mode)
        tr = _lib.trace_launches(lambda: out.__setitem__(mode, model.predict_device(x8)[0].clone()), torch)
        names = [n for n, _ in tr]
        assert names.count("stem_resblock2d_persist_f16_kernel") == 1 and names.count("resblock2d_f16_kernel") == 8, names
    monkeypatch.delenv("MMLA_NET_F16_ACT")
    assert torch.equal(out["0"], out["1"])
    # block 1 on the one-CTA-per-item kernel instead of the persistent one: the same bits
    monkeypatch.setenv("MMLA_NET_PERSIST", "0")
    tr = _lib.trace_launches(lambda: out.__setitem__("np", model.predict_device(x8)[0].clone()), torch)
    assert [n for n, _ in tr].count("stem_resblock2d_f16_kernel") == 1 and [n for n, _ in tr].count("resblock2d_f16_kernel") == 8
    # ... and blocks 2, 3 on the persistent kernel's fp16 form (fill = the producer's fp16 operand, epilogue 2 writes the next one)
    monkeypatch.setenv("MMLA_NET_PERSIST", "2")
    tr = _lib.trace_launches(lambda: out.__setitem__("p2", model.predict_device(x8)[0].clone()), torch)
    assert [n for n, _ in tr].count("resblock2d_persist_f16_kernel") == 2
    monkeypatch.delenv("MMLA_NET_PERSIST")
    assert torch.equal(out["np"], out["1"]) and torch.equal(out["p2"], out["1"])
    # opt-in MMLA_NET_F16_Z=1: the pooled blocks' row-pooled conv output travels to pool_shortcut_kernel as fp16: not bitwise
    # neutral — one more 11-bit rounding of a value that is already two 11-bit-operand convolutions deep — but small
    monkeypatch.setenv("MMLA_NET_F16_Z", "1")
    z16 = model.predict_device(x8)[0].clone()
    monkeypatch.delenv("MMLA_NET_F16_Z")
    dz = (out["1"] - z16).abs().max().item()
    model.set_precision("tf32")
    ref = model.predict_device(x8)[0]
    d = (out["1"] - ref).abs().max().item()
    print(f"fp16-operand mode vs TF32 mode: max |dprob| {d:.2e}; fp16 vs fp32 pooled conv output: {dz:.2e}")
    assert d <= 2e-3 and dz <= 1e-3


def test_block_fusion_switches_keep_overlap_net_output_bitwise(cuda, monkeypatch):
    """Whole overlap net, TF32 mode, uint8 and float32 images: MMLA_NET_FUSE_BLOCKS=0 (two conv_slab launches per block) vs
    MMLA_NET_FUSE_STEM2D=0 (one resblock2d_fused_kernel launch per block, stem1x1_kernel on its own) vs the default (the stem
    Conv2D(16, 1x1) also computed inside the first block's conv-pair and pooling kernels): identical probabilities, and the
    launch traces show which kernels ran; MMLA_NET_FUSE_HPOOL=0 keeps the pooled blocks' conv output at full resolution
    (default: the row half of the MaxPool is taken in the conv-pair kernel's epilogue); MMLA_NET_PERSIST=0 keeps the C = 32
    blocks on the one-CTA-per-item kernel (default: block 1 on resblock2d_persist_kernel, csrc/resblock2d_persist.cu; 2: blocks
    2 and 3 as well)."""
    from mmla_audio_b200 import _lib, models, weights as W
    torch = cuda
    spec = W.OVERLAP
    model = models.Model(spec, W.synthetic_weights(spec, 1234), precision="tf32")
    x8 = torch.randint(0, 256, (5, 128, 151, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    for x in (x8, x8.float() * 0.37 - 20.0):
        out, names = {}, {}
        sw = ("MMLA_NET_FUSE_BLOCKS", "MMLA_NET_FUSE_STEM2D", "MMLA_NET_FUSE_HPOOL", "MMLA_NET_PERSIST")
        for mode, env in (("two", {"MMLA_NET_FUSE_BLOCKS": "0"}), ("one", {"MMLA_NET_FUSE_STEM2D": "0", "MMLA_NET_FUSE_HPOOL": "0"}),
                          ("nohp", {"MMLA_NET_FUSE_HPOOL": "0"}), ("nopersist", {"MMLA_NET_PERSIST": "0"}),
                          ("persist3", {"MMLA_NET_PERSIST": "2"}), ("stem", {})):
            for k in sw:
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            tr = _lib.trace_launches(lambda: out.__setitem__(mode, model.predict_device(x)), torch)
            names[mode] = [n for n, _ in tr]
        pairs = lambda n: n.count("resblock2d_fused_kernel") + n.count("resblock2d_persist_kernel")
        stems = lambda n: n.count("stem_resblock2d_fused_kernel") + n.count("stem_resblock2d_persist_kernel")
        assert names["two"].count("conv_slab_kernel") == 18 and pairs(names["two"]) + stems(names["two"]) == 0
        assert pairs(names["one"]) == 9 and "conv_slab_kernel" not in names["one"]
        assert names["one"].count("stem1x1_kernel") == 1 and names["two"].count("stem1x1_kernel") == 1
        assert pairs(names["stem"]) == 8 and stems(names["stem"]) == 1 and "stem1x1_kernel" not in names["stem"]
        # block 1 runs on the persistent, warp-specialised kernel (MMLA_NET_PERSIST=0: never, 2: blocks 2 and 3 as well)
        assert names["stem"].count("stem_resblock2d_persist_kernel") == 1 and names["stem"].count("resblock2d_persist_kernel") == 0
        assert names["persist3"].count("stem_resblock2d_persist_kernel") == 1 and names["persist3"].count("resblock2d_persist_kernel") == 2
        assert not any("persist" in n for n in names["nopersist"]) and names["nopersist"].count("resblock2d_fused_kernel") == 8
        for mode in ("one", "nohp", "nopersist", "persist3", "stem"):
            assert torch.equal(out["two"][0], out[mode][0]) and torch.equal(out["two"][1], out[mode][1]), mode


@pytest.mark.parametrize("H,W,Cin,Cc,with_res,B", [(64, 76, 32, 64, False, 3), (32, 38, 64, 64, True, 5), (16, 19, 128, 128, True, 7),
                                                   (32, 38, 64, 128, False, 2)])
def test_cta_pair_mode_is_bit_identical(cuda, monkeypatch, H, W, Cin, Cc, with_res, B):
    """MMLA_RB_PAIR=1 (opt-in): `tcgen05.mma.cta_group::2` — two CTAs of a cluster run every MMA together (M = 256), each
    holding half of every weight chunk; an odd CTA count pads the grid with a protocol-only CTA (B odd here).  Same operands,
    same K order => the same bits as one CTA per MMA, with and without the row-pooled epilogue."""
    torch = cuda
    from mmla_audio_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(7 + H + Cin + Cc)
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w1 = (torch.randn(3, 3, Cin, Cc, generator=g) * (2.0 / (9 * Cin)) ** 0.5).numpy()
    w2 = (torch.randn(4, 1, Cc, Cc, generator=g) * (2.0 / (4 * Cc)) ** 0.5).numpy()
    b1 = (torch.randn(Cc, generator=g) * 0.1).cuda()
    b2 = (torch.randn(Cc, generator=g) * 0.1).cuda()
    bn1 = ((torch.rand(Cin, generator=g) + 0.5).cuda(), (torch.randn(Cin, generator=g) * 0.3).cuda())
    bn2 = ((torch.rand(Cc, generator=g) + 0.5).cuda(), (torch.randn(Cc, generator=g) * 0.3).cuda())
    res = torch.randn(B, H, W, Cc, generator=g).cuda() if with_res else None
    modes = [0] if with_res else [0, 1]
    monkeypatch.delenv("MMLA_RB_PAIR", raising=False)
    one = [_fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res, hpool=hp) for hp in modes]
    monkeypatch.setenv("MMLA_RB_PAIR", "1")
    two = [_fused(torch, lib, x, w1, b1, bn1, w2, b2, bn2, res, hpool=hp) for hp in modes]
    for a, b in zip(one, two):
        assert not torch.isnan(b).any() and torch.equal(a, b)
