"""GPU parity of conv_slab_kernel (tcgen05 conv with filter taps as operand row shifts, csrc/conv_slab.cu), layer by
layer through the C-ABI's mmla_debug_conv2d: every stride-1 3x3 / 4x1 layer shape of the overlap classifier
(overlap_detector_temp.py:253-280) plus ragged geometries, against
  * a torch fp32 convolution with explicit Keras 'same' padding (1 before / 2 after for k = 4) — TF32 operands, so the
    bar is 3e-3 of the layer's largest output, and
  * conv_tc_kernel (same TF32 rounding of both operands, other accumulation order): 2e-5 of the largest output.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _conv(torch, lib, x, w, bias, bn, act, res, kernel):
    B, H, W, Cin = x.shape
    kh, kw, _, N = w.shape
    y = torch.empty(B, H, W, N, device="cuda", dtype=torch.float32)
    wh = np.ascontiguousarray(w.reshape(kh * kw * Cin, N), dtype=np.float32)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    rc = lib.mmla_debug_conv2d(ptr(x), wh.ctypes.data_as(C.c_void_p), ptr(bias), ptr(bn[0]) if bn else None,
                               ptr(bn[1]) if bn else None, act, ptr(res), ptr(y), B, H, W, Cin, N, kh, kw, kernel,
                               C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.mmla_last_error().decode()
    return y


def _torch_ref(torch, x, w, bias, bn, act, res):
    F = torch.nn.functional
    kh, kw = w.shape[0], w.shape[1]
    a = x.double()
    if bn:
        a = a * bn[0].double() + bn[1].double()
        a = torch.where(a > 0, a, torch.expm1(a)) if act == 2 else (a.clamp_min(0) if act == 1 else a)
    a = a.permute(0, 3, 1, 2)
    pt, pl = (kh - 1) // 2, (kw - 1) // 2
    a = F.pad(a, (pl, kw - 1 - pl, pt, kh - 1 - pt))
    y = F.conv2d(a, torch.as_tensor(w, device=x.device).double().permute(3, 2, 0, 1)) + bias.double()[None, :, None, None]
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + res.double()
    return y.float()


LAYERS = [  # (H, W, Cin, N, kh, kw, act, with_res, B)
    (128, 151, 16, 32, 3, 3, 2, False, 2),
    (128, 151, 32, 32, 4, 1, 2, False, 2),
    (64, 76, 32, 32, 3, 3, 2, False, 3),
    (64, 76, 32, 32, 4, 1, 2, True, 3),
    (64, 76, 32, 64, 3, 3, 2, False, 2),
    (64, 76, 64, 64, 4, 1, 2, False, 2),
    (32, 38, 64, 64, 3, 3, 2, False, 5),
    (32, 38, 64, 64, 4, 1, 2, True, 5),
    (32, 38, 64, 128, 3, 3, 2, False, 3),
    (32, 38, 128, 128, 4, 1, 2, False, 3),
    (16, 19, 128, 128, 3, 3, 2, False, 7),
    (16, 19, 128, 128, 4, 1, 2, True, 7),
    (5, 7, 32, 32, 3, 3, 1, True, 3),          # less than one tile per image
    (9, 130, 16, 64, 3, 3, 0, False, 2),        # no prologue
    (37, 3, 64, 32, 4, 1, 1, False, 4),
    (2, 2, 128, 128, 2, 2, 2, True, 1),
]


@pytest.mark.parametrize("H,W,Cin,N,kh,kw,act,with_res,B", LAYERS)
def test_slab_conv_layer(cuda, H, W, Cin, N, kh, kw, act, with_res, B):
    torch = cuda
    from mmla_audio_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(H * 1000 + W * 10 + Cin + N + kh)
    x = torch.randn(B, H, W, Cin, generator=g).cuda()
    w = (torch.randn(kh, kw, Cin, N, generator=g) * (2.0 / (kh * kw * Cin)) ** 0.5).numpy()
    bias = (torch.randn(N, generator=g) * 0.1).cuda()
    use_bn = act != 0
    bn = ((torch.rand(Cin, generator=g) + 0.5).cuda(), (torch.randn(Cin, generator=g) * 0.3).cuda()) if use_bn else None
    res = torch.randn(B, H, W, N, generator=g).cuda() if with_res else None
    ref = _torch_ref(torch, x, w, bias, bn, act, res)
    slab = _conv(torch, lib, x, w, bias, bn, act, res, 2)
    gather = _conv(torch, lib, x, w, bias, bn, act, res, 1)
    scale = ref.abs().max().item()
    d_ref = (slab - ref).abs().max().item()
    d_tc = (slab - gather).abs().max().item()
    print(f"slab vs fp64 conv {d_ref / scale:.2e}, vs gather kernel {d_tc / scale:.2e} (of max |y| = {scale:.2f})")
    assert d_ref <= 3e-3 * scale
    assert d_tc <= 2e-5 * scale


def test_slab_switch_keeps_overlap_net_output(cuda, monkeypatch):
    """Whole overlap net: MMLA_CONV_SLAB=0 (gather kernel for every conv) vs the default."""
    from mmla_audio_b200 import models, weights as W
    torch = cuda
    spec = W.OVERLAP
    model = models.Model(spec, W.synthetic_weights(spec, 1234), precision="tf32")
    x = torch.randint(0, 256, (5, 128, 151, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    monkeypatch.setenv("MMLA_CONV_SLAB", "0")
    p0, l0 = model.predict_device(x)
    monkeypatch.setenv("MMLA_CONV_SLAB", "1")
    p1, l1 = model.predict_device(x)
    d = (p0 - p1).abs().max().item()
    print("slab vs gather, overlap net: max |dprob|", d)
    assert d <= 2e-4
