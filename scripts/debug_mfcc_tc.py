"""GPU-side stage-by-stage check of the tensor-core MFCC kernel (development aid).
Dumps the raw stage-1 / stage-2 accumulators through mmla_debug_mfcc_tc_dump and compares them
and the final cepstra with float64 numpy.  Run on the GPU box:  python scripts/debug_mfcc_tc.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import _lib, synth
from mmla_audio_b200 import speaker_identification as si
from oracle import psf

lib = _lib.load()
n_clips, L, nfilt = 3, 40000, int(os.environ.get("NFILT", "26"))
pcm = synth.synth_clips(0, n_clips, L)
x = pcm.cpu().numpy()
cfg = si.MfccConfig(nfilt=nfilt)
T = cfg.num_frames(L)
gpc = (T + 15) // 16
n_groups = n_clips * gpc
n_tiles = (n_groups + 3) // 4
dbg = torch.full((n_tiles, 2, 128, 256), float("nan"), dtype=torch.float32, device="cuda")
lib.mmla_debug_mfcc_tc_dump(dbg.data_ptr(), None)
out = torch.empty((n_clips, T, 13), dtype=torch.float32, device="cuda")
si.mfcc_batch(pcm, cfg, out=out)
torch.cuda.synchronize()
lib.mmla_debug_mfcc_tc_dump(None, None)
d = dbg.cpu().numpy().astype(np.float64)
got = out.cpu().numpy()

# expected stage values
n1 = np.arange(32)
B1 = np.zeros((32, 32))
B1[0] = 1.0
B1[1] = (-1.0) ** n1
for j in range(1, 16):
    B1[2 * j] = np.cos(2 * np.pi * n1 * j / 32)
    B1[2 * j + 1] = -np.sin(2 * np.pi * n1 * j / 32)
B1[:, 25:] = 0
e1 = e2 = 0.0
s1max = s2max = 0.0
for tile in range(n_tiles):
    for g in range(4):
        G = tile * 4 + g
        if G >= n_groups:
            continue
        clip, f0 = G // gpc, (G % gpc) * 16
        y = psf.preemphasis(x[clip], 0.97) * 0.5
        ypad = np.concatenate([y, np.zeros(16 * 160 + 512 + 160 * 16)])
        ypad[len(y):] = 0
        for fl in range(16):
            t = f0 + fl
            fr = ypad[t * 160:t * 160 + 512].copy()
            fr[400:] = 0
            for h in range(2):
                for r in range(8):
                    n2 = 8 * h + r
                    exp = B1 @ fr[16 * n1 + n2]
                    if h == 1:          # the h = 1 copy of B1 carries W64^k1 (half of the twiddle)
                        for j in range(1, 16):
                            z = (exp[2 * j] + 1j * exp[2 * j + 1]) * np.exp(-2j * np.pi * j / 64)
                            exp[2 * j], exp[2 * j + 1] = z.real, z.imag
                    lane = 8 * fl + r
                    gotv = d[tile, 0, lane, (2 * g + h) * 32:(2 * g + h) * 32 + 32]
                    e1 = max(e1, np.abs(gotv - exp).max())
                    s1max = max(s1max, np.abs(exp).max())
            X = np.fft.fft(fr, 512) * (2.0 ** -5)
            for p in range(2):
                row = p * 64 + 16 * g + fl
                for j in range(8):
                    k1 = 2 * j + p
                    for k2 in range(16):
                        v = X[k1 + 32 * k2]
                        cre = 4 * (k2 >> 1) + (k2 & 1)
                        gr_, gi_ = d[tile, 1, row, j * 32 + cre], d[tile, 1, row, j * 32 + cre + 2]
                        e2 = max(e2, abs(gr_ - v.real), abs(gi_ - v.imag))
                        s2max = max(s2max, abs(v))
print(f"stage 1: max abs err {e1:.4g} (scale {s1max:.4g}, rel {e1 / max(s1max, 1e-30):.3g})")
print(f"stage 2: max abs err {e2:.4g} (scale {s2max:.4g}, rel {e2 / max(s2max, 1e-30):.3g})")
worst = 0.0
for c in range(n_clips):
    ref = psf.mfcc(x[c], 16000, 0.025, 0.01, 13, nfilt, 512)
    tol = 1e-4 * np.abs(ref) + 1e-4 * np.abs(ref).max()
    err = np.abs(got[c] - ref) / tol
    worst = max(worst, float(np.nanmax(err)))
    if not np.isfinite(got[c]).all():
        print("clip", c, "non-finite outputs:", int((~np.isfinite(got[c])).sum()))
print(f"mfcc: worst |err|/tol = {worst:.4g}")
print("first row got", got[0, 0, :5], "ref", psf.mfcc(x[0], 16000, 0.025, 0.01, 13, nfilt, 512)[0, :5])
