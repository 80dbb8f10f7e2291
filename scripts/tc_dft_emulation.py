"""Numerical emulation (numpy) of the two-stage tensor-core DFT used by csrc/mfcc_tc.cu:
512 = 32 x 16 Cooley-Tukey, operands split into fp16 hi + lo, three MMA passes per stage,
fp32 accumulation.  Checks the end-to-end MFCC error against the float64 oracle before any
CUDA is written.  Development aid only (imports oracle/)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import psf, synth

f16 = np.float16
f32 = np.float32


def split16(a):
    hi = a.astype(f16)
    lo = (a.astype(f32) - hi.astype(f32)).astype(f16)
    return hi, lo


def mm3(ah, al, bh, bl, passes=3):
    # fp16 products are exact in fp32; accumulate in fp32 (float64 here then round: optimistic by ~1 ulp)
    A_h, A_l = ah.astype(np.float64), al.astype(np.float64)
    B_h, B_l = bh.astype(np.float64), bl.astype(np.float64)
    r = A_h @ B_h
    if passes >= 2:
        r = r + A_l @ B_h
    if passes >= 3:
        r = r + A_h @ B_l
    return r.astype(f32)


def build_b1():
    n1 = np.arange(32)[:, None]
    B = np.zeros((32, 32))
    B[:, 0] = 1.0
    B[:, 1] = (-1.0) ** np.arange(32)
    for j in range(1, 16):
        B[:, 2 * j] = np.cos(2 * np.pi * n1[:, 0] * j / 32)
        B[:, 2 * j + 1] = -np.sin(2 * np.pi * n1[:, 0] * j / 32)
    B[25:, :] = 0.0
    return B


def build_b2():
    B = np.zeros((32, 32))
    for n2 in range(16):
        for k2 in range(16):
            th = 2 * np.pi * n2 * k2 / 16
            B[2 * n2, 2 * k2] = np.cos(th)
            B[2 * n2 + 1, 2 * k2] = np.sin(th)
            B[2 * n2, 2 * k2 + 1] = -np.sin(th)
            B[2 * n2 + 1, 2 * k2 + 1] = np.cos(th)
    return B


def tc_powspec(y, passes=3, s1=0.5, s2=2.0 ** -5):
    """y: float32 [T, 512] pre-emphasised frames zero padded (only n<400 non-zero in practice,
    but the kernel reads the following samples and relies on zero rows of B1)."""
    T = y.shape[0]
    b1h, b1l = split16(build_b1().astype(f32))
    b2h, b2l = split16(build_b2().astype(f32))
    ys = (y * f32(s1)).astype(f32)
    yh, yl = split16(ys)
    # stage 1: rows (f, n2), K = n1
    A1h = yh.reshape(T, 32, 16).transpose(0, 2, 1).reshape(T * 16, 32)
    A1l = yl.reshape(T, 32, 16).transpose(0, 2, 1).reshape(T * 16, 32)
    D1 = mm3(A1h, A1l, b1h, b1l, passes).reshape(T, 16, 32)          # [f][n2][col]
    # twiddle (fp32 CUDA cores)
    S = np.zeros((T, 16, 17), np.complex64)
    S[:, :, 0] = D1[:, :, 0]
    S[:, :, 16] = D1[:, :, 1]
    for j in range(1, 16):
        S[:, :, j] = D1[:, :, 2 * j] + 1j * D1[:, :, 2 * j + 1]
    n2 = np.arange(16)[:, None]
    k1 = np.arange(17)[None, :]
    tw = np.exp(-2j * np.pi * n2 * k1 / 512).astype(np.complex64)
    Tt = (S * tw[None] * f32(s2)).astype(np.complex64)                # [f][n2][k1]
    P = np.zeros((T, 257), f32)
    scale = f32(1.0 / (512.0 * s1 * s1 * s2 * s2))
    for k1i in range(17):
        A2 = np.empty((T, 32), f32)
        A2[:, 0::2] = Tt[:, :, k1i].real
        A2[:, 1::2] = Tt[:, :, k1i].imag
        a2h, a2l = split16(A2)
        D2 = mm3(a2h, a2l, b2h, b2l, passes)                          # [f][(k2,c)]
        re, im = D2[:, 0::2], D2[:, 1::2]
        pw = (re * re + im * im) * scale
        for k2 in range(16):
            k = k1i + 32 * k2
            b = k if k <= 256 else 512 - k
            P[:, b] = pw[:, k2]
    return P


def mfcc_from_pspec(P, nfilt):
    energy = P.astype(np.float64).sum(1)
    energy = np.where(energy == 0, psf.EPS, energy)
    fb = psf.get_filterbanks(nfilt, 512, 16000, 0, 8000)
    feat = P.astype(np.float64) @ fb.T
    feat = np.where(feat == 0, psf.EPS, feat)
    feat = psf.lifter(psf.dct2_ortho(np.log(feat), 13), 22)
    feat[:, 0] = np.log(energy)
    return feat


def main():
    worst = 0.0
    for nfilt in (26, 40):
        for passes in (3, 2):
            wr = 0.0
            pcm = synth.synth_clips(0, 12, 40000)
            extra = [np.zeros(40000, np.int16), np.full(40000, 1234, np.int16),
                     (np.where(np.arange(40000) % 100 < 50, 32767, -32768)).astype(np.int16),
                     (3000 * np.sin(2 * np.pi * 440 * np.arange(40000) / 16000)).astype(np.int16),
                     (np.random.default_rng(0).integers(-3, 4, 40000)).astype(np.int16)]
            clips = [pcm[i] for i in range(pcm.shape[0])] + extra
            for x in clips:
                ref = psf.mfcc(x, 16000, 0.025, 0.01, 13, nfilt, 512)
                y = psf.preemphasis(x, 0.97).astype(f32)
                T = ref.shape[0]
                ypad = np.concatenate([y, np.zeros((T - 1) * 160 + 512 - len(y), f32)])
                idx = np.arange(512)[None, :] + 160 * np.arange(T)[:, None]
                fr = ypad[idx]        # frames incl. the 112 following samples (zero rows of B1 kill them)
                P = tc_powspec(fr, passes)
                got = mfcc_from_pspec(P, nfilt)
                tol = 1e-4 * np.abs(ref) + 1e-4 * np.abs(ref).max()
                r = float((np.abs(got - ref) / tol).max())
                wr = max(wr, r)
            print(f"nfilt={nfilt} passes={passes}: worst |err|/tol = {wr:.4f}")
            if passes == 3:
                worst = max(worst, wr)
    print("OK" if worst < 1.0 else "FAIL")


if __name__ == "__main__":
    main()
