"""Oracle: python_speech_features 0.6 ``mfcc`` exactly as the reference calls it, plus the
reference's own ``delta`` / padding / chunking.  TEST INFRASTRUCTURE ONLY (see oracle/__init__).

PARITY UNPINNED — python_speech_features is an un-pinned, un-vendored dependency
(``setup.py:32-41``); this file restates its published algorithm (release 0.6, the only
release with the ``nfft`` keyword and a rectangular default window) and is anchored on the
reference call sites:

    mfcc(sig, rate, winlen=0.025, winstep=0.01, nfft=512)
        SpeakerIdentification/scripts/speaker_identification.py:89,285,341,386
        SpeakerIdentification/scripts/speaker_identification_post_processing.py:256

Everything is float64, at int16 scale (``scipy.io.wavfile.read`` output is NOT rescaled).
"""
from __future__ import annotations

import decimal
import math

import numpy as np

EPS = float(np.finfo(float).eps)


def round_half_up(number) -> int:
    return int(decimal.Decimal(number).quantize(decimal.Decimal("1"),
                                                rounding=decimal.ROUND_HALF_UP))


def preemphasis(signal, coeff=0.95):
    """psf.sigproc.preemphasis: y[0]=x[0]; y[n]=x[n]-coeff*x[n-1]."""
    signal = np.asarray(signal)
    return np.append(signal[0], signal[1:] - coeff * signal[:-1]).astype(np.float64)


def num_frames(slen: int, frame_len: int = 400, frame_step: int = 160) -> int:
    """psf.sigproc.framesig frame count: 1 if slen<=frame_len else 1+ceil((slen-len)/step)."""
    if slen <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * slen - frame_len) / frame_step))


def framesig(sig, frame_len, frame_step, winfunc=lambda x: np.ones((x,))):
    """psf.sigproc.framesig: zero-pad the tail so the last frame is full; multiply by window."""
    slen = len(sig)
    frame_len = int(round_half_up(frame_len))
    frame_step = int(round_half_up(frame_step))
    nf = num_frames(slen, frame_len, frame_step)
    padlen = int((nf - 1) * frame_step + frame_len)
    padsignal = np.concatenate((sig, np.zeros((padlen - slen,))))
    idx = np.arange(frame_len)[None, :] + frame_step * np.arange(nf)[:, None]
    return padsignal[idx] * winfunc(frame_len)


def powspec(frames, nfft):
    """psf.sigproc.powspec: 1/NFFT * |rfft(frame, NFFT)|^2 (frame zero-padded to NFFT)."""
    return 1.0 / nfft * np.square(np.absolute(np.fft.rfft(frames, nfft)))


def hz2mel(hz):
    return 2595 * np.log10(1 + hz / 700.0)


def mel2hz(mel):
    return 700 * (10 ** (mel / 2595.0) - 1)


def get_filterbanks(nfilt=20, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    """psf.base.get_filterbanks: HTK-mel triangles on floor((nfft+1)*hz/sr) bin edges,
    no area normalisation.  Returns [nfilt, nfft//2+1] float64."""
    highfreq = highfreq or samplerate / 2
    assert highfreq <= samplerate / 2
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    bins = np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fbank = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fbank[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fbank[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fbank


def fbank(signal, samplerate=16000, winlen=0.025, winstep=0.01, nfilt=26, nfft=512,
          lowfreq=0, highfreq=None, preemph=0.97, winfunc=lambda x: np.ones((x,))):
    """psf.base.fbank → (filterbank energies [T,nfilt], frame energy [T])."""
    highfreq = highfreq or samplerate / 2
    signal = preemphasis(signal, preemph)
    frames = framesig(signal, winlen * samplerate, winstep * samplerate, winfunc)
    pspec = powspec(frames, nfft)
    energy = np.sum(pspec, 1)
    energy = np.where(energy == 0, EPS, energy)
    fb = get_filterbanks(nfilt, nfft, samplerate, lowfreq, highfreq)
    feat = np.dot(pspec, fb.T)
    feat = np.where(feat == 0, EPS, feat)
    return feat, energy


def dct2_ortho(x, numcep):
    """scipy.fftpack.dct(x, type=2, axis=1, norm='ortho')[:, :numcep], written out."""
    n = x.shape[1]
    k = np.arange(numcep)[:, None]
    i = np.arange(n)[None, :]
    basis = np.cos(np.pi * k * (2 * i + 1) / (2.0 * n))
    scale = np.full((numcep, 1), np.sqrt(2.0 / n))
    scale[0, 0] = np.sqrt(1.0 / n)
    return x @ (basis * scale).T


def lifter(cepstra, L=22):
    if L > 0:
        _, ncoeff = np.shape(cepstra)
        n = np.arange(ncoeff)
        lift = 1 + (L / 2.0) * np.sin(np.pi * n / L)
        return lift * cepstra
    return cepstra


def mfcc(signal, samplerate=16000, winlen=0.025, winstep=0.01, numcep=13, nfilt=26,
         nfft=None, lowfreq=0, highfreq=None, preemph=0.97, ceplifter=22, appendEnergy=True,
         winfunc=lambda x: np.ones((x,))):
    """psf.base.mfcc.  Returns float64 [T, numcep]."""
    if nfft is None:
        nfft = 1
        while nfft < winlen * samplerate:
            nfft *= 2
    feat, energy = fbank(signal, samplerate, winlen, winstep, nfilt, nfft, lowfreq, highfreq,
                         preemph, winfunc)
    feat = np.log(feat)
    feat = dct2_ortho(feat, numcep)
    feat = lifter(feat, ceplifter)
    if appendEnergy:
        feat[:, 0] = np.log(energy)
    return feat


def delta(feat, N):
    """Reference ``delta`` (speaker_identification.py:141-151): edge-padded regression
    sum_{k=-N..N} k*x[t+k] / (2*sum k^2)."""
    nf = len(feat)
    denominator = 2 * sum(i ** 2 for i in range(1, N + 1))
    padded = np.pad(feat, ((N, N), (0, 0)), mode="edge")
    out = np.empty_like(feat)
    w = np.arange(-N, N + 1)
    for t in range(nf):
        out[t] = np.dot(w, padded[t:t + 2 * N + 1]) / denominator
    return out


def mfcc39(sig, rate=16000, nfilt=26):
    """MFCC ‖ Δ ‖ ΔΔ as built at speaker_identification.py:386-389."""
    m = mfcc(sig, rate, winlen=0.025, winstep=0.01, nfft=512, nfilt=nfilt)
    d = delta(m, 2)
    dd = delta(d, 2)
    return np.concatenate((m, d, dd), axis=1)


def input_feature_gen(sig, rate=16000):
    """Reference ``input_feature_gen`` (speaker_identification.py:372-398) on an in-memory
    int16 array: 'silent' if len<4000, else float64 [1,256,39] (zero-padded / truncated)."""
    if len(sig) < 4000:
        return "silent"
    f = mfcc39(sig, rate)
    length = f.shape[0]
    if length < 256:
        f = np.concatenate((f, np.zeros((256 - length, 39))), axis=0)
    else:
        f = f[:256, :]
    return np.asarray([f])


def chunked_features(sig, rate=16000):
    """Whole-file MFCC-39 padded to a multiple of 256 frames and cut into chunks
    (speaker_identification.py:341-353; speaker_identification_post_processing.py:255-269).
    Returns float64 [ceil(T/256), 256, 39]."""
    f = mfcc39(sig, rate)
    length = f.shape[0]
    segs = math.ceil(length / 256)
    f = np.concatenate((f, np.zeros((segs * 256 - length, 39))), axis=0)
    return f.reshape(segs, 256, 39)
