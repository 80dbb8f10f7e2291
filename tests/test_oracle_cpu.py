"""CPU tests of the oracle itself: golden fixtures + cross-checks against INDEPENDENT
implementations available in this image (scipy, torchaudio, torch.nn).  The reference ships no
test vectors, so this is how the restatement is pinned ("parity unpinned" with
respect to the reference's own libraries, which cannot run here)."""
import math
import os

import numpy as np
import pytest
import scipy.fft
import scipy.fftpack
import torch

from oracle import librosa_mel as lm, nets as onets, psf, synth, tally as otally
from mmla_audio_b200 import weights as W

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_synth_is_deterministic_and_matches_golden(golden):
    clips = synth.synth_clips(0, 2, 8000)
    np.testing.assert_array_equal(clips, golden["pcm"])
    # clip i depends only on (seed, index): slicing the index range changes nothing
    np.testing.assert_array_equal(synth.synth_clips(1, 1, 8000)[0], clips[1])
    assert clips.dtype == np.int16 and np.abs(clips).max() > 1000


def test_psf_oracle_matches_golden(golden):
    pcm = golden["pcm"]
    np.testing.assert_allclose(psf.mfcc(pcm[0], 16000, 0.025, 0.01, nfft=512), golden["mfcc13"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(psf.mfcc(pcm[1], 16000, 0.025, 0.01, nfft=512, nfilt=40), golden["mfcc13_nfilt40"],
                               rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(psf.input_feature_gen(pcm[0])[0][:60], golden["feat39"], rtol=1e-10, atol=1e-9)


def test_librosa_oracle_matches_golden(golden):
    pcm = golden["pcm"]
    s_db, _ = lm.generate_mels(pcm[1])
    np.testing.assert_allclose(s_db[:, ::10], golden["s_db_cols"], rtol=0, atol=1e-4)
    np.testing.assert_array_equal(lm.generate_zcr(pcm[1]), golden["zcr"])
    img = lm.imsave_rgb_uint8(lm.generate_zcr_image(pcm[1]))[:, ::10]
    assert (np.abs(img.astype(int) - golden["image_cols"].astype(int)) <= 1).all()


def test_psf_shapes_pinned_by_reference():
    """Frame counts the reference relies on: 2.56 s -> 255 (<= 256 rows), 2.5 s -> 249, 1.5 s -> 149."""
    assert psf.num_frames(40960) == 255 and psf.num_frames(40000) == 249 and psf.num_frames(24000) == 149
    assert psf.num_frames(400) == 1 and psf.num_frames(401) == 2 and psf.num_frames(10) == 1
    sig = synth.synth_clips(3, 1, 40960)[0]
    assert psf.input_feature_gen(sig).shape == (1, 256, 39)
    assert psf.input_feature_gen(sig[:3999]) == "silent"
    assert psf.chunked_features(synth.synth_clips(0, 3, 40000).reshape(-1)).shape == (3, 256, 39)


def test_psf_pieces_vs_scipy():
    sig = synth.synth_clips(9, 1, 6010)[0]
    pre = psf.preemphasis(sig, 0.97)
    assert pre[0] == sig[0] and pre[5] == sig[5] - 0.97 * sig[4]
    frames = psf.framesig(pre, 400, 160)
    assert frames.shape == (psf.num_frames(6010), 400) == (37, 400)
    np.testing.assert_array_equal(frames[3], pre[480:880])
    assert frames[-1][-1] == 0.0                                    # zero-padded tail
    ps = psf.powspec(frames, 512)
    ref = np.abs(scipy.fft.rfft(frames, 512, axis=1)) ** 2 / 512
    np.testing.assert_allclose(ps, ref, rtol=1e-12)
    x = np.random.default_rng(0).normal(size=(7, 26))
    np.testing.assert_allclose(psf.dct2_ortho(x, 13), scipy.fftpack.dct(x, type=2, axis=1, norm="ortho")[:, :13],
                               rtol=1e-12, atol=1e-12)
    fb = psf.get_filterbanks(26, 512, 16000)
    assert fb.shape == (26, 257) and (fb >= 0).all() and fb.max() <= 1.0
    assert ((fb > 0).sum(axis=0) <= 2).all()                        # a bin feeds at most two triangles
    assert ((fb > 0).sum(axis=1) >= 1).all() and ((psf.get_filterbanks(40, 512, 16000) > 0).sum(axis=1) >= 1).all()
    assert (fb[:, 256] == 0).all()                                  # Nyquist bin carries no filter weight
    lift = psf.lifter(np.ones((1, 13)), 22)[0]
    np.testing.assert_allclose(lift, 1 + 11 * np.sin(np.pi * np.arange(13) / 22))


def test_psf_zero_signal_uses_eps():
    out = psf.mfcc(np.zeros(2000, np.int16), 16000, 0.025, 0.01, nfft=512)
    assert np.isfinite(out).all()
    assert out[0, 0] == pytest.approx(math.log(np.finfo(float).eps))


def test_delta_vs_torchaudio():
    import torchaudio
    x = np.random.default_rng(1).normal(size=(40, 13))
    ref = torchaudio.functional.compute_deltas(torch.from_numpy(x.T.copy())[None], win_length=5, mode="replicate")[0].numpy().T
    np.testing.assert_allclose(psf.delta(x, 2), ref, rtol=1e-10, atol=1e-12)


def test_slaney_mel_basis_vs_torchaudio():
    import torchaudio
    ta = torchaudio.functional.melscale_fbanks(201, 0.0, 8000.0, 128, 16000, norm="slaney", mel_scale="slaney").numpy().T
    fb = lm.mel_basis()
    assert fb.dtype == np.float32 and fb.shape == (128, 201)
    np.testing.assert_allclose(fb, ta, rtol=1e-4, atol=1e-7)
    assert (fb.sum(axis=1) > 0).all()                               # no empty filters at 128 mels / 201 bins


def test_librosa_pieces():
    sig = synth.synth_clips(2, 1, 30000)[0]
    y = lm.pad_or_truncate(lm.load_pcm(sig))
    assert y.dtype == np.float32 and len(y) == 24000 and np.abs(y).max() < 1
    spec = lm.stft(y)
    assert spec.shape == (201, 151) and spec.dtype == np.complex64
    # frame 10 is fully interior: equals a direct windowed rfft
    seg = y[10 * 160 - 200:10 * 160 + 200].astype(np.float64) * lm.hann_periodic(400)
    np.testing.assert_allclose(spec[:, 10], scipy.fft.rfft(seg).astype(np.complex64), rtol=1e-6, atol=1e-6)
    # frame 0 uses reflect padding
    pad = np.concatenate([y[1:201][::-1], y[:200]]).astype(np.float64) * lm.hann_periodic(400)
    np.testing.assert_allclose(spec[:, 0], scipy.fft.rfft(pad).astype(np.complex64), rtol=1e-6, atol=1e-6)
    s_db, norm = lm.generate_mels(sig)
    assert s_db.shape == (128, 151) and s_db.max() <= 1e-5 and s_db.min() >= -80.001
    assert norm.min() == 0.0 and norm.max() == 1.0
    zcr = lm.generate_zcr(sig)
    assert zcr.shape == (1, 151) and zcr.dtype == np.float64
    np.testing.assert_allclose(zcr * 400, np.rint(zcr * 400), atol=1e-9)      # k/400 with integer k
    short = lm.generate_zcr(sig[:1000])
    assert short[0, -1] == 0.0                                       # zero padding never crosses
    img = lm.classifier_input(sig)
    assert img.shape == (128, 151, 3) and img.dtype == np.float32 and img.max() <= 255 and img.min() >= 0


def test_imsave_truncates_and_flips():
    img = np.zeros((2, 1, 3))
    img[0, 0] = [0.999, 0.5, 1.0]
    out = lm.imsave_rgb_uint8(img)
    np.testing.assert_array_equal(out[1, 0], [254, 127, 255])       # truncation, rows flipped
    with pytest.raises(ValueError):
        lm.imsave_rgb_uint8(np.full((1, 1, 3), 1.5))


def test_lstm_vs_torch_nn_lstm():
    """Keras gate order i,f,c,o equals torch's i,f,g,o: check the restated LSTM against nn.LSTM
    with the kernel -> weight_ih / recurrent -> weight_hh transposes."""
    rng = np.random.default_rng(3)
    F_, u, T, B = 12, 16, 5, 3
    k = rng.normal(size=(F_, 4 * u)).astype(np.float32) * 0.3
    r = rng.normal(size=(u, 4 * u)).astype(np.float32) * 0.3
    b = rng.normal(size=(4 * u,)).astype(np.float32) * 0.1
    x = torch.from_numpy(rng.normal(size=(B, T, F_)).astype(np.float32))
    lstm = torch.nn.LSTM(F_, u, batch_first=True)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(torch.from_numpy(k.T))
        lstm.weight_hh_l0.copy_(torch.from_numpy(r.T))
        lstm.bias_ih_l0.copy_(torch.from_numpy(b))
        lstm.bias_hh_l0.zero_()
        ref_f = lstm(x)[0][:, -1]
        ref_b = lstm(torch.flip(x, dims=[1]))[0][:, -1]
    got_f = onets.lstm_last(x, torch.from_numpy(k), torch.from_numpy(r), torch.from_numpy(b), reverse=False)
    got_b = onets.lstm_last(x, torch.from_numpy(k), torch.from_numpy(r), torch.from_numpy(b), reverse=True)
    np.testing.assert_allclose(got_f.numpy(), ref_f.numpy(), atol=1e-6)
    np.testing.assert_allclose(got_b.numpy(), ref_b.numpy(), atol=1e-6)


def test_keras_same_padding_rules():
    assert onets._same_pads(128, 4, 1) == (1, 2)        # (4,1) kernel: 1 above, 2 below
    assert onets._same_pads(151, 3, 1) == (1, 1)
    assert onets._same_pads(151, 2, 2) == (0, 1)        # MaxPool on odd width pads on the right
    assert onets._same_pads(151, 1, 2) == (0, 0) and onets._same_pads(128, 1, 2) == (0, 0)
    assert onets._same_pads(256, 4, 1) == (1, 2)        # Conv1D k4


def test_nets_shapes_and_param_counts():
    wo = W.synthetic_weights(W.OVERLAP, 1)
    assert sum(v.size for v in wo.values()) == 1548706 - 0 + sum(2 * c for c in ())  # trainable+BN stats
    ws = W.synthetic_weights(W.SPEAKER_BASE, 1)
    assert sum(v.size for v in ws.values()) == 1491382 + 0
    x = np.random.default_rng(0).integers(0, 256, (2, 128, 151, 3)).astype(np.float32)
    p = onets.overlap_forward(x, wo, W.OVERLAP)
    assert p.shape == (2, 2) and np.allclose(p.sum(1), 1, atol=1e-6)
    xs = np.random.default_rng(0).normal(size=(2, 256, 39)).astype(np.float32)
    sp = W.speaker_spec(10, "sigmoid")
    q = onets.speaker_forward(xs, W.synthetic_weights(sp, 2), sp)
    assert q.shape == (2, 10) and (q > 0).all() and (q < 1).all()


def test_tally_oracle_matches_reference_expressions():
    from datetime import datetime
    names = ["overlapped", "non-overlapped", "overlapped", "silent", "overlapped"]
    lines = otally.log_rows(names, datetime(2021, 5, 4, 10, 0, 0, 250000), 1.5, "overlapped degree", False)
    assert lines[0] == "segment\toverlapped degree\ttimestamp"
    assert lines[1].split("\t")[2] == "2021-05-04 10:00:00.250000" and lines[2].split("\t")[2] == "2021-05-04 10:00:01.750000"
    counts, secs, total = otally.tally_from_log(lines, ["non-overlapped", "overlapped"])
    assert counts == {"non-overlapped": 1, "overlapped": 3, "silent": 1}
    assert total == 6.0 and secs == {"non-overlapped": int(0.2 * 6), "overlapped": int(0.6 * 6), "silent": int(0.2 * 6)}
    assert otally.num_windows(460800000, 24000, 24000) == 19200 and otally.num_windows(40000, 24000, 24000) == 1


# ---------------------------------------------------------------------------------------------
# END-TO-END cross-checks of the two feature chains against independent implementations.
# They cannot turn "parity unpinned" into pinned — the independent code is not the reference's own library either —
# but they catch a wrong window, pad mode, mel scale, normalisation or log convention in the restatement as a whole.
# ---------------------------------------------------------------------------------------------
def test_librosa_chain_vs_torchaudio_end_to_end():
    """oracle: librosa.load -> melspectrogram(n_fft=400, hop=160, 128 mel) -> power_to_db(ref=max)  versus
    torchaudio MelSpectrogram(norm='slaney', mel_scale='slaney', pad_mode='reflect', power=2) + 10 log10, top_db 80.
    Residuals (recorded): mel power <= 2e-5 relative to the clip maximum (float32 STFT in torchaudio vs float64 FFT
    stored as complex64 in the oracle); dB map <= 2e-3 dB wherever the value is above the -80 dB floor."""
    import torchaudio.transforms as T
    mel = T.MelSpectrogram(sample_rate=16000, n_fft=400, win_length=400, hop_length=160, n_mels=128, f_min=0.0, f_max=8000.0,
                           window_fn=torch.hann_window, power=2.0, center=True, pad_mode="reflect", norm="slaney",
                           mel_scale="slaney")
    worst_p, worst_db = 0.0, 0.0
    for clip in (3, 4, 5):
        sig = synth.synth_clips(clip, 1, 24000)[0]
        y = lm.pad_or_truncate(lm.load_pcm(sig))
        ref = mel(torch.from_numpy(y)[None]).numpy()[0]                      # [128, 151]
        got = lm.melspectrogram(y)
        assert got.shape == ref.shape == (128, 151)
        worst_p = max(worst_p, float(np.abs(got - ref).max() / ref.max()))
        db_ref = 10.0 * np.log10(np.maximum(ref, 1e-10)) - 10.0 * np.log10(max(ref.max(), 1e-10))
        db_ref = np.maximum(db_ref, db_ref.max() - 80.0)
        db_got = lm.power_to_db_refmax(got)
        live = db_ref > -79.0
        worst_db = max(worst_db, float(np.abs(db_got - db_ref)[live].max()))
        assert db_got.min() >= -80.0 - 1e-4 and abs(float(db_got.max())) <= 1e-5
    print("librosa chain vs torchaudio: mel power rel", worst_p, "dB", worst_db)
    assert worst_p <= 2e-5 and worst_db <= 2e-3


def test_psf_chain_vs_independent_scipy_pipeline_end_to_end():
    """oracle psf.mfcc versus a pipeline written a different way with scipy primitives only: scipy.signal.lfilter
    pre-emphasis, stride-trick framing, scipy.fft.rfft, a filterbank built from the closed-form triangle expression
    (not the per-bin loops), scipy.fftpack.dct(type=2, norm='ortho'), sinusoidal lifter, log-energy in c0.
    Residual (recorded): <= 1e-9 absolute on cepstra of magnitude up to ~40 (float64 on both sides)."""
    import scipy.signal
    worst = 0.0
    for clip, n, nfilt in ((6, 24000, 26), (7, 40960, 26), (8, 40000, 40), (9, 4321, 26)):
        sig = synth.synth_clips(clip, 1, n)[0].astype(np.float64)
        emph = scipy.signal.lfilter([1.0, -0.97], [1.0], sig)
        emph[0] = sig[0]
        T_ = 1 if n <= 400 else 1 + int(math.ceil((n - 400) / 160.0))
        padded = np.concatenate([emph, np.zeros((T_ - 1) * 160 + 400 - n)])
        frames = np.lib.stride_tricks.sliding_window_view(padded, 400)[::160][:T_]
        pspec = np.abs(scipy.fft.rfft(frames, 512, axis=1)) ** 2 / 512.0
        energy = pspec.sum(1)
        energy[energy == 0] = np.finfo(float).eps
        mel_pts = np.linspace(0.0, 2595 * np.log10(1 + 8000 / 700.0), nfilt + 2)
        bins = np.floor(513 * (700 * (10 ** (mel_pts / 2595.0) - 1)) / 16000.0)
        k = np.arange(257)[None, :]
        lo, ce, hi = bins[:-2, None], bins[1:-1, None], bins[2:, None]
        fb = np.where((k >= lo) & (k < ce), (k - lo) / (ce - lo), 0.0) + np.where((k >= ce) & (k < hi), (hi - k) / (hi - ce), 0.0)
        fe = pspec @ fb.T
        fe[fe == 0] = np.finfo(float).eps
        cep = scipy.fftpack.dct(np.log(fe), type=2, axis=1, norm="ortho")[:, :13]
        cep *= 1 + 11.0 * np.sin(np.pi * np.arange(13) / 22.0)
        cep[:, 0] = np.log(energy)
        got = psf.mfcc(synth.synth_clips(clip, 1, n)[0], 16000, winlen=0.025, winstep=0.01, nfft=512, nfilt=nfilt)
        assert got.shape == cep.shape == (T_, 13)
        worst = max(worst, float(np.abs(got - cep).max()))
    print("psf chain vs independent scipy pipeline: max |d|", worst)
    assert worst <= 1e-9


def test_noisereduce_oracle_against_hand_rolled_stft():
    """The spectral-gate oracle leans on scipy.signal.stft / istft; check the conventions it assumes about them (periodic
    Hann, boundary zeros, 1/sum(w) scaling, perfect reconstruction with an all-ones mask)."""
    from oracle import noisereduce_stationary as onr
    from scipy.signal import stft, istft
    x = synth.synth_clips(2, 1, 8192)[0].astype(np.float64) / 32768
    _, _, Z = stft(x, nfft=1024, noverlap=768, nperseg=1024, padded=False)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024)
    xp = np.concatenate([np.zeros(512), x, np.zeros(512)])
    m = 7
    mine = np.fft.rfft(w * xp[256 * m: 256 * m + 1024]) / w.sum()
    np.testing.assert_allclose(Z[:, m], mine, atol=1e-12)
    assert Z.shape == (513, (len(x) + 256) // 256)
    _, back = istft(Z, nfft=1024, noverlap=768, nperseg=1024)
    np.testing.assert_allclose(back[: len(x)], x, atol=1e-12)
    f = onr._smoothing_filter(16, 3)
    assert f.shape == (33, 7) and abs(f.sum() - 1) < 1e-12 and np.isclose(f[16, 3] * 68, 1.0)
