"""GPU parity of the tensor-core MFCC kernel (csrc/mfcc_tc.cu) and of the kernel dispatch.

The reference parameterisation (mfcc(sig, rate, winlen=0.025, winstep=0.01, nfft=512), nfilt 26 or
40) runs on the tcgen05 two-stage DFT kernel; anything else runs on the general fp32 FFT kernel
(csrc/mfcc.cu).  Both must agree with the float64 oracle within the tolerance stated in
test_mfcc_gpu.py, and with each other.
"""
import os

import numpy as np
import pytest

from oracle import psf, synth
from test_mfcc_gpu import assert_mfcc_close, ref_mfcc

pytestmark = pytest.mark.gpu


def _with_kernel(kind, fn):
    old = os.environ.get("MMLA_MFCC_KERNEL")
    try:
        if kind is None:
            os.environ.pop("MMLA_MFCC_KERNEL", None)
        else:
            os.environ["MMLA_MFCC_KERNEL"] = kind
        return fn()
    finally:
        if old is None:
            os.environ.pop("MMLA_MFCC_KERNEL", None)
        else:
            os.environ["MMLA_MFCC_KERNEL"] = old


def _launch_names(fn):
    from mmla_audio_b200 import _lib
    import torch
    return [n for n, _ in _lib.trace_launches(fn, torch)]


@pytest.mark.parametrize("nfilt", [26, 40])
def test_tc_and_fft_kernels_agree(cuda, nfilt):
    from mmla_audio_b200 import speaker_identification as si
    pcm = synth.synth_clips(300, 5, 40000)
    cfg = si.MfccConfig(nfilt=nfilt)
    names_tc = _launch_names(lambda: si.mfcc_batch(pcm, cfg))
    names_fft = _with_kernel("fft", lambda: _launch_names(lambda: si.mfcc_batch(pcm, cfg)))
    assert names_tc == ["mfcc_tc_kernel"] and names_fft == ["mfcc_fused_kernel"]
    tc = si.mfcc_batch(pcm, cfg).cpu().numpy()
    fft = _with_kernel("fft", lambda: si.mfcc_batch(pcm, cfg).cpu().numpy())
    for i in range(pcm.shape[0]):
        ref = ref_mfcc(pcm[i], nfilt=nfilt)
        assert_mfcc_close(tc[i], ref)
        assert_mfcc_close(fft[i], ref)
        assert_mfcc_close(tc[i], fft[i])


def test_non_reference_parameters_use_the_general_kernel(cuda):
    """Hamming window / nfilt 30 are outside the tensor-core specialisation."""
    from mmla_audio_b200 import speaker_identification as si
    pcm = synth.synth_clips(310, 2, 24000)
    cfg = si.MfccConfig(nfilt=30)
    assert _launch_names(lambda: si.mfcc_batch(pcm, cfg)) == ["mfcc_fused_kernel"]
    out = si.mfcc_batch(pcm, cfg).cpu().numpy()
    for i in range(2):
        assert_mfcc_close(out[i], psf.mfcc(pcm[i], 16000, 0.025, 0.01, 13, 30, 512))


def test_tc_speaker_features_and_padding(cuda):
    """MFCC || delta || delta-delta, zero rows up to 256: mfcc_tc_kernel + mfcc_finish_kernel."""
    from mmla_audio_b200 import speaker_identification as si
    for clip_len in (24000, 40000, 40960, 4000, 4321):
        pcm = synth.synth_clips(320, 3, clip_len)
        names = _launch_names(lambda: si.speaker_features_batch(pcm))
        # clip starts must be 16-byte aligned for the TMA path: a 4321-sample stride is not
        want = ["mfcc_tc_kernel", "mfcc_finish_kernel"] if clip_len % 8 == 0 else ["mfcc_fused_kernel"]
        assert names == want, (clip_len, names)
        out = si.speaker_features_batch(pcm).cpu().numpy()
        assert out.shape == (3, 256, 39)
        for i in range(3):
            ref = psf.input_feature_gen(pcm[i])[0]
            assert_mfcc_close(out[i], ref)
            T = psf.num_frames(clip_len)
            assert not out[i, T:].any()


def test_tc_truncated_context_falls_back(cuda):
    """Clips longer than 256 frames with deltas need frames past the cut for the delta context."""
    from mmla_audio_b200 import speaker_identification as si
    pcm = synth.synth_clips(330, 2, 48000)                     # 299 frames -> truncated to 256
    names = _launch_names(lambda: si.speaker_features_batch(pcm))
    assert names == ["mfcc_fused_kernel"]
    out = si.speaker_features_batch(pcm).cpu().numpy()
    for i in range(2):
        assert_mfcc_close(out[i], psf.input_feature_gen(pcm[i])[0])


def test_tc_overlapping_windows_zero_copy(cuda):
    """segmentation() with step < win: a strided view of one recording (clip_stride < clip_len)."""
    import torch
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.pipeline import window_view
    rec = synth.synth_clips(340, 4, 40000).reshape(-1)
    x = torch.from_numpy(rec).cuda()
    wins = window_view(x, 24000, 8000)
    assert wins.shape[0] == int((rec.size - 24000) / 8000 + 1)
    out = si.mfcc_batch(wins).cpu().numpy()
    for j in (0, 1, wins.shape[0] - 1):
        assert_mfcc_close(out[j], ref_mfcc(rec[j * 8000:j * 8000 + 24000]))


def test_tc_stage_accumulators_match_numpy(cuda):
    """The raw tcgen05 accumulators of both DFT stages (diagnostic dump) against float64 numpy:
    stage 1 = 32-point real DFTs over the stride-16 polyphase components (the h = 1 half carries
    W64^k1), stage 2 = the 512-point spectrum scaled by 2^-6."""
    import torch
    from mmla_audio_b200 import _lib, speaker_identification as si
    lib = _lib.load()
    n_clips, L = 2, 24000
    pcm = synth.synth_clips(350, n_clips, L)
    cfg = si.MfccConfig()
    T = cfg.num_frames(L)
    gpc = (T + 15) // 16
    n_tiles = (n_clips * gpc + 3) // 4
    dbg = torch.full((n_tiles, 2, 128, 256), float("nan"), dtype=torch.float32, device="cuda")
    lib.mmla_debug_mfcc_tc_dump(dbg.data_ptr(), None)
    try:
        si.mfcc_batch(torch.from_numpy(pcm).cuda(), cfg)
        torch.cuda.synchronize()
    finally:
        lib.mmla_debug_mfcc_tc_dump(None, None)
    d = dbg.cpu().numpy().astype(np.float64)
    n1 = np.arange(32)
    k = np.arange(16)
    rng = np.random.default_rng(0)
    worst1 = worst2 = 0.0
    for _ in range(40):
        G = int(rng.integers(0, n_clips * gpc))
        fl, h, r = int(rng.integers(0, 16)), int(rng.integers(0, 2)), int(rng.integers(0, 8))
        clip, f0 = G // gpc, (G % gpc) * 16
        tile, g = G // 4, G % 4
        y = np.concatenate([psf.preemphasis(pcm[clip], 0.97) * 0.5, np.zeros(4096)])
        fr = y[(f0 + fl) * 160:(f0 + fl) * 160 + 512].copy()
        fr[400:] = 0
        if (f0 + fl) * 160 >= L:
            continue
        seq = fr[16 * n1 + 8 * h + r]
        S = np.array([np.sum(seq * np.exp(-2j * np.pi * n1 * kk / 32)) for kk in range(17)])
        S[1:16] *= np.exp(-2j * np.pi * k[1:] * h / 64)
        got = d[tile, 0, 8 * fl + r, (2 * g + h) * 32:(2 * g + h) * 32 + 32]
        exp = np.empty(32)
        exp[0], exp[1] = S[0].real, S[16].real
        exp[2::2], exp[3::2] = S[1:16].real, S[1:16].imag
        scale = max(np.abs(exp).max(), 1.0)
        worst1 = max(worst1, np.abs(got - exp).max() / scale)
        X = np.fft.fft(fr, 512) * 2.0 ** -5
        for p in range(2):
            row = p * 64 + 16 * g + fl
            for j in range(8):
                k1 = 2 * j + p
                for k2 in range(16):
                    c = 4 * (k2 >> 1) + (k2 & 1)
                    z = d[tile, 1, row, j * 32 + c] + 1j * d[tile, 1, row, j * 32 + c + 2]
                    worst2 = max(worst2, abs(z - X[k1 + 32 * k2]) / max(np.abs(X).max(), 1.0))
    assert worst1 < 2e-6 and worst2 < 4e-6, (worst1, worst2)


def test_row_stride_40_layout_and_pad40_classifier_input(cuda):
    """speaker_features_batch(row_stride=40): columns 0..38 identical to the dense layout, column 39
    zero, on both MFCC kernels; the TF32 speaker classifier gives identical probabilities from the
    [B,256,40] layout (MMLA_INPUT_F32_PAD40, no pad pass) and from [B,256,39]."""
    import torch
    from mmla_audio_b200 import models, speaker_identification as si, weights as W
    pcm = synth.synth_clips(360, 6, 24000)
    dense = si.speaker_features_batch(pcm)
    for kind in (None, "fft"):
        wide = _with_kernel(kind, lambda: si.speaker_features_batch(pcm, row_stride=40))
        ref = _with_kernel(kind, lambda: si.speaker_features_batch(pcm))
        assert wide.shape == (6, 256, 40)
        assert torch.equal(wide[:, :, :39], ref) and not wide[:, :, 39].any()
    wide13 = si.mfcc_batch(pcm, row_stride=16)
    assert wide13.shape[2] == 16 and torch.equal(wide13[:, :, :13], si.mfcc_batch(pcm)) and not wide13[:, :, 13:].any()
    spec = W.speaker_spec(10, "sigmoid")
    model = models.Model(spec, W.synthetic_weights(spec, 4321), precision="tf32")
    wide = si.speaker_features_batch(pcm, row_stride=40)
    names = _launch_names(lambda: model.predict_device(wide))
    assert "pad_channels_kernel" not in names
    assert "pad_channels_kernel" in _launch_names(lambda: model.predict_device(dense))
    p40, l40 = model.predict_device(wide)
    p39, l39 = model.predict_device(dense)
    assert torch.equal(p40, p39) and torch.equal(l40, l39)


@pytest.mark.parametrize("clip_len", [24000, 40000, 40960, 4000, 8800])
def test_classifier_from_cepstra_equals_classifier_from_features(cuda, clip_len):
    """mmla_net_forward_cepstra (delta / delta-delta / zero rows built inside the stem kernel from the MFCC-13 rows)
    must give what mmla_net_forward gives on the materialised [B,256,39] features, and both must match the oracle
    features -> torch-CPU classifier within the TF32 bar of test_nets_gpu.py."""
    import torch
    from mmla_audio_b200 import models, speaker_identification as si, weights as W
    from oracle import nets as onets
    pcm = synth.synth_clips(370, 5, clip_len)
    spec = W.speaker_spec(10, "sigmoid")
    w = W.synthetic_weights(spec, 4321)
    model = models.Model(spec, w, precision="tf32")
    cep = si.mfcc_batch(pcm, row_stride=16)
    assert cep.shape == (5, psf.num_frames(clip_len), 16) and not cep[:, :, 13:].any()
    names = _launch_names(lambda: model.predict_device_cepstra(cep))
    # the stem runs either in its own launch or inside the first ResNet stage's kernel; never a separate feature pass
    assert names[0] in ("stem_delta_fused_kernel", "stem_resstage_fused_kernel") and "mfcc_finish_kernel" not in names
    p_cep, l_cep = model.predict_device_cepstra(cep)
    p_feat, l_feat = model.predict_device(si.speaker_features_batch(pcm))
    assert torch.allclose(p_cep, p_feat, rtol=0, atol=2e-6), float((p_cep - p_feat).abs().max())
    assert torch.equal(l_cep, l_feat)
    ref_feat = np.concatenate([psf.input_feature_gen(pcm[i]) for i in range(5)]).astype(np.float32)
    ref = onets.speaker_forward(ref_feat, w, spec)
    assert np.abs(p_cep.cpu().numpy() - ref).max() <= 5e-3


def test_stem_inside_first_stage_matches_separate_stem(cuda, monkeypatch):
    """Label pipeline from MFCC-13 rows: the stem computed inside the first ResNet stage's kernel (features split by row
    parity, max-pool and shortcut operand straight from TMEM) vs the stand-alone stem kernel followed by the plain stage
    kernel.  Same MMAs in the same order on the same TF32 operands: probabilities must agree to fp32 round-off, labels
    exactly; clip lengths cover T < 128, T = 150, T = 250 and the full 256 frames, batches cover partial tile groups."""
    import torch
    from mmla_audio_b200 import models, speaker_identification as si, weights as W
    spec = W.speaker_spec(10, "sigmoid")
    w = W.synthetic_weights(spec, 4321)
    monkeypatch.setenv("MMLA_NET_FUSE_STEM", "0")
    separate = models.Model(spec, w, precision="tf32")
    monkeypatch.setenv("MMLA_NET_FUSE_STEM", "1")
    fused = models.Model(spec, w, precision="tf32")
    for n, clip_len in ((1, 24000), (5, 8800), (37, 40000), (300, 24000), (9, 40960)):
        cep = si.mfcc_batch(synth.synth_clips(11, n, clip_len), row_stride=16)
        assert "stem_resstage_fused_kernel" in _launch_names(lambda: fused.predict_device_cepstra(cep))
        assert "stem_delta_fused_kernel" in _launch_names(lambda: separate.predict_device_cepstra(cep))
        pf, lf = fused.predict_device_cepstra(cep)
        ps, ls = separate.predict_device_cepstra(cep)
        d = float((pf - ps).abs().max())
        print(f"stem in stage vs separate, {n} clips x {clip_len}: max |dprob| {d:.2e}")
        assert d <= 2e-6 and torch.equal(lf, ls)


# ---------------------------------------------------------------------------------------------
# mfcc_tc2_kernel: the 64 x 8 formulation (csrc/mfcc_tc2.inc), selected with MMLA_MFCC_TC=2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nfilt", [26, 40])
def test_tc2_formulation_matches_oracle_and_tc1(cuda, monkeypatch, nfilt):
    """Same bar as the 32 x 16 kernel (1e-4 rel + 1e-4 of max |ref|), several clip lengths incl. partial last groups."""
    from mmla_audio_b200 import speaker_identification as si
    cfg = si.MfccConfig(nfilt=nfilt)
    for clip_len in (40000, 24000, 40960, 4000, 8000):
        pcm = synth.synth_clips(700 + clip_len % 97, 5, clip_len)
        monkeypatch.delenv("MMLA_MFCC_TC", raising=False)
        tc1 = si.mfcc_batch(pcm, cfg).cpu().numpy()
        monkeypatch.setenv("MMLA_MFCC_TC", "2")
        assert _launch_names(lambda: si.mfcc_batch(pcm, cfg)) == ["mfcc_tc2_kernel"]
        tc2 = si.mfcc_batch(pcm, cfg).cpu().numpy()
        for i in range(pcm.shape[0]):
            ref = ref_mfcc(pcm[i], nfilt=nfilt)
            assert_mfcc_close(tc2[i], ref)
            assert_mfcc_close(tc2[i], tc1[i])


def test_tc2_features_padding_ragged_edges_and_windows(cuda, monkeypatch):
    """The callers of the kernel: MFCC-39 padded to 256 rows (finish pass), a ragged batch, the edge clips of
    test_mfcc_gpu.py, overlapping windows as a strided view, 16-float rows with the zero tail written by the epilogue."""
    import torch
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.pipeline import window_view
    monkeypatch.setenv("MMLA_MFCC_TC", "2")
    pcm = synth.synth_clips(720, 3, 24000)
    assert _launch_names(lambda: si.speaker_features_batch(pcm)) == ["mfcc_tc2_kernel", "mfcc_finish_kernel"]
    out = si.speaker_features_batch(pcm).cpu().numpy()
    for i in range(3):
        assert_mfcc_close(out[i], psf.input_feature_gen(pcm[i])[0])
    # ragged
    lens = [40000, 4001, 24000, 163, 33333]
    clips = [synth.synth_clips(750 + i, 1, n)[0] for i, n in enumerate(lens)]
    offs, pos = [], 0
    for c in clips:
        pos = (pos + 7) // 8 * 8
        offs.append(pos)
        pos += len(c)
    buf = np.zeros(pos + 8, np.int16)
    for o, c in zip(offs, clips):
        buf[o:o + len(c)] = c
    got, rows = si.mfcc_ragged(buf, offs, lens, with_deltas=True)
    got = got.cpu().numpy()
    for i, c in enumerate(clips):
        assert_mfcc_close(got[i, :rows[i]], psf.mfcc39(c))
    # edge clips
    rng = np.random.default_rng(5)
    cases = {
        "zero": np.zeros(8000, np.int16),
        "dc": np.full(8000, 1234, np.int16),
        "square": (np.where((np.arange(8000) // 40) % 2 == 0, 32767, -32768)).astype(np.int16),
        "short": rng.integers(-3000, 3000, 137).astype(np.int16),
        "one_frame": rng.integers(-3000, 3000, 400).astype(np.int16),
        "ragged": rng.integers(-20000, 20000, 12345).astype(np.int16),
        "noise_fs": rng.integers(-32768, 32767, 16000).astype(np.int16),
    }
    for name, sig in cases.items():
        g = si.mfcc_batch(sig)[0].cpu().numpy()
        ref = ref_mfcc(sig)
        assert g.shape == ref.shape, name
        if name == "zero":
            np.testing.assert_allclose(g, ref, rtol=1e-5, atol=1e-4, err_msg=name)
        else:
            assert_mfcc_close(g, ref)
    # overlapping windows (clip_stride < clip_len) and 16-float rows
    rec = synth.synth_clips(760, 4, 40000).reshape(-1)
    wins = window_view(torch.from_numpy(rec).cuda(), 24000, 8000)
    w = si.mfcc_batch(wins).cpu().numpy()
    for j in (0, 1, wins.shape[0] - 1):
        assert_mfcc_close(w[j], ref_mfcc(rec[j * 8000:j * 8000 + 24000]))
    wide = si.mfcc_batch(pcm, row_stride=16).cpu().numpy()
    assert wide.shape[-1] == 16 and not wide[..., 13:].any()
    for i in range(3):
        assert_mfcc_close(wide[i, :, :13], ref_mfcc(pcm[i]))
