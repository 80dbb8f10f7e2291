"""GPU parity: silence removal (WebRTC VAD mode 3 + vad_collector + clip rewrite) through the C-ABI `mmla_vad_trim`
against the CPU restatement (oracle/webrtc_vad.c + oracle/webrtc_vad.py).  Everything here is integer work: the bar is
BIT-EXACT — per-frame decisions, collector mask, rewritten length, rewritten samples, and the per-clip 'silent' labels."""
import numpy as np
import pytest

from oracle import librosa_mel as lm, nets as onets, psf, synth, webrtc_vad as ovad

pytestmark = pytest.mark.gpu


def _edge_clips(L):
    rng = np.random.default_rng(7)
    t = np.arange(L)
    clips = {
        "zeros": np.zeros(L, np.int16),
        "dc": np.full(L, 1234, np.int16),
        "fs_square": np.where((t // 40) % 2 == 0, 32767, -32768).astype(np.int16),
        "min_value": np.full(L, -32768, np.int16),
        "faint_noise": (rng.standard_normal(L) * 20).astype(np.int16),
        "loud_noise": (rng.standard_normal(L) * 6000).clip(-32768, 32767).astype(np.int16),
        "burst": np.concatenate([np.zeros(L // 3, np.int16), synth.synth_clips(5, 1, L)[0][: L - 2 * (L // 3)],
                                 np.zeros(L // 3, np.int16)]),
        "chirp": (12000 * np.sin(2 * np.pi * (100 + 3000 * t / L) * t / 16000)).astype(np.int16),
    }
    return clips


def _oracle_batch(pcm, lengths=None, chained=False):
    """Per clip: (flags, keep, trimmed) from the oracle; `chained` carries one Vad across the clips in order."""
    out = []
    vad = ovad.Vad(3)
    for i in range(pcm.shape[0]):
        if not chained:
            vad.reset()
        sig = pcm[i] if lengths is None else pcm[i][: lengths[i]]
        trimmed, flags, keep = ovad.remove_silence(sig, vad)
        out.append((flags, keep, trimmed))
    return out


def _check(res, ref, max_frames):
    speech, keep = res.speech.cpu().numpy(), res.keep.cpu().numpy()
    vlen, out = res.voiced_len.cpu().numpy(), res.pcm.cpu().numpy()
    for i, (flags, k, trimmed) in enumerate(ref):
        nf = len(flags)
        np.testing.assert_array_equal(speech[i, :nf], flags, err_msg=f"is_speech flags, clip {i}")
        np.testing.assert_array_equal(keep[i, :nf], k, err_msg=f"collector mask, clip {i}")
        assert not speech[i, nf:].any() and not keep[i, nf:].any()
        assert vlen[i] == len(trimmed) == 480 * int(k.sum())
        np.testing.assert_array_equal(out[i, : vlen[i]], trimmed, err_msg=f"rewritten samples, clip {i}")


@pytest.mark.parametrize("L", [40960, 24000, 8000])
def test_vad_independent_clips_bit_exact(cuda, L):
    """Batch mode: every clip starts from a fresh Vad(3); 2.56 s (85 frames), 1.5 s (exact multiple of 480: the last
    frame is dropped, 49 frames) and 0.5 s clips, synthetic multi-speaker clips plus edge clips."""
    from mmla_audio_b200 import vad
    clips = [synth.synth_clips(100, 40, L)] + [c[None, :] for c in _edge_clips(L).values()]
    pcm = np.concatenate(clips)
    assert vad.num_frames(L) == ovad.num_frames(L) == (L - 1) // 480
    res = vad.vad_trim(pcm)
    _check(res, _oracle_batch(pcm), vad.num_frames(L))
    assert res.speech.cpu().numpy().any() and not res.speech.cpu().numpy().all()   # the batch has both outcomes


def test_vad_ragged_lengths_and_short_clips(cuda):
    from mmla_audio_b200 import vad
    L = 20000
    pcm = synth.synth_clips(300, 12, L)
    lengths = np.array([20000, 19999, 481, 480, 479, 0, 960, 961, 4799, 4800, 4801, 12345], np.int32)
    res = vad.vad_trim(pcm, lengths=lengths)
    _check(res, _oracle_batch(pcm, lengths), vad.num_frames(L))


def test_vad_session_state_carries_across_clips(cuda):
    """The reference keeps ONE module-global Vad object: clips of a session are one stream.  clips_per_stream = B chains
    the detector through all clips in order; clips_per_stream = 4 restarts it every 4 clips."""
    from mmla_audio_b200 import vad
    pcm = synth.synth_clips(700, 24, 24000)
    res = vad.vad_trim(pcm, clips_per_stream=24)
    chained = _oracle_batch(pcm, chained=True)
    _check(res, chained, vad.num_frames(24000))
    fresh = _oracle_batch(pcm)
    assert any(not np.array_equal(a[0], b[0]) for a, b in zip(chained, fresh)), "state carry-over had no effect"
    res4 = vad.vad_trim(pcm, clips_per_stream=4)
    ref4 = []
    for g in range(0, 24, 4):
        ref4 += _oracle_batch(pcm[g:g + 4], chained=True)
    _check(res4, ref4, vad.num_frames(24000))


def test_reference_signature_frame_generator_and_collector(cuda):
    """`frame_generator(30, audio_bytes, 16000)` + `vad_collector(16000, 30, 300, vad, frames)` as the scripts call them."""
    from mmla_audio_b200 import vad
    sig = synth.synth_clips(41, 1, 40960)[0]
    frames = list(vad.frame_generator(30, sig.tobytes(), 16000))
    assert len(frames) == 85 and len(frames[0].bytes) == 960 and abs(frames[1].timestamp - 0.03) < 1e-12
    segments = list(vad.vad_collector(16000, 30, 300, vad.Vad(3), frames))
    got = np.frombuffer(b"".join(segments), dtype=np.int16)
    want, _, _ = ovad.remove_silence(sig, ovad.Vad(3))
    np.testing.assert_array_equal(got, want)


def _silence_mix(n, L):
    """Clips of which some have too little voiced audio to survive the 4000-sample rule."""
    pcm = synth.synth_clips(2000, n, L)
    rng = np.random.default_rng(3)
    for i in range(0, n, 3):                                   # every third clip: faint noise only
        pcm[i] = (rng.standard_normal(L) * 15).astype(np.int16)
    for i in range(1, n, 6):                                   # some: a short burst (< 10 voiced frames at the start)
        pcm[i, 3000:] = 0
    return pcm


def test_speaker_pipeline_silence_removed_labels_vs_oracle(cuda):
    """SI record_on_pc.py:117-140 per clip: save_wave_file(silence_remove=True) -> input_feature_gen ('silent' if the
    rewritten clip has < 4000 samples) -> predict -> argmax."""
    from mmla_audio_b200 import models, weights as W
    from mmla_audio_b200.pipeline import SpeakerPipeline
    spec = W.speaker_spec(10, "sigmoid")
    w = W.synthetic_weights(spec, 4321)
    pcm = _silence_mix(48, 40960)
    pipe = SpeakerPipeline(models.Model(spec, w, precision="fp32"))
    labels, prob = pipe.run_device(cuda.from_numpy(pcm).cuda(), silence_removed=True)
    got = labels.cpu().numpy()
    ref = _oracle_batch(pcm)
    n_silent = 0
    for i, (_f, _k, trimmed) in enumerate(ref):
        x = psf.input_feature_gen(trimmed)
        if isinstance(x, str):
            assert got[i] == -1, f"clip {i} must be 'silent'"
            n_silent += 1
            continue
        p = onets.speaker_forward(x.astype(np.float32), w, spec)
        assert np.abs(prob[i].cpu().numpy() - p[0]).max() <= 2e-4
        s = np.sort(p[0])
        if s[-1] - s[-2] > 1e-3:
            assert got[i] == int(p[0].argmax())
    assert 0 < n_silent < 48


def test_overlap_pipeline_silence_removed_labels_vs_oracle(cuda):
    """record_on_pc.py:133-160 per clip (overlap): the image is computed from the rewritten clip."""
    from mmla_audio_b200 import models, weights as W
    from mmla_audio_b200.pipeline import OverlapPipeline
    w = W.synthetic_weights(W.OVERLAP, 1234)
    pcm = _silence_mix(18, 40960)
    pipe = OverlapPipeline(models.Model(W.OVERLAP, w, precision="fp32"))
    labels, prob = pipe.run_device(cuda.from_numpy(pcm).cuda(), silence_removed=True)
    got = labels.cpu().numpy()
    n_silent = 0
    for i, (_f, _k, trimmed) in enumerate(_oracle_batch(pcm)):
        if len(trimmed) < 4000:
            assert got[i] == -1
            n_silent += 1
            continue
        p = onets.overlap_forward(lm.classifier_input(trimmed)[None], w, W.OVERLAP)
        assert np.abs(prob[i].cpu().numpy() - p[0]).max() <= 1e-3          # image may differ by 1 LSB on < 1 % of pixels
        if abs(p[0, 0] - p[0, 1]) > 5e-3:
            assert got[i] == int(p[0].argmax())
    assert 0 < n_silent < 18


def test_per_clip_silent_rule_with_explicit_lengths(cuda):
    """`len(sig) < 4000` is a per-clip decision: 3999 -> 'silent', 4000 -> classified."""
    from mmla_audio_b200 import models, weights as W
    from mmla_audio_b200.pipeline import OverlapPipeline, SpeakerPipeline
    pcm = cuda.from_numpy(synth.synth_clips(9, 6, 24000)).cuda()
    lengths = np.array([24000, 3999, 4000, 0, 12000, 3999], np.int32)
    spec = W.speaker_spec(10, "sigmoid")
    sp = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision="tf32"))
    ls, ps = sp.run_device(pcm, lengths=lengths)
    op = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision="tf32"))
    lo, _ = op.run_device(pcm, lengths=lengths)
    for lab in (ls.cpu().numpy(), lo.cpu().numpy()):
        assert (lab[[1, 3, 5]] == -1).all() and (lab[[0, 2, 4]] >= 0).all()
    # and the classified ones equal the same clips run on their own at that length
    host = pcm.cpu().numpy()
    for i in (2, 4):                                            # (uniform entry: MFCC-13 rows -> stem kernel; ragged: features)
        _l1, p1 = sp.run_device(cuda.from_numpy(host[i:i + 1, : lengths[i]].copy()).cuda())
        assert (p1[0] - ps[i]).abs().max().item() <= 5e-3
