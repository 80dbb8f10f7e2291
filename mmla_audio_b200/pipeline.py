"""End-to-end hot path, batched: PCM → features → classifier → arg-max labels → tallies.

These are the loops of the reference's entry scripts with the microphone / WAV / PNG plumbing
removed (BASELINE.json north_star: "audio fed from synthetic buffers instead of PyAudio"):
  * ``SpeakerPipeline``  — SpeakerIdentification/scripts/record_on_pc.py:97-151 (per clip) and
    speaker_identification_post_processing.py:253-312 (whole session)
  * ``OverlapPipeline``  — OverlapDetection/scripts/record_on_pc.py:114-173 (per clip) and
    overlap_detection_post_processing.py:189-226 (whole session)
"""
from __future__ import annotations

from datetime import datetime
from typing import Dict, Optional

import numpy as np

from . import _lib, tally
from .models import Model
from .overlap_features_generator import OverlapFeaturesGenerator
from .params import (MfccConfig, OVERLAP_CLIP_SAMPLES, SILENT_MIN_SAMPLES, SPEAKER_FRAMES)
from .speaker_identification import (mfcc_batch, mfcc_ragged, speaker_features_batch, whole_file_chunks,
                                     _to_device_pcm)


def segmentation_windows(n_samples: int, win: int, step: int) -> int:
    """``cut_num = int((nframes - win)/step + 1)`` (overlap_detection_post_processing.py:55-59)."""
    return int(((n_samples - win) / step) + 1)


def window_view(pcm_1d, win: int, step: int):
    """Zero-copy [cut_num, win] strided view of a long recording (window j = [j*step, j*step+win))."""
    n = segmentation_windows(pcm_1d.numel(), win, step)
    return pcm_1d.as_strided((max(n, 0), win), (step, 1))


class PendingResult:
    """Handle of one in-flight :meth:`SpeakerPipeline.submit_host` batch.

    The caller's ``pcm_host`` buffer is read by asynchronous copies: it must stay untouched until
    :meth:`wait_uploaded` returns (or ``uploaded.query()`` is true).  A handle is valid until its
    result slot is recycled, i.e. for ``depth`` further submissions; after that :meth:`result`
    raises instead of returning another batch's labels."""

    def __init__(self, slot, done, uploaded, generation):
        self._slot, self._done = slot, done
        self.uploaded = uploaded                # CUDA event: the last host->device slice of this batch has landed
        self._generation = generation

    def _check(self):
        if self._slot["generation"] != self._generation:
            raise RuntimeError("this PendingResult's slot was recycled by a later submit_host(); "
                               "fetch results within `depth` submissions or raise `depth`")

    @property
    def labels_dev(self):
        """int32 CUDA [B]; valid until ``depth`` more submissions."""
        self._check()
        return self._slot["labels"]

    def wait_uploaded(self):
        """Block until every slice of the caller's ``pcm_host`` has been copied to the device; the
        host buffer may then be refilled with the next batch."""
        self.uploaded.synchronize()

    def result(self):
        """Block until this batch's labels and tallies are on the host; returns numpy copies."""
        self._check()
        self._done.synchronize()
        self._check()
        self._slot["fetched"] = True
        return self._slot["labels_h"].numpy().copy(), self._slot["counts_h"].numpy().copy()


class _HostApi:
    """Host-buffer entry points shared by both pipelines (they only need ``self.run_device``)."""

    def submit_host(self, pcm_host, n_classes: int, n_chunks: int = 2, depth: int = 3, reduce=None):
        """Asynchronous end-to-end pass from HOST memory: ``pcm_host`` int16 [B, L] (pinned for
        full speed).  The batch is cut into ``n_chunks`` slices; a copy stream uploads slice i+1
        while slice i runs features + classifier on the compute stream, and the labels + tallies
        are read back into pinned host buffers behind the last slice.  Nothing blocks the host:
        the call returns a :class:`PendingResult`; up to ``depth`` submissions may be in flight, so
        the upload of batch k+1 overlaps the compute of batch k (the recording loop of the
        reference scripts, pipelined).  ``PendingResult.result()`` waits for that batch only.
        ``pcm_host`` is read asynchronously: do not modify it before ``PendingResult.wait_uploaded()``.

        ``reduce(labels_dev, counts_dev) -> (labels_all, counts_all)``: optional device-side step run
        on the compute stream before the read-back — the multi-GPU label all_gather / tally
        all_reduce, one collective (``sharding.exchange_labels_and_counts``); its outputs are what is
        copied to the host."""
        torch = _lib.require_cuda()
        B, L = pcm_host.shape
        n_chunks = max(1, min(n_chunks, B))
        bounds = [(B * i) // n_chunks for i in range(n_chunks + 1)]
        cmax = max(bounds[i + 1] - bounds[i] for i in range(n_chunks))
        key = (B, cmax, L, n_classes, depth)
        if getattr(self, "_stage_key", None) != key:
            nbuf = 3
            self._stage = [torch.empty((cmax, L), dtype=torch.int16, device="cuda") for _ in range(nbuf)]
            self._stage_free = [None] * nbuf            # event: last consumer of the buffer is done
            self._stage_next = 0
            self._copy_stream = torch.cuda.Stream()
            self._slots = [dict(labels=torch.empty((B,), dtype=torch.int32, device="cuda"),
                                labels_h=torch.empty((B,), dtype=torch.int32).pin_memory(),
                                counts_h=torch.empty((n_classes + 1,), dtype=torch.int64).pin_memory(),
                                done=None, generation=0, fetched=True) for _ in range(depth)]
            self._slot_next = 0
            self._stage_key = key
        slot = self._slots[self._slot_next]
        self._slot_next = (self._slot_next + 1) % len(self._slots)
        if slot["done"] is not None:
            slot["done"].synchronize()                  # the slot's previous batch must have finished
        slot["generation"] += 1                         # handles of the slot's previous batch now raise
        slot["fetched"] = False
        compute = torch.cuda.current_stream()
        uploaded = None
        labels = slot["labels"]
        for i in range(n_chunks):
            lo, hi = bounds[i], bounds[i + 1]
            k = self._stage_next
            self._stage_next = (k + 1) % len(self._stage)
            buf = self._stage[k][: hi - lo]
            copied = torch.cuda.Event()
            with torch.cuda.stream(self._copy_stream):
                if self._stage_free[k] is not None:
                    self._copy_stream.wait_event(self._stage_free[k])
                buf.copy_(pcm_host[lo:hi], non_blocking=True)
                copied.record(self._copy_stream)
            uploaded = copied
            compute.wait_event(copied)
            lab, _ = self.run_device(buf)
            labels[lo:hi] = lab
            free = torch.cuda.Event()
            free.record(compute)
            self._stage_free[k] = free
        counts = tally.device_counts(labels, n_classes)
        labels_out, counts_out = (labels, counts) if reduce is None else reduce(labels, counts)
        if slot["labels_h"].numel() != labels_out.numel():
            slot["labels_h"] = torch.empty((labels_out.numel(),), dtype=torch.int32).pin_memory()
        slot["labels_h"].copy_(labels_out.reshape(-1), non_blocking=True)
        slot["counts_h"].copy_(counts_out, non_blocking=True)
        done = torch.cuda.Event()
        done.record(compute)
        slot["done"] = done
        return PendingResult(slot, done, uploaded, slot["generation"])

    def run_host(self, pcm_host, n_classes: int, n_chunks: int = 2):
        """Synchronous form of :meth:`submit_host`: returns (labels int32 numpy [B], counts int64
        numpy [n_classes+1])."""
        return self.submit_host(pcm_host, n_classes, n_chunks).result()


class SpeakerPipeline(_HostApi):
    def __init__(self, model: Model, cfg: MfccConfig = MfccConfig(), n_streams: int = 1, split_min_clips: int = 1024):
        """``n_streams`` > 1: batches of at least ``split_min_clips`` clips are cut into that many slices that
        run on their own CUDA streams, so kernels that cannot fill the GPU on their own (the persistent
        BiLSTM kernel runs 128 clips per CTA: 64 CTAs for 4096 clips on 148 SMs) overlap with the other
        slice's convolutions instead of leaving SMs idle.  Measured on B200 (4096 clips): 1 stream 2.09 ms,
        2 streams 2.20 ms, 4 streams 2.23 ms — the kernels are persistent / grid-filling, so slices mostly
        serialise and lose batch efficiency; the default therefore stays 1."""
        if model.spec.ndim != 1:
            raise ValueError("SpeakerPipeline needs the speaker net")
        self.model = model
        self.cfg = cfg
        self.n_streams = max(1, int(n_streams))
        self.split_min_clips = int(split_min_clips)
        self._feat = {}
        self._side = None

    def _run_slice(self, torch, pcm_dev, slot: int):
        B, L = pcm_dev.shape
        T = self.cfg.num_frames(L)
        if self.model.precision in ("tf32", "fp16") and self.cfg.numcep == 13 and T <= SPEAKER_FRAMES:
            # label pipeline: MFCC-13 rows only; delta / delta-delta / padding happen inside the classifier's stem
            # kernel, so the [B,256,39] feature tensor never exists in HBM
            cep = self._feat.get(slot)
            if cep is None or tuple(cep.shape) != (B, T, 16) or cep.device != pcm_dev.device:
                cep = self._feat[slot] = torch.empty((B, T, 16), dtype=torch.float32, device=pcm_dev.device)
            mfcc_batch(pcm_dev, self.cfg, out=cep, row_stride=16)
            prob, labels = self.model.predict_device_cepstra(cep)
            return labels, prob
        width = 3 * self.cfg.numcep
        feat = self._feat.get(slot)
        if feat is None or tuple(feat.shape) != (B, SPEAKER_FRAMES, width) or feat.device != pcm_dev.device:
            feat = self._feat[slot] = torch.empty((B, SPEAKER_FRAMES, width), dtype=torch.float32, device=pcm_dev.device)
        speaker_features_batch(pcm_dev, self.cfg, out=feat)
        prob, labels = self.model.predict_device(feat)
        return labels, prob

    def run_device(self, pcm_dev, lengths=None, silence_removed: bool = False, vad_clips_per_stream: int = 1):
        """pcm_dev: int16 CUDA [B, L].  → (labels int32 [B], prob [B,n]).

        ``lengths`` (int32 [B], host or device): ragged clips, clip i = ``pcm_dev[i, :lengths[i]]``.
        ``silence_removed=True``: every clip first goes through the WebRTC VAD + ``vad_collector``
        (``save_wave_file(.., silence_remove=True)``, SI record_on_pc.py:117 → :185-204) and the features are
        taken from the rewritten (voiced-only) clip.  Either way the reference's rule ``len(sig) < 4000 =>
        'silent'`` (speaker_identification.py:375) is applied PER CLIP: such clips get label -1 and skip the
        classifier; their ``prob`` rows are zero."""
        torch = _lib.require_cuda()
        B = pcm_dev.shape[0]
        if silence_removed:
            from .vad import vad_trim
            res = vad_trim(pcm_dev, lengths, clips_per_stream=vad_clips_per_stream)
            pcm_dev, lengths = res.pcm, res.voiced_len
        if lengths is not None:
            return self._run_ragged(torch, pcm_dev, lengths)
        if pcm_dev.shape[1] < SILENT_MIN_SAMPLES:          # every clip is 'silent'
            return (torch.full((B,), tally.SILENT, dtype=torch.int32, device=pcm_dev.device), None)
        n = self.n_streams if B >= self.split_min_clips else 1
        if n == 1:
            return self._run_slice(torch, pcm_dev, 0)
        if self._side is None or len(self._side) != n - 1:
            self._side = [torch.cuda.Stream() for _ in range(n - 1)]
        main = torch.cuda.current_stream()
        bounds = [(B * i) // n for i in range(n + 1)]
        ready = torch.cuda.Event()
        ready.record(main)
        outs = [None] * n
        for i in range(1, n):
            st = self._side[i - 1]
            st.wait_event(ready)                              # the input was produced on the caller's stream
            with torch.cuda.stream(st):
                outs[i] = self._run_slice(torch, pcm_dev[bounds[i]:bounds[i + 1]], i)
        outs[0] = self._run_slice(torch, pcm_dev[bounds[0]:bounds[1]], 0)
        for st in self._side:
            main.wait_stream(st)
        labels = torch.cat([o[0] for o in outs])
        prob = torch.cat([o[1] for o in outs])
        for o in outs[1:]:                                    # tensors allocated on side streams, consumed on `main`
            o[0].record_stream(main)
            o[1].record_stream(main)
        return labels, prob

    def _run_ragged(self, torch, pcm_dev, lengths):
        """Per-clip lengths: clips shorter than 4000 samples are 'silent' (-1); the others go through the ragged
        feature entry (``mmla_psf_mfcc`` with per-clip offsets / lengths → [n,256,39]) and the classifier."""
        B, L = pcm_dev.shape
        ln = torch.as_tensor(lengths).to(torch.int32).cpu().numpy()          # one small read-back: the rule is a host decision
        if ln.shape != (B,):
            raise ValueError("lengths must have one entry per clip")
        ln = np.minimum(ln, L)
        live = np.nonzero(ln >= SILENT_MIN_SAMPLES)[0]
        labels = torch.full((B,), tally.SILENT, dtype=torch.int32, device=pcm_dev.device)
        prob = torch.zeros((B, self.model.spec.n_classes), dtype=torch.float32, device=pcm_dev.device)
        if len(live) == 0:
            return labels, prob
        stride0 = pcm_dev.stride(0) if B > 1 else L
        flat = pcm_dev.as_strided(((B - 1) * stride0 + L,), (1,))
        feat, _rows = mfcc_ragged(flat, live.astype(np.int64) * stride0, ln[live], self.cfg, with_deltas=True,
                                  pad_frames=SPEAKER_FRAMES)
        p_live, l_live = self.model.predict_device(feat)
        idx = torch.from_numpy(live).to(pcm_dev.device)
        labels[idx] = l_live
        prob[idx] = p_live
        return labels, prob

    def run_session(self, pcm_long, speaker_names: Dict[int, str], t0: Optional[datetime] = None,
                    silent_index=(), log_path: Optional[str] = None):
        """Offline session: MFCC-39 over the whole recording, 256-frame chunks, one predict,
        rows every 2.56 s (speaker_identification_post_processing.py:253-312).  With ``log_path`` the
        TSV log the reference appends row by row (:278-312) is written for ``visualization()``."""
        torch = _lib.require_cuda()
        speaker_names = tally.normalize_names(speaker_names)     # make_feature_experiment's dict has str keys
        chunks = whole_file_chunks(pcm_long, self.cfg)
        prob, labels = self.model.predict_device(chunks)
        if len(silent_index):
            idx = torch.as_tensor(list(silent_index), dtype=torch.long, device=labels.device)
            labels[idx] = tally.SILENT
        t0 = t0 or datetime.today()
        if log_path is not None:
            from . import distributions
            rows = [speaker_names.get(int(l), "silent") for l in labels.cpu().tolist()]
            distributions.write_log(log_path, tally.log_rows(rows, t0, 2.56, "speaker", add_before_first=True))
        return labels, tally.tally_session(labels, speaker_names, t0, 2.56, add_before_first=True)


    def session_chunks_sharded(self, pcm_long, rank: int, world: int):
        """This rank's share of ``whole_file_chunks(pcm_long)``: float32 CUDA [chunk_hi - chunk_lo, 256, 39], identical
        to rows [chunk_lo, chunk_hi) of the single-GPU result.  The rank reads its chunk range plus a halo (5 frames
        before, 4 frames + one window after — ``sharding.session_slice``) from the shared recording; nothing is exchanged."""
        from .sharding import session_slice
        torch = _lib.require_cuda()
        x = _to_device_pcm(torch, pcm_long).reshape(-1)
        sl = session_slice(x.numel(), rank, world, self.cfg.frame_len, self.cfg.frame_step, SPEAKER_FRAMES)
        n_own = sl["chunk_hi"] - sl["chunk_lo"]
        width = 3 * self.cfg.numcep
        if n_own <= 0:
            return torch.empty((0, SPEAKER_FRAMES, width), dtype=torch.float32, device=x.device), sl
        piece = x[sl["sample_lo"]:sl["sample_hi"]]
        feat = mfcc_batch(piece.reshape(1, -1), self.cfg, with_deltas=True)[0]          # [frames of the slice, 39]
        rows = feat[sl["skip_rows"]:sl["skip_rows"] + n_own * SPEAKER_FRAMES]
        if sl["chunk_hi"] < sl["n_chunks_total"]:
            rows = rows[: n_own * SPEAKER_FRAMES]                                        # the tail rows are halo
        else:                                                                            # last chunk of the file: zero rows
            real = sl["n_frames_total"] - sl["chunk_lo"] * SPEAKER_FRAMES
            rows = rows[:real]
        out = torch.zeros((n_own * SPEAKER_FRAMES, width), dtype=torch.float32, device=x.device)
        out[: rows.shape[0]] = rows
        return out.view(n_own, SPEAKER_FRAMES, width), sl

    def run_session_sharded(self, pcm_long, speaker_names: Dict[int, str], rank: int, world: int,
                            t0: Optional[datetime] = None, silent_index=()):
        """:meth:`run_session` with the recording's chunks split across ``world`` ranks (one process per GPU): every
        rank labels its chunk range, one all_gather (``sharding.exchange_labels_and_counts``) gives every rank all
        labels, and the tallies are computed from the gathered labels — identical on every rank and to the
        single-GPU session."""
        from .sharding import exchange_labels_and_counts
        torch = _lib.require_cuda()
        speaker_names = tally.normalize_names(speaker_names)
        chunks, sl = self.session_chunks_sharded(pcm_long, rank, world)
        n_classes = self.model.spec.n_classes
        if chunks.shape[0]:
            _prob, labels = self.model.predict_device(chunks)
        else:
            labels = torch.empty((0,), dtype=torch.int32, device="cuda")
        counts = tally.device_counts(labels, n_classes)
        labels, _counts = exchange_labels_and_counts(labels, counts, sl["n_chunks_total"], rank, world)
        labels = labels.clone()
        if len(silent_index):
            labels[torch.as_tensor(list(silent_index), dtype=torch.long, device=labels.device)] = tally.SILENT
        return labels, tally.tally_session(labels, speaker_names, t0 or datetime.today(), 2.56, add_before_first=True)


class OverlapPipeline(_HostApi):
    def __init__(self, model: Model):
        if model.spec.ndim != 2:
            raise ValueError("OverlapPipeline needs the overlap net")
        self.model = model
        self.ofg = OverlapFeaturesGenerator(wl=25, hl=10)

    def run_device(self, pcm_dev, lengths=None, silence_removed: bool = False, vad_clips_per_stream: int = 1):
        """pcm_dev: int16 CUDA [B, L].  → (labels int32 [B], prob [B,2]); label -1 = 'silent' for every clip
        with fewer than 4000 samples (record_on_pc.py:141-154), decided PER CLIP when ``lengths`` are given or
        ``silence_removed=True`` runs the WebRTC VAD + ``vad_collector`` first (``save_wave_file(..,
        silence_remove=True)``, record_on_pc.py:133 → :214-226).  The features of a trimmed clip are taken from
        its first 24000 voiced samples, zero-padded (overlap_features_generator.py:73-80)."""
        torch = _lib.require_cuda()
        B, L = pcm_dev.shape
        if silence_removed:
            from .vad import vad_trim
            res = vad_trim(pcm_dev, lengths, clips_per_stream=vad_clips_per_stream)
            pcm_dev, lengths = res.pcm, res.voiced_len
            L = pcm_dev.shape[1]
        if lengths is not None:
            ln = np.minimum(torch.as_tensor(lengths).to(torch.int32).cpu().numpy(), L)
            if ln.shape != (B,):
                raise ValueError("lengths must have one entry per clip")
            live = np.nonzero(ln >= SILENT_MIN_SAMPLES)[0]
            labels = torch.full((B,), tally.SILENT, dtype=torch.int32, device=pcm_dev.device)
            prob = torch.zeros((B, 2), dtype=torch.float32, device=pcm_dev.device)
            if len(live):
                if len(live) == B:
                    img = self.ofg.classifier_input_batch(pcm_dev, lengths_host=ln)
                else:
                    idx = torch.from_numpy(live).to(pcm_dev.device)
                    img = self.ofg.classifier_input_batch(pcm_dev.index_select(0, idx), lengths_host=ln[live])
                p_live, l_live = self.model.predict_device(img)
                if len(live) == B:
                    return l_live, p_live
                labels[idx] = l_live
                prob[idx] = p_live
            return labels, prob
        if L < SILENT_MIN_SAMPLES:
            return (torch.full((B,), tally.SILENT, dtype=torch.int32, device=pcm_dev.device), None)
        img = self.ofg.classifier_input_batch(pcm_dev)
        prob, labels = self.model.predict_device(img)
        return labels, prob

    def run_session(self, pcm_long, t0: Optional[datetime] = None, win_s: float = 1.5, step_s: float = 1.5,
                    sr: int = 16000, chunk: int = 4096, log_path: Optional[str] = None):
        """Offline session: cut 1.5 s windows (zero-copy), features + predict per window, rows every
        1.5 s (overlap_detection_post_processing.py:189-226).  With ``log_path`` the TSV log of
        :213-224 is written for ``visualization()``."""
        torch = _lib.require_cuda()
        x = _to_device_pcm(torch, pcm_long).reshape(-1)
        wins = window_view(x, int(sr * win_s), int(sr * step_s))
        labels = torch.empty((wins.shape[0],), dtype=torch.int32, device=x.device)
        for i in range(0, wins.shape[0], chunk):
            l, _ = self.run_device(wins[i:i + chunk])
            labels[i:i + chunk] = l
        t0 = t0 or datetime.today()
        names = {int(k): v for k, v in tally.OVERLAP_DEGREE_DICT.items()}
        if log_path is not None:
            from . import distributions
            rows = [names.get(int(l), "silent") for l in labels.cpu().tolist()]
            distributions.write_log(log_path, tally.log_rows(rows, t0, win_s, "overlapped degree", add_before_first=False))
        return labels, tally.tally_session(labels, names, t0, win_s, add_before_first=False,
                                           initial_order=list(tally.OVERLAP_DEGREE_DICT.values()))
