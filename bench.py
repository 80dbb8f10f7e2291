#!/usr/bin/env python
"""bench.py — headline benchmark of the mmla-audio hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload speaker_id|bulk_mfcc|overlap]

Metric (BASELINE.json): audio-seconds per second, MFCC + classifier, whole job over N GPUs.
Default workload = BASELINE.json configs[1]: speaker identification, 1.5 s windows, 4096
synthetic clips per GPU, 10 registered speakers.  One step = one pass of the hot path over the
rank's clips: fused MFCC-39 features -> speaker classifier forward -> arg-max labels -> label
tally (+ label all_gather / tally all_reduce when N > 1).  Clips are sharded across ranks with
no data-path collective (weak scaling).

Printed JSON line (rank 0): the driver contract plus `roofline`, `cpu_baseline`, `e2e`, `clocks`,
`gpu_launches`, and an `extra` object with per-stage timings and the bulk MFCC-only number
(configs[2] shape: 2.5 s clips, nfilt 40) on the same GPU.

`--impl reference` times the reference-equivalent CPU path (oracle/ port: numpy float64
python_speech_features restatement + torch-CPU classifier — the upstream libraries cannot be
installed here) on the host cores over a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
WORKLOADS = {
    "speaker_id": dict(name="speaker_id_1.5s_x4096_10spk", clips=4096, clip_len=24000, nfilt=26),
    "bulk_mfcc": dict(name="bulk_mfcc13_2.5s_nfilt40", clips=65536, clip_len=40000, nfilt=40),
    "overlap": dict(name="overlap_1.5s_x512", clips=512, clip_len=24000, nfilt=0),
}
SPEAKER_FLOP_PER_CLIP = 46.9e6
OVERLAP_FLOP_PER_CLIP = 1.838e9


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt, self.active = threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            if self.active:
                try:
                    self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.01)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------------
# CPU reference-equivalent path (oracle port) — cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------
def _cpu_features_speaker(args):
    first, n, clip_len = args
    from oracle import psf, synth
    pcm = synth.synth_clips(first, n, clip_len)
    return np.concatenate([psf.input_feature_gen(pcm[i]) for i in range(n)]).astype(np.float32)


def cpu_step(workload: str, first_clip: int, n_clips: int, clip_len: int, state: dict, pool=None):
    """One bounded CPU pass of the same hot path on `n_clips` synthetic clips."""
    from oracle import librosa_mel as lm, nets as onets, psf, synth
    if workload == "speaker_id":
        if pool is not None:
            per = max(1, n_clips // pool._processes)
            jobs = [(first_clip + i, min(per, n_clips - i), clip_len) for i in range(0, n_clips, per)]
            x = np.concatenate(pool.map(_cpu_features_speaker, jobs))
        else:
            x = _cpu_features_speaker((first_clip, n_clips, clip_len))
        prob = onets.speaker_forward(x, state["w"], state["spec"])
        labels = np.argmax(prob, axis=1)
    elif workload == "bulk_mfcc":
        pcm = synth.synth_clips(first_clip, n_clips, clip_len)
        for i in range(n_clips):
            psf.mfcc(pcm[i], SR, winlen=0.025, winstep=0.01, nfft=512, nfilt=40)
        labels = np.zeros(n_clips, np.int64)
    else:
        pcm = synth.synth_clips(first_clip, n_clips, clip_len)
        x = np.stack([lm.classifier_input(pcm[i]) for i in range(n_clips)])
        prob = onets.overlap_forward(x, state["w"], state["spec"])
        labels = np.argmax(prob, axis=1)
    return np.bincount(labels, minlength=2)


def cpu_state(workload: str):
    from mmla_audio_b200 import weights as W
    if workload == "speaker_id":
        spec = W.speaker_spec(10, "sigmoid")
        return {"spec": spec, "w": W.synthetic_weights(spec, 4321)}
    if workload == "overlap":
        return {"spec": W.OVERLAP, "w": W.synthetic_weights(W.OVERLAP, 1234)}
    return {}


_THREAD_ENV = ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMEXPR_NUM_THREADS")


def _single_thread_worker():
    """Pool initializer: one BLAS/OpenMP thread per worker process, whatever launched bench.py.  Without it the
    workers inherit un-capped thread pools when the arm is started as plain `python` (N = 1) and oversubscribe the
    host ~4x, while under torchrun (N > 1, OMP_NUM_THREADS=1 exported) they do not — the r01 reference arm read
    690 audio-s/s at N = 1 and 2.6-2.7 K at N = 2/4/8 for that reason alone."""
    for k in _THREAD_ENV:
        os.environ[k] = "1"
    try:
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass


def make_feature_pool(clip_len: int):
    """All host cores as single-threaded feature workers (the reference's feature code is single-threaded numpy)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    saved = {k: os.environ.get(k) for k in _THREAD_ENV}
    for k in _THREAD_ENV:                       # spawn children read these before numpy / torch initialise
        os.environ[k] = "1"
    try:
        pool = mp.get_context("spawn").Pool(min(cores, 32), initializer=_single_thread_worker)
        pool.map(_cpu_features_speaker, [(0, 1, clip_len)] * pool._processes)   # import warm-up
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return pool, cores


def time_cpu(workload: str, clip_len: int, sample_clips: int, steps: int, warmup: int, all_cores: bool):
    import torch
    state = cpu_state(workload)
    pool = None
    cores = torch.get_num_threads()
    if all_cores:
        torch.set_num_threads(os.cpu_count() or 1)     # torchrun exports OMP_NUM_THREADS=1: undo it for the classifier
        cores = torch.get_num_threads()
    if all_cores and workload == "speaker_id":
        pool, cores = make_feature_pool(clip_len)
    for i in range(warmup):
        cpu_step(workload, i * sample_clips, sample_clips, clip_len, state, pool)
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_step(workload, (warmup + i) * sample_clips, sample_clips, clip_len, state, pool)
    dt = (time.perf_counter() - t0) / steps
    if pool is not None:
        pool.close()
    return sample_clips * clip_len / SR / dt, dt, cores


def run_reference_arm(a, wl):
    """`--impl reference`: the reference-equivalent CPU path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # --steps / --warmup are honoured as given; the per-step sample is bounded so that the default
    # driver call (20 + 5 steps) stays under a minute of CPU time
    sample = {"speaker_id": 1024, "bulk_mfcc": 256, "overlap": 8}[a.workload]
    steps, warmup = max(1, a.steps), max(0, a.warmup)
    value, dt, cores = time_cpu(a.workload, wl["clip_len"], sample, steps, warmup, all_cores=True)
    desc = (f"{sample} synthetic clips of {wl['clip_len'] / SR:.2f} s per step (oracle port: numpy float64 features in "
            f"{min(cores, 32)} single-threaded worker processes + torch-CPU fp32 classifier on {cores} threads)")
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "clips_per_step": sample, "clip_seconds": wl["clip_len"] / SR,
                   "note": "CPU baseline does not scale with --gpus; rank 0 only. clips_per_step is a bounded sample "
                           "(the GPU arm runs %d clips per GPU per step); the metric is throughput-normalised "
                           "(audio-s/s), so the arms compare per second, not per step" % wl["clips"]},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# Algorithmic work per step of each kernel name (DESIGN.md section 4 states the per-unit figures)
# ---------------------------------------------------------------------------------------------
def kernel_work(workload: str, B: int, L: int, pipe, bulk_frames: int):
    """{kernel name: {"bound", "per_step" (bytes or FLOP over all of that kernel's launches in one step), "what"}}"""
    T = 1 if L <= 400 else 1 + -(-(L - 400) // 160)                 # psf frame count
    work = {}
    if workload == "bulk_mfcc":
        nb = B * (L * 2 + bulk_frames * 13 * 4)
        for k in ("mfcc_tc_kernel", "mfcc_fused_kernel"):
            work[k] = {"bound": "hbm", "per_step": nb, "what": "int16 PCM in + float32 [T,13] cepstra out"}
        return work
    if workload == "overlap":
        for k in ("overlap_features_tc_kernel", "overlap_features_kernel"):
            work[k] = {"bound": "hbm", "per_step": B * (24000 * 2 + 128 * 151 * 3),
                       "what": "int16 PCM (24000 samples) in + uint8 [128,151,3] image out"}
        spec, T0, H0 = pipe.model.spec, 151, 128
    else:
        rows = min(T, 256)
        work["mfcc_tc_kernel"] = {"bound": "hbm", "per_step": B * (L * 2 + rows * 13 * 4),
                                  "what": "int16 PCM in + float32 [T,13] cepstra out (deltas / padding: mfcc_finish_kernel)"}
        work["mfcc_fused_kernel"] = {"bound": "hbm", "per_step": B * (L * 2 + 256 * 39 * 4),
                                     "what": "int16 PCM in + float32 [256,39] features out"}
        work["mfcc_finish_kernel"] = {"bound": "hbm", "per_step": B * (rows * 13 * 4 + (256 * 39 - rows * 13) * 4),
                                      "what": "[T,13] cepstra in + delta / delta-delta columns and zero rows out"}
        spec, T0, H0 = pipe.model.spec, 256, 1
    # classifier convolutions: 2 FLOP per multiply-add, Keras 'same' output sizes
    h, w_ = H0, T0
    res = 0.0
    slab = 0.0                                   # overlap: the stride-1 3x3 / 4x1 convs (conv_slab_kernel)
    slab_first = 0.0                             # ... of the first block (stem_resblock2d_fused_kernel)
    slab_23 = 0.0                                # ... of blocks 2 and 3 (resblock2d_persist_kernel)
    pool_bytes = 0.0                             # overlap: pool_shortcut_kernel's compulsory traffic
    res_first_stage = 0.0                        # the first three residual units (the stage the stem is folded into)
    stem = 2.0 * h * w_ * spec.stem.kh * spec.stem.kw * spec.stem.cin * spec.stem.cout
    for bi, blk in enumerate(spec.blocks):
        if blk.pool:
            h2, w2 = (-(-h // 2) if spec.ndim == 2 else h), -(-w_ // 2)
        else:
            h2, w2 = h, w_
        if spec.ndim == 2:                       # overlap: both convs at full size, then MaxPool; shortcut strided
            res += 2.0 * h * w_ * blk.conv1.kh * blk.conv1.kw * blk.conv1.cin * blk.conv1.cout
            res += 2.0 * h * w_ * blk.conv2.kh * blk.conv2.kw * blk.conv2.cin * blk.conv2.cout
            slab += 2.0 * h * w_ * (blk.conv1.kh * blk.conv1.kw * blk.conv1.cin * blk.conv1.cout +
                                    blk.conv2.kh * blk.conv2.kw * blk.conv2.cin * blk.conv2.cout)
            if bi == 0:
                slab_first = slab
            if bi == 2:
                slab_23 = slab - slab_first
        else:                                    # speaker: MaxPool first, both convs at the pooled length
            res += 2.0 * h2 * w2 * blk.conv1.kh * blk.conv1.kw * blk.conv1.cin * blk.conv1.cout
            res += 2.0 * h2 * w2 * blk.conv2.kh * blk.conv2.kw * blk.conv2.cin * blk.conv2.cout
        if blk.shortcut is not None:
            res += 2.0 * h2 * w2 * blk.shortcut.cin * blk.shortcut.cout
            pool_bytes += 4.0 * (h * w_ * blk.conv2.cout + h2 * w2 * blk.shortcut.cin + h2 * w2 * blk.shortcut.cout)
        h, w_ = h2, w2
        if bi == 2:
            res_first_stage = res
    t_lstm = w_ if spec.ndim == 2 else w_ // 4                       # mean over H (overlap) / AvgPool1D(4) (speaker)
    feat = spec.blocks[-1].conv2.cout
    xproj = 2.0 * 2 * t_lstm * feat * 1024
    work["lstm_fused_kernel"] = {"bound": "tensor", "per_step": B * 2.0 * 2 * (t_lstm - 1) * 256 * 1024,
                                 "what": "recurrent products h U of both directions, T-1 steps"}
    work["xproj_fused_kernel"] = {"bound": "hbm", "per_step": B * t_lstm * (128 + 2 * 1024) * 4,
                                  "what": "[B*T,128] features in + both directions' [B*T,1024] input projections out"}
    conv_flop = B * (stem + res + xproj)
    if spec.ndim == 1:
        work["resunit_fused_kernel"] = {"bound": "tensor", "per_step": B * res, "what": "the 9 residual units (18 convs + 3 shortcuts)"}
        work["resstage_fused_kernel"] = {"bound": "tensor", "per_step": B * (res - res_first_stage),
                                         "what": "ResNet stages 2 and 3 = 6 residual units (12 convs + 2 shortcuts), TF32"}
        work["stem_resstage_fused_kernel"] = {"bound": "tensor", "per_step": B * (stem + res_first_stage),
                                              "what": "stem Conv1D(32,4) + ResNet stage 1 (3 residual units), features built "
                                                      "from the MFCC-13 rows, TF32"}
        work["conv_tc_kernel"] = {"bound": "tensor", "per_step": B * xproj, "what": "both LSTM input projections"}

        work["stem_fused_kernel"] = {"bound": "hbm", "per_step": B * 256 * (40 + 32) * 4,
                                     "what": "[256,40] features in + [256,32] activations out (Conv1D k=4)"}
        work["stem_delta_fused_kernel"] = {"bound": "hbm", "per_step": B * (rows * 13 + 256 * 32) * 4,
                                           "what": "[T,13] cepstra in + [256,32] activations out (delta, delta-delta, padding, Conv1D k=4)"}
        work["conv_igemm_kernel"] = {"bound": "tensor", "per_step": conv_flop, "what": "all convolutions + LSTM input projections"}
    else:
        work["conv_igemm_kernel"] = {"bound": "tensor", "per_step": conv_flop, "what": "all convolutions + LSTM input projections"}
        work["conv_slab_kernel"] = {"bound": "tensor", "per_step": B * slab,
                                    "what": "the 18 stride-1 3x3 / 4x1 convolutions of the residual blocks (TF32, tap-shifted slabs)"}
        work["resblock2d_fused_kernel"] = {"bound": "tensor", "per_step": B * (slab - slab_first),
                                           "what": "residual blocks 2-9: conv pairs (3x3 then 4x1, TF32, tap-shifted slabs, the "
                                                   "intermediate stays in shared memory): 16 convolutions in 8 launches"}
        work["stem_resblock2d_fused_kernel"] = {"bound": "tensor", "per_step": B * (slab_first + stem),
                                                "what": "stem Conv2D(16, 1x1) from the uint8 image + residual block 1's conv pair "
                                                        "at 128 x 151 (TF32)"}
        work["resblock2d_persist_kernel"] = {"bound": "tensor", "per_step": B * slab_23,
                                             "what": "residual blocks 2-3 (64 x 76 x 32): conv pairs on the persistent, "
                                                     "warp-specialised kernel (TF32, weights resident in shared memory)"}
        work["stem_resblock2d_persist_kernel"] = {"bound": "tensor", "per_step": B * (slab_first + stem),
                                                  "what": "stem Conv2D(16, 1x1) from the uint8 image + residual block 1's conv pair "
                                                          "at 128 x 151 (TF32), persistent warp-specialised kernel"}
        work["resblock2d_f16_kernel"] = {"bound": "tensor", "per_step": B * (slab - slab_first),
                                         "what": "residual blocks 2-9: conv pairs with fp16 operands / fp32 accumulation "
                                                 "(precision fp16): 16 convolutions in 8 launches"}
        work["stem_resblock2d_f16_kernel"] = {"bound": "tensor", "per_step": B * (slab_first + stem),
                                              "what": "stem Conv2D(16, 1x1) from the uint8 image + residual block 1's conv pair "
                                                      "at 128 x 151, fp16 operands / fp32 accumulation"}
        work["resblock2d_persist_f16_kernel"] = {"bound": "tensor", "per_step": B * slab_23,
                                                 "what": "residual blocks 2-3 (64 x 76 x 32): conv pairs with fp16 operands on the "
                                                         "persistent, warp-specialised kernel"}
        work["stem_resblock2d_persist_f16_kernel"] = dict(work["stem_resblock2d_f16_kernel"],
                                                          what=work["stem_resblock2d_f16_kernel"]["what"] + ", persistent warp-specialised kernel")
        if os.environ.get("MMLA_NET_PERSIST") == "2":
            work["resblock2d_f16_kernel"] = dict(work["resblock2d_f16_kernel"], per_step=B * (slab - slab_first - slab_23),
                                                 what="residual blocks 4-9 (C >= 64): conv pairs with fp16 operands / fp32 accumulation")
        if os.environ.get("MMLA_NET_PERSIST") == "2":        # blocks 2-3 on the persistent kernel as well
            work["resblock2d_fused_kernel"]["per_step"] = B * (slab - slab_first - slab_23)
            work["resblock2d_fused_kernel"]["what"] = ("residual blocks 4-9 (C >= 64): conv pairs (3x3 then 4x1, TF32, tap-shifted "
                                                       "slabs, the intermediate stays in shared memory)")
        work["conv_tc_kernel"] = {"bound": "tensor", "per_step": B * (res - slab),
                                  "what": "the three stride-2 1x1 shortcut convs (TF32, im2col gather)"}
        work["pool_shortcut_kernel"] = {"bound": "hbm", "per_step": B * pool_bytes,
                                        "what": "per pooled block: full-resolution conv output in + every other block-input pixel in "
                                                "+ pooled output out (MaxPool2x2 + stride-2 1x1 shortcut + add)"}
        work["stem1x1_kernel"] = {"bound": "hbm", "per_step": B * H0 * T0 * (3 + 16 * 4),
                                  "what": "uint8 [128,151,3] image in + float32 [128,151,16] stem activations out"}
    return work


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="speaker_id", choices=sorted(WORKLOADS))
    ap.add_argument("--clips", type=int, default=0, help="clips per GPU (default: workload's)")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32", "fp16"],
                    help="classifier matmul arithmetic: tf32 = tcgen05 tensor cores, fp32 = CUDA cores")
    ap.add_argument("--e2e-chunks", type=int, default=2, help="host pipeline: upload/compute overlap slices per pass")
    ap.add_argument("--e2e-depth", type=int, default=2, help="host pipeline: passes in flight (1 = synchronous)")
    ap.add_argument("--streams", type=int, default=1,
                    help="speaker_id: slices of a batch run on this many CUDA streams so under-filled kernels overlap")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    a = ap.parse_args()
    wl = dict(WORKLOADS[a.workload])
    if a.clips and a.clips != wl["clips"]:
        # keep the label honest: "..._x512" -> "..._x19200" when --clips overrides the workload's batch
        wl["name"] = wl["name"].replace("_x%d" % wl["clips"], "_x%d" % a.clips) if ("_x%d" % wl["clips"]) in wl["name"] \
            else "%s_x%d" % (wl["name"], a.clips)
        wl["clips"] = a.clips
    if a.impl == "reference":
        run_reference_arm(a, wl)
        return
    if a.warmup < 3:
        a.warmup = 3

    import torch
    import torch.distributed as dist
    from mmla_audio_b200 import _lib, models, synth, tally, weights as W
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.pipeline import OverlapPipeline, SpeakerPipeline
    from mmla_audio_b200.sharding import bind_host_thread_to_gpu, exchange_labels_and_counts, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for --impl ours")
    torch.cuda.set_device(local_rank)
    numa_bound = bind_host_thread_to_gpu(physical_gpu_index(local_rank)) if world > 1 else False
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    peaks = load_peaks()

    B, L = wl["clips"], wl["clip_len"]
    n_total = B * world
    lo, hi = shard_range(n_total, rank, world)
    pcm = synth.synth_clips(lo, hi - lo, L)                       # this rank's shard, resident in HBM
    pcm_host = torch.empty((B, L), dtype=torch.int16).pin_memory()
    pcm_host.copy_(pcm.cpu())
    pcm_dev2 = torch.empty_like(pcm)                              # e2e staging target

    if a.workload == "speaker_id":
        spec = W.speaker_spec(10, "sigmoid")
        pipe = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision=a.precision),
                               n_streams=a.streams)
        n_classes = 10
    elif a.workload == "overlap":
        pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision=a.precision))
        n_classes = 2
    else:
        pipe, n_classes = None, 1
        cfg40 = si.MfccConfig(nfilt=40)
        out_bulk = torch.empty((B, cfg40.num_frames(L), 13), dtype=torch.float32, device="cuda")

    def step(x):
        if pipe is None:
            si.mfcc_batch(x, cfg40, out=out_bulk)
            return None, None
        labels, _ = pipe.run_device(x)
        counts = tally.device_counts(labels, n_classes)
        if world > 1:
            labels, counts = exchange_labels_and_counts(labels, counts, n_total, rank, world)
        return labels, counts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    for _ in range(a.warmup):
        step(pcm)
    launches0 = lib.mmla_launch_count()
    sampler.active = True
    ms_step = timed(lambda: step(pcm), a.steps)
    sampler.active = False
    gpu_launches = int(lib.mmla_launch_count() - launches0)
    audio_s = n_total * L / SR
    value = audio_s / (ms_step / 1e3)

    # ---- e2e: host buffers in, labels + tallies out, copies inside the timed region -----------
    def e2e_finish(pending):
        return pending.result()                                  # host numpy: the step's (gathered) labels + tallies

    def e2e_reduce(labels_dev, counts_dev):                      # device side, on the compute stream, before the read-back
        return exchange_labels_and_counts(labels_dev, counts_dev, n_total, rank, world)

    def e2e_run(steps):
        """`steps` passes through the public host API.  Every pass uploads its own PCM from pinned
        host memory and reads its own labels + tallies back; up to `--e2e-depth` passes are in
        flight so the upload of pass k+1 overlaps the compute of pass k."""
        if a.workload in ("speaker_id", "overlap"):
            from collections import deque
            pend = deque()
            chunks = a.e2e_chunks if a.workload == "speaker_id" else 1      # overlap is compute-bound: whole batches
            for _ in range(steps):
                pend.append(pipe.submit_host(pcm_host, n_classes, n_chunks=chunks, depth=a.e2e_depth + 1,
                                             reduce=e2e_reduce if world > 1 else None))
                if len(pend) >= a.e2e_depth:
                    e2e_finish(pend.popleft())
            while pend:
                e2e_finish(pend.popleft())
            return
        for _ in range(steps):
            pcm_dev2.copy_(pcm_host, non_blocking=True)
            labels, counts = step(pcm_dev2)
            if labels is None:
                out_bulk[:, 0, 0].sum().item()
            else:
                labels.cpu(), counts.cpu()
    e2e_run(3)
    e2e_steps = max(6, a.steps // 2)
    sampler.active = True
    ms_e2e = timed(lambda: e2e_run(e2e_steps), 1) / e2e_steps
    sampler.active = False
    h2d = B * L * 2
    ms_h2d = timed(lambda: pcm_dev2.copy_(pcm_host, non_blocking=True), 5)     # the PCIe floor of e2e
    d2h = (n_total * 4 + (n_classes + 1) * 8) if pipe is not None else 4

    # ---- per-kernel device times (library launch trace: one CUDA event per launch on the step's stream, taken
    #      live over `steps` passes of the timed step) and the roofline of the dominant kernel ---------------------
    extra = {"h2d_only_ms": ms_h2d, "h2d_gbs": h2d / (ms_h2d * 1e-3) / 1e9}
    barrier()
    streams_saved = getattr(pipe, "n_streams", 1)
    if pipe is not None and streams_saved > 1:
        pipe.n_streams = 1            # per-kernel durations need the launches serialised on one stream
    trace = _lib.trace_launches(lambda: [step(pcm) for _ in range(a.steps)], torch)
    barrier()
    if pipe is not None and streams_saved > 1:
        pipe.n_streams = streams_saved
    per = {}
    for name, ms in trace:
        d = per.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ms
    traced_ms = sum(v[1] for v in per.values()) / a.steps
    kernels = [{"kernel": k, "launches_per_step": v[0] / a.steps, "ms_per_step": v[1] / a.steps,
                "share_of_step": v[1] / a.steps / ms_step} for k, v in per.items()]
    kernels.sort(key=lambda d: -d["ms_per_step"])
    extra["kernels"] = kernels
    extra["traced_ms_per_step"] = traced_ms
    extra["kernels_note"] = ("per-kernel device times from a single-stream pass over the same steps (launch trace); "
                             "the timed step runs %d stream slice(s), so shares are relative to ms_per_step" % streams_saved)
    work = kernel_work(a.workload, B, L, pipe, out_bulk.shape[1] if pipe is None else 0)
    dom = kernels[0]
    w = work.get(dom["kernel"])
    launch_ms = dom["ms_per_step"] / dom["launches_per_step"]
    if w is None:
        roofline = {"bound": None, "kernel": dom["kernel"], "achieved": None, "peak": None, "unit": None, "frac": None,
                    "traffic": None, "note": "no algorithmic figure recorded for this kernel"}
    elif w["bound"] == "hbm":
        gbs = w["per_step"] / (dom["ms_per_step"] * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "traffic": None,
                    "algorithmic_bytes_per_launch": w["per_step"] / dom["launches_per_step"], "what": w["what"],
                    "launch_ms": launch_ms, "launches_per_step": dom["launches_per_step"],
                    "share_of_step": dom["share_of_step"], "peak_source": peaks["source"] + " HBM copy"}
    else:
        tf = w["per_step"] / (dom["ms_per_step"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": dom["kernel"], "achieved": tf, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops_sustained"], "traffic": None,
                    "algorithmic_flop_per_step": w["per_step"], "what": w["what"], "launch_ms": launch_ms,
                    "launches_per_step": dom["launches_per_step"], "share_of_step": dom["share_of_step"],
                    "peak_source": peaks["source"] + " cuBLAS bf16 sustained (the kernel computes in "
                                   + ("tf32, nominal peak half of bf16)" if a.precision == "tf32" else
                                      "fp16 operands, fp32 accumulation: the same nominal peak as bf16)" if a.precision == "fp16" else
                                      "fp32 on CUDA cores)")}
    # the HBM-bound feature kernel is the metric's second half ("HBM GB/s as % of peak"): always reported
    for k in kernels:
        wk = work.get(k["kernel"])
        if wk and wk["bound"] == "hbm" and k["kernel"].startswith(("mfcc", "overlap_features")):
            gbs = wk["per_step"] / (k["ms_per_step"] * 1e-3) / 1e9
            extra.setdefault("feature_kernel_rooflines", []).append(
                {"kernel": k["kernel"], "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": wk["per_step"] / k["launches_per_step"],
                 "what": wk["what"]})
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            roofline["traffic"] = json.load(open(traffic_path)).get(a.workload, {}).get(roofline.get("kernel", ""))
        except Exception:
            pass

    # ---- the same step with the classifier on the fp32 CUDA-core path (the bit-faithful arithmetic) ----------
    summary = {"h2d_only_ms": round(ms_h2d, 4), "h2d_gbs_per_gpu": round(h2d / (ms_h2d * 1e-3) / 1e9, 2),
               "e2e_floor_note": "e2e >= h2d_only_ms per step: PCIe host->device upload of the step's int16 PCM"}
    labels_tc = labels_32 = None
    if pipe is not None:
        labels_tc, _ = step(pcm)
        labels_tc = labels_tc[: hi - lo].clone() if world == 1 else labels_tc[lo:hi].clone()
        if a.precision == "tf32" and not a.no_extra:
            pipe.model.set_precision("fp32")
            for _ in range(2):
                step(pcm)
            ms_32 = timed(lambda: step(pcm), max(3, a.steps // 4))
            l32, _ = step(pcm)
            labels_32 = l32[: hi - lo].clone() if world == 1 else l32[lo:hi].clone()
            pipe.model.set_precision("tf32")
            extra["fp32_classifier"] = {"value": audio_s / (ms_32 / 1e3), "unit": "audio-s/s", "ms_per_step": ms_32,
                                        "label_agreement_tf32_vs_fp32": float((labels_tc == labels_32).float().mean().item())}
            summary["fp32_classifier_audio_s_per_s"] = round(audio_s / (ms_32 / 1e3), 1)

    # ---- >= 1 s sustained loop of the timed step with its own clock record ------------------------------------
    if not a.no_extra:
        sus_sampler = ClockSampler(physical_gpu_index(local_rank))
        sus_sampler.start()
        n_sus = max(a.steps, int(1200.0 / max(ms_step, 1e-3)))
        sus_sampler.active = True
        ms_sus = timed(lambda: step(pcm), n_sus)
        sus_sampler.active = False
        sus_sampler.stop()
        extra["sustained"] = {"steps": n_sus, "seconds": ms_sus * n_sus / 1e3, "ms_per_step": ms_sus,
                              "value": audio_s / (ms_sus / 1e3), "clocks": sus_sampler.summary()}
        summary["sustained_1s_audio_s_per_s"] = round(audio_s / (ms_sus / 1e3), 1)
        summary["sustained_sm_mhz"] = sus_sampler.summary().get("sm_mhz")

    # ---- speaker features materialised ([256,39]) and classifier alone, as stages -----------------------------
    if a.workload == "speaker_id":
        feat = torch.empty((B, 256, 39), dtype=torch.float32, device="cuda")
        ms_feat = timed(lambda: si.speaker_features_batch(pcm, out=feat), a.steps)
        ms_cls = timed(lambda: pipe.model.predict_device(feat), a.steps)
        extra["stages_ms"] = {"speaker_features (mfcc_tc + mfcc_finish)": ms_feat, "speaker_classifier": ms_cls}
        fb = B * (L * 2 + 256 * 39 * 4)
        extra["speaker_features_roofline"] = {"bound": "hbm", "achieved": fb / (ms_feat * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                              "unit": "GB/s", "frac": fb / (ms_feat * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                              "algorithmic_bytes_per_step": fb}
        del feat

    def kernel_rooflines(trace, steps, ms_total, wk):
        """[{kernel, ms_per_step, share, roofline fields}] for the kernels of a traced secondary workload."""
        per2 = {}
        for name, ms in trace:
            d = per2.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += ms
        out = []
        for k, v in sorted(per2.items(), key=lambda kv: -kv[1][1]):
            ms_k = v[1] / steps
            row = {"kernel": k, "launches_per_step": v[0] / steps, "ms_per_step": ms_k, "share_of_step": ms_k / ms_total}
            w2 = wk.get(k)
            if w2 is not None:
                if w2["bound"] == "hbm":
                    g = w2["per_step"] / (ms_k * 1e-3) / 1e9
                    row.update({"bound": "hbm", "achieved": g, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": g / peaks["hbm_gbs"]})
                else:
                    t = w2["per_step"] / (ms_k * 1e-3) / 1e12
                    row.update({"bound": "tensor", "achieved": t, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                "frac": t / peaks["bf16_tflops_sustained"]})
            out.append(row)
        return out

    # ---- secondary workloads in the same line, so the driver's BENCH / SCALE records carry them at every N ------
    if a.workload == "speaker_id" and not a.no_extra:
        del pcm_dev2
        # (1) overlap path (BASELINE configs[0]/[3]): 512 clips of 1.5 s per GPU; classifier in its "fp16" mode (conv pairs with
        #     fp16 operands / fp32 accumulation, same parity bars as TF32: tests/test_r02_parity_gpu.py) and in "tf32"
        Bo, Lo = 512, 24000
        po = synth.synth_clips(20_000_000 + rank * Bo, Bo, Lo)
        po_host = torch.empty((Bo, Lo), dtype=torch.int16).pin_memory()
        po_host.copy_(po.cpu())
        po_dev = torch.empty_like(po)
        opipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision="tf32"))

        def ostep(x):
            lab, _ = opipe.run_device(x)
            cnt = tally.device_counts(lab, 2)
            if world > 1:
                lab, cnt = exchange_labels_and_counts(lab, cnt, Bo * world, rank, world)
            return lab, cnt

        def o_reduce(labels_dev, counts_dev):
            return exchange_labels_and_counts(labels_dev, counts_dev, Bo * world, rank, world)

        def ostep_e2e(n=1):
            # the public host API, as for the speaker workload: every pass uploads its own PCM from pinned host memory and
            # reads its own labels + tallies back; two passes in flight, one slice per pass (the path is compute-bound)
            from collections import deque
            pend = deque()
            for _ in range(n):
                pend.append(opipe.submit_host(po_host, 2, n_chunks=1, depth=3, reduce=o_reduce if world > 1 else None))
                if len(pend) >= 2:
                    pend.popleft().result()
            while pend:
                pend.popleft().result()

        by_prec = {}
        for prec in ("tf32", "fp16"):
            opipe.model.set_precision(prec)
            for _ in range(3):
                ostep(po)
            n_o = max(5, a.steps // 2)
            ms_o = timed(lambda: ostep(po), n_o)
            ostep_e2e(2)
            n_oe = max(4, n_o // 2)
            ms_oe = timed(lambda: ostep_e2e(n_oe), 1) / n_oe
            tr = _lib.trace_launches(lambda: [ostep(po) for _ in range(3)], torch)
            barrier()
            ok = kernel_rooflines(tr, 3, ms_o, kernel_work("overlap", Bo, Lo, opipe, 0))
            by_prec[prec] = {
                "clips_per_gpu": Bo, "ms_per_step": ms_o, "audio_s_per_s": world * Bo * Lo / SR / (ms_o * 1e-3),
                "e2e_audio_s_per_s": world * Bo * Lo / SR / (ms_oe * 1e-3), "e2e_ms_per_step": ms_oe,
                "classifier_precision": prec, "kernels": ok[:6]}
        # headline of the overlap path = the fp16-operand mode; the TF32 mode's line rides along
        extra["overlap_1.5s_x512"] = dict(by_prec["fp16"], tf32=by_prec["tf32"])
        summary["overlap_x512_audio_s_per_s"] = round(by_prec["fp16"]["audio_s_per_s"], 1)
        summary["overlap_x512_e2e_audio_s_per_s"] = round(by_prec["fp16"]["e2e_audio_s_per_s"], 1)
        summary["overlap_x512_classifier_precision"] = "fp16 operands / fp32 accumulation in the conv pairs, tf32 elsewhere"
        summary["overlap_x512_tf32_audio_s_per_s"] = round(by_prec["tf32"]["audio_s_per_s"], 1)
        summary["overlap_x512_tf32_e2e_audio_s_per_s"] = round(by_prec["tf32"]["e2e_audio_s_per_s"], 1)
        for prec in ("tf32", "fp16"):
            for row in by_prec[prec]["kernels"]:
                if row["kernel"] in ("conv_slab_kernel", "resblock2d_fused_kernel", "resblock2d_persist_kernel",
                                     "stem_resblock2d_persist_kernel", "resblock2d_f16_kernel", "stem_resblock2d_f16_kernel",
                                     "stem_resblock2d_persist_f16_kernel", "resblock2d_persist_f16_kernel",
                                     "overlap_features_kernel", "overlap_features_tc_kernel") and "frac" in row:
                    summary["overlap_%s_frac" % row["kernel"]] = round(row["frac"], 4)
        del po, po_dev, po_host, opipe

        # (2) BASELINE configs[2] at its stated size: 1 M clips of 2.5 s GLOBAL (nfilt 40, 13 cepstra), sharded over
        #     the ranks (strong scaling for this sub-number): 80 GB of int16 PCM resident at N = 1
        del pcm
        torch.cuda.empty_cache()
        n_glob = int(os.environ.get("MMLA_BENCH_BULK_CLIPS", "1000000"))
        blo, bhi = shard_range(n_glob, rank, world)
        Bb, Lb = bhi - blo, 40000
        cfgb = si.MfccConfig(nfilt=40)
        Tb = cfgb.num_frames(Lb)
        per_clip = Lb * 2 + Tb * 13 * 4
        free_b, _tot = torch.cuda.mem_get_info()
        scaled = False
        if Bb * per_clip > 0.9 * free_b:                   # never drive the box out of memory: shrink and say so
            Bb = int(0.9 * free_b // per_clip)
            scaled = True
        pb = synth.synth_clips(10_000_000 + blo, Bb, Lb)
        ob = torch.empty((Bb, Tb, 13), dtype=torch.float32, device="cuda")
        for _ in range(2):
            si.mfcc_batch(pb, cfgb, out=ob)
        ms_b = timed(lambda: si.mfcc_batch(pb, cfgb, out=ob), 5)
        nb = Bb * per_clip
        cnt_b = torch.tensor([Bb], device="cuda", dtype=torch.int64)
        if world > 1:
            dist.all_reduce(cnt_b)
        tot_b = int(cnt_b.item())
        extra["bulk_mfcc13_nfilt40_2.5s_1M"] = {
            "global_clips": tot_b, "clips_this_gpu": Bb, "scaled_down_to_fit_memory": scaled, "ms": ms_b,
            "audio_s_per_s": tot_b * Lb / SR / (ms_b * 1e-3), "scaling": "strong (1 M clips global, sharded)",
            "roofline": {"bound": "hbm", "kernel": "mfcc_tc_kernel", "achieved": nb / (ms_b * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": nb / (ms_b * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "algorithmic_bytes_per_launch": nb, "what": "int16 PCM in + float32 [249,13] cepstra out, per GPU"}}
        summary["bulk_1M_audio_s_per_s"] = round(tot_b * Lb / SR / (ms_b * 1e-3), 1)
        summary["bulk_1M_ms"] = round(ms_b, 3)
        summary["bulk_1M_hbm_frac_per_gpu"] = round(nb / (ms_b * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)
        del pb, ob
        torch.cuda.empty_cache()

        # (3) BASELINE configs[3], speaker half: ONE 8 h recording (460.8 M samples -> 2 879 999 frames -> 11 250 chunks of
        #     256 frames), whole-file MFCC + delta + delta-delta, chunks split across the ranks with a read-only halo
        #     (sharding.session_slice: no exchange on the data path), classifier, label all_gather, tallies
        from datetime import datetime as _dt
        n8h = 11250 * 40960
        rec = synth.synth_clips(30_000_000, 11250, 40960).reshape(-1)          # every rank holds the shared recording
        spk_names = {i: "spk%d" % i for i in range(10)}
        spipe = SpeakerPipeline(models.Model(W.speaker_spec(10, "sigmoid"), W.synthetic_weights(W.speaker_spec(10, "sigmoid"), 4321),
                                             precision="tf32"))
        t_fix = _dt(2021, 6, 1, 9, 0, 0, 123456)

        def session():
            return spipe.run_session_sharded(rec, spk_names, rank, world, t0=t_fix)

        for _ in range(2):
            session()
        ms_s = timed(session, 5)
        lab_s, (cnt_s, sec_s, tot_s) = session()
        extra["speaker_8h_session"] = {"audio_hours": 8.0, "chunks": int(lab_s.numel()), "ms_per_session": ms_s,
                                       "audio_s_per_s": n8h / SR / (ms_s * 1e-3), "total_seconds": tot_s,
                                       "sharding": "256-frame chunk ranges + 800 / 880-sample halo per rank, labels all_gather",
                                       "scaling": "strong (one recording)"}
        summary["speaker_8h_session_ms"] = round(ms_s, 3)
        summary["speaker_8h_session_audio_s_per_s"] = round(n8h / SR / (ms_s * 1e-3), 1)
        del rec, spipe
        torch.cuda.empty_cache()
    sampler.stop()

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only) and label agreement against the oracle -------
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sample = {"speaker_id": 1024, "bulk_mfcc": 192, "overlap": 6}[a.workload]
        v, dt, cores = time_cpu(a.workload, L, sample, steps=3, warmup=1, all_cores=True)
        cpu_baseline = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                        "sample": f"{sample} clips x 3 steps of the same workload through oracle/ "
                                  f"(numpy float64 psf/librosa restatement in single-threaded worker processes on all "
                                  f"{cores} host cores + torch-CPU fp32 classifier); upstream libs not installable here"}
    if rank == 0 and a.workload == "speaker_id" and labels_tc is not None and not a.no_cpu_baseline:
        from oracle import nets as onets
        n_chk = min(1024, hi - lo)
        pool, _cores = make_feature_pool(L)
        per_job = max(1, n_chk // pool._processes)
        x = np.concatenate(pool.map(_cpu_features_speaker, [(lo + i, min(per_job, n_chk - i), L) for i in range(0, n_chk, per_job)]))
        pool.close()
        st8 = cpu_state("speaker_id")
        ref_lab = np.argmax(onets.speaker_forward(x, st8["w"], st8["spec"]), axis=1)
        agree = {"clips": n_chk, "tf32": float((labels_tc[:n_chk].cpu().numpy() == ref_lab).mean())}
        if labels_32 is not None:
            agree["fp32"] = float((labels_32[:n_chk].cpu().numpy() == ref_lab).mean())
        extra["label_agreement_vs_oracle"] = agree
        summary["label_agreement_vs_oracle_tf32"] = round(agree["tf32"], 4)
        if "fp32" in agree:
            summary["label_agreement_vs_oracle_fp32"] = round(agree["fp32"], 4)
        summary["label_agreement_clips"] = n_chk

    if rank == 0:
        line = {
            "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if (pipe is None or a.precision == "fp32") else ("f16" if a.precision == "fp16" else "tf32"), "data": "synthetic",
            "config": {"workload": wl["name"], "classifier_precision": a.precision if pipe is not None else None,
                       "clips_per_gpu": B, "clip_seconds": L / SR, "global_clips": n_total,
                       "stream_slices": getattr(pipe, "n_streams", 1) if pipe is not None else 1,
                       "host_numa_binding": bool(numa_bound),
                       "sharding": f"clips x{world}, no data-path collective; labels all_gather + tally all_reduce",
                       "l2": "inputs larger than L2 (%.0f MB int16 PCM per GPU per step)" % (B * L * 2 / 1e6),
                       "weights": "seeded synthetic, reference shapes (real .data shards stripped from the mount)"},
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": audio_s / (ms_e2e / 1e3), "unit": "audio-s/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "bound": "pcie_h2d" if ms_e2e < 1.25 * ms_h2d else "compute", "h2d_only_ms": ms_h2d},
            "gpu_launches": gpu_launches, "clocks": sampler.summary(), "extra": extra,
            # compact digest LAST: the driver keeps the tail of the line
            "summary": summary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
