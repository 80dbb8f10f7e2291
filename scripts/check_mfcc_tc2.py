"""Scratch check of the two tensor-core MFCC formulations against each other and against the float64 oracle, plus
CUDA-event timings of both (inputs > L2).  Not the bench contract.
    python scripts/check_mfcc_tc2.py [clips] [samples] [nfilt]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from mmla_audio_b200 import speaker_identification as si, synth  # noqa: E402
from oracle import psf  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 24000
nfilt = int(sys.argv[3]) if len(sys.argv) > 3 else 26
pcm = synth.synth_clips(0, B, L)
cfg = si.MfccConfig(nfilt=nfilt)


def run(form):
    if form:
        os.environ["MMLA_MFCC_TC"] = form
    else:
        os.environ.pop("MMLA_MFCC_TC", None)
    out = si.mfcc_batch(pcm, cfg)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        si.mfcc_batch(pcm, cfg, out=out)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return out.clone(), min(ts)


o1, ms1 = run(None)
o2, ms2 = run("2")
d = (o1 - o2).abs()
host = pcm[:8].cpu().numpy()
ref = np.stack([psf.mfcc(host[i].astype(np.float64), 16000, 0.025, 0.01, 13, nfilt, 512) for i in range(8)])
res = {}
for name, o in (("tc1", o1), ("tc2", o2)):
    g = o[:8].cpu().numpy().astype(np.float64)
    tol = 1e-4 * np.abs(ref) + 1e-4 * np.abs(ref).max()
    res[name + "_worst_err_over_tol"] = float((np.abs(g - ref) / tol).max())
    res[name + "_finite"] = bool(torch.isfinite(o).all().item())
T = o1.shape[1]
bytes_ = B * (L * 2 + T * 13 * 4)
print(json.dumps(dict(B=B, L=L, nfilt=nfilt, frames=T, max_abs_diff_tc1_tc2=float(d.max()), ref_absmax=float(np.abs(ref).max()),
                      ms_tc1=ms1, ms_tc2=ms2, frac_tc1=bytes_ / ms1 / 1e6 / 6441, frac_tc2=bytes_ / ms2 / 1e6 / 6441, **res)))
