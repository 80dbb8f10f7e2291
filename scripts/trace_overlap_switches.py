import os, sys
sys.path.insert(0, "/root/repo")
import torch
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import OverlapPipeline
pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision="tf32"))
pcm = synth.synth_clips(0, 512, 24000)
for _ in range(2): pipe.run_device(pcm)
for env in ({}, {"MMLA_NET_FUSE_HPOOL": "0"}, {"MMLA_NET_FUSE_HPOOL": "0", "MMLA_NET_FUSE_STEM2D": "0"}):
    for k in ("MMLA_NET_FUSE_HPOOL", "MMLA_NET_FUSE_STEM2D"): os.environ.pop(k, None)
    os.environ.update(env)
    pipe.run_device(pcm)
    acc = {}
    for _ in range(5):
        tr = _lib.trace_launches(lambda: pipe.run_device(pcm), torch)
        for i, (n, ms) in enumerate(tr):
            acc[(i, n)] = acc.get((i, n), 0) + ms / 5
    print(env, " ".join(f"{n[:12]}={ms:.3f}" for (i, n), ms in sorted(acc.items())), "total", round(sum(acc.values()), 3))
