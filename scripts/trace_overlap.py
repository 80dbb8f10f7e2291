import sys; sys.path.insert(0,'.')
import torch
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import OverlapPipeline
pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision=__import__("os").environ.get("PRECISION", "tf32")))
pcm = synth.synth_clips(0, 512, 24000)
for _ in range(2): pipe.run_device(pcm)
torch.cuda.synchronize()
tr = _lib.trace_launches(lambda: pipe.run_device(pcm), torch)
tot = sum(ms for _, ms in tr)
for n, ms in tr: print(f"{n:28s} {ms:8.4f}")
print("total", tot)
