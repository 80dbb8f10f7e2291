"""One overlap step (512 x 1.5 s clips: PCM -> image -> classifier -> labels) between cudaProfilerStart/Stop — the target of
`ncu --profile-from-start off --set full`.  PRECISION=fp16|tf32 (default fp16).  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import models, synth, weights as W
from mmla_audio_b200.pipeline import OverlapPipeline

pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision=os.environ.get("PRECISION", "fp16")))
pcm = synth.synth_clips(0, int(os.environ.get("CLIPS", "512")), 24000)
for _ in range(3):
    pipe.run_device(pcm)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
pipe.run_device(pcm)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
