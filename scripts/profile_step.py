"""One bench-sized speaker step (4096 x 1.5 s clips: PCM -> MFCC -> classifier -> labels) between
cudaProfilerStart/Stop — the target of `ncu --profile-from-start off --set full`.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import models, synth, weights as W
from mmla_audio_b200.pipeline import SpeakerPipeline

spec = W.speaker_spec(10, "sigmoid")
pipe = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision="tf32"))
pcm = synth.synth_clips(0, int(os.environ.get("CLIPS", "4096")), 24000)
for _ in range(3):
    pipe.run_device(pcm)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
pipe.run_device(pcm)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
