"""clock64 timeline of one CTA of the fused BiLSTM launch (lstm_fused.cu), per time step.  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import SpeakerPipeline

lib = _lib.load()
spec = W.speaker_spec(10, "sigmoid")
pipe = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision=__import__("os").environ.get("PRECISION", "tf32")))
pcm = synth.synth_clips(0, int(os.environ.get("CLIPS", "4096")), 24000)
for _ in range(3):
    pipe.run_device(pcm)
torch.cuda.synchronize()
T = 8
stamps = torch.zeros(T * 16, dtype=torch.int64, device="cuda")
lib.mmla_debug_lstm_stamps(stamps.data_ptr(), int(os.environ.get("CTA", "20")))
pipe.run_device(pcm)
torch.cuda.synchronize()
lib.mmla_debug_lstm_stamps(None, 0)
P = stamps.cpu().numpy().reshape(T, 16)
t0 = P[P > 0].min()
names = ["step start", "xp prefetch issued", "pass0 acc ready", "pass0 cells done", "pass1 acc ready", "pass1 cells done",
         "peer handshake done", "h restaged", "MMA: h ready", "MMA: pass0 issued", "MMA: pass1 issued"]
for st in range(T):
    ev = sorted((int(P[st][k] - t0), names[k]) for k in range(len(names)) if P[st][k] > 0)
    print(f"--- step {st}")
    prev = None
    for v, n in ev:
        print(f"  {v:8d}  (+{0 if prev is None else v - prev:6d})  {n}")
        prev = v
