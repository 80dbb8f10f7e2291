import os, sys
sys.path.insert(0, "/root/repo")
import torch
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import OverlapPipeline
lib = _lib.load()
pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision=__import__("os").environ.get("PRECISION", "tf32")))
pcm = synth.synth_clips(0, 512, 24000)
for _ in range(2): pipe.run_device(pcm)
torch.cuda.synchronize()
T = 19
stamps = torch.zeros(T * 16, dtype=torch.int64, device="cuda")
lib.mmla_debug_lstm_stamps(stamps.data_ptr(), int(os.environ.get("CTA", "2")))
pipe.run_device(pcm)
torch.cuda.synchronize()
lib.mmla_debug_lstm_stamps(None, 0)
P = stamps.cpu().numpy().reshape(T, 16)
t0 = P[P > 0].min()
names = ["step start", "xp prefetch issued", "pass0 acc ready", "pass0 cells done", "pass1 acc ready", "pass1 cells done",
         "peer handshake done", "h restaged", "MMA: h ready", "MMA: pass0 issued", "MMA: pass1 issued"]
for st in (3, 4):
    ev = sorted((int(P[st][k] - t0), names[k]) for k in range(len(names)) if P[st][k] > 0)
    print(f"--- step {st}")
    prev = None
    for v, n in ev:
        print(f"  {v:8d}  (+{0 if prev is None else v - prev:6d})  {n}")
        prev = v
print("step period", int(P[5][0]-P[4][0]))
