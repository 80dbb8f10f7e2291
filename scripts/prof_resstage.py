"""clock64 timeline of one CTA of each fused ResNet-stage launch (resstage_fused.cu).  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import SpeakerPipeline

lib = _lib.load()
spec = W.speaker_spec(10, "sigmoid")
pipe = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision="tf32"))
pcm = synth.synth_clips(0, int(os.environ.get("CLIPS", "4096")), 24000)
for _ in range(3):
    pipe.run_device(pcm)
torch.cuda.synchronize()
cta = int(os.environ.get("CTA", "400"))
stamps = torch.zeros(3 * 64, dtype=torch.int64, device="cuda")
lib.mmla_debug_resstage_stamps(stamps.data_ptr(), cta)
pipe.run_device(pcm)
torch.cuda.synchronize()
lib.mmla_debug_resstage_stamps(None, 0)
P = stamps.cpu().numpy().reshape(3, 64)
names = {0: "mma start", 1: "shortcut issued", 16: "epi start", 30: "staged", 31: "stored"}
for u in range(3):
    names[2 + 4 * u] = f"u{u} a_ready0"
    names[3 + 4 * u] = f"u{u} conv1 issued"
    names[4 + 4 * u] = f"u{u} a_ready1"
    names[5 + 4 * u] = f"u{u} conv2 issued"
    names[18 + 4 * u] = f"u{u} tfull0"
    names[19 + 4 * u] = f"u{u} epi1 done"
    names[20 + 4 * u] = f"u{u} tfull1"
    names[21 + 4 * u] = f"u{u} epi2 done"
for j in range(4):
    names[32 + j] = f"tile{j} operands written"
    names[40 + j] = f"stem: clip{j} features built"
names[44] = "stem: all MMAs retired (seen by epilogue)"
for st in range(3):
    row = P[st]
    t0 = row[row > 0].min()
    print(f"--- stage {st + 1} (CTA {cta}) cycles from CTA start")
    if os.environ.get("BRIEF"):
        d = lambda x, y: int(row[x] - row[y])
        print("  load", d(2, 0), " conv", [d(3 + 2 * i, 2 + 2 * i) for i in range(6)], " epi", [d(19 + 2 * i, 18 + 2 * i) for i in range(5)],
              " stage+store", d(31, 28), " total", d(31, 0))
        continue
    for k, v in sorted(((k, row[k] - t0) for k in names if row[k] > 0), key=lambda kv: kv[1]):
        print(f"  {v:8d}  {names[k]}")
