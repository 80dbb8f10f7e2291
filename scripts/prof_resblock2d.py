"""Per-launch device times and the clock64 timeline of one CTA of every resblock2d_fused_kernel launch of an overlap-net
step (csrc/resblock2d_fused.cu).  Development aid: `MMLA_RB_TILES / MMLA_RB_KB / MMLA_RB_STAGES` force a configuration,
`BLOCK=<i>` restricts the forcing to nothing (the env applies to every launch) — run once per candidate and compare."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import OverlapPipeline

lib = _lib.load()
pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision=os.environ.get("PRECISION", "tf32")))
pcm = synth.synth_clips(0, int(os.environ.get("CLIPS", "512")), 24000)
for _ in range(2):
    pipe.run_device(pcm)
torch.cuda.synchronize()
acc = {}
reps = int(os.environ.get("REPS", "5"))
for _ in range(reps):
    tr = _lib.trace_launches(lambda: pipe.run_device(pcm), torch)
    i = 0
    for n, ms in tr:
        if n in ("resblock2d_persist_f16_kernel", "stem_resblock2d_persist_f16_kernel", "resblock2d_f16_kernel", "stem_resblock2d_f16_kernel", "resblock2d_fused_kernel", "stem_resblock2d_fused_kernel", "resblock2d_persist_kernel", "stem_resblock2d_persist_kernel", "conv_slab_kernel", "pool_shortcut_kernel", "stem1x1_kernel"):
            acc[(i, n)] = acc.get((i, n), 0.0) + ms / reps
            i += 1
tot = {}
for (i, n), ms in sorted(acc.items()):
    tot[n] = tot.get(n, 0.0) + ms
print(" ".join(f"{ms:.4f}" for (_, n), ms in sorted(acc.items()) if n not in ("pool_shortcut_kernel", "stem1x1_kernel")))
print("totals:", {k: round(v, 4) for k, v in tot.items()}, "step", round(sum(ms for _, ms in tr), 4))
if os.environ.get("STAMPS", "1") != "0":
    stamps = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
    lib.mmla_debug_resblock2d_stamps(stamps.data_ptr(), int(os.environ.get("IMAGE", "300")))
    pipe.run_device(pcm)
    torch.cuda.synchronize()
    lib.mmla_debug_resblock2d_stamps(None, 0)
    P = stamps.cpu().numpy().reshape(16, 16)
    print("launch:  setup   fill  fsync  issue1  mma1-tail    epi1  issue2  mma2-tail    epi2   exit   total   (cycles)")
    for i, r in enumerate(P):
        if r[0] == 0:
            continue
        d = lambda x, y: int(r[x] - r[y])
        print(f"{i:5d}: {d(1,0):7d}{d(2,1):7d}{d(3,2):7d}{d(4,3):8d}{d(5,4):11d}{d(6,5):8d}{d(7,6):8d}{d(8,7):11d}{d(9,8):8d}{d(10,9):7d}{d(10,0):8d}")
