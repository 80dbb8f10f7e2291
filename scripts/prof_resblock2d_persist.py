"""clock64 timeline of the roles of resblock2d_persist_kernel (csrc/resblock2d_persist.cu): work items 4..7 of a mid-grid CTA of
each persistent launch of an overlap-net step.  Development aid."""
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
stamps = torch.zeros(4 * 64, dtype=torch.int64, device="cuda")
lib.mmla_debug_resblock2d_persist_stamps(stamps.data_ptr())
pipe.run_device(pcm)
torch.cuda.synchronize()
lib.mmla_debug_resblock2d_persist_stamps(None)
P = stamps.cpu().numpy().reshape(4, 4, 16)
names = ["fill0", "fill1", "c1 i0", "c1 i1", "c2 i0", "c2 i1", "e1 rdy", "e1 slab", "e1 end", "e2 rdy", "e2 drn", "e2 end"]
for L in range(4):
    if P[L].max() == 0:
        continue
    t0 = P[L][P[L] > 0].min()
    print(f"launch {L}: " + " ".join(f"{n:>8s}" for n in names))
    for k in range(4):
        print(f"  item {k + 4}: " + " ".join(f"{int(P[L, k, s] - t0):8d}" for s in range(12)))
    d = P[L, 3] - P[L, 1]
    print(f"  period (items 5 -> 7, per item): {int((P[L, 3, 3] - P[L, 1, 3]) / 2)} cycles; "
          f"fill {int(P[L,2,1]-P[L,2,0])}, conv1 issue {int(P[L,2,3]-P[L,2,2])}, conv2 issue {int(P[L,2,5]-P[L,2,4])}, "
          f"e1 {int(P[L,2,8]-P[L,2,6])} (slab wait {int(P[L,2,7]-P[L,2,6])}), e2 {int(P[L,2,11]-P[L,2,9])}")
