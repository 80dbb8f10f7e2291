"""clock64 timeline of one CTA of every conv_slab_kernel launch of an overlap-net step (conv_slab.cu).  Development aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import OverlapPipeline

lib = _lib.load()
pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision="tf32"))
pcm = synth.synth_clips(0, int(os.environ.get("CLIPS", "512")), 24000)
for _ in range(2):
    pipe.run_device(pcm)
torch.cuda.synchronize()
stamps = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
lib.mmla_debug_conv_slab_stamps(stamps.data_ptr(), int(os.environ.get("IMAGE", "300")))
pipe.run_device(pcm)
torch.cuda.synchronize()
lib.mmla_debug_conv_slab_stamps(None, 0)
P = stamps.cpu().numpy().reshape(64, 16)
print("launch:  setup  (f.issue f.wait f.xform) fill  fsync  w-wait  issue  mma-tail  epilogue  exit   total   (cycles)")
for i, r in enumerate(P):
    if r[0] == 0:
        continue
    d = lambda x, y: int(r[x] - r[y])
    print(f"{i:5d}: {d(1,0):7d}  ({d(9,1):6d}{d(10,9):7d}{d(2,10):8d}){d(2,1):6d}{d(3,2):7d}{d(4,3):8d}{d(5,4):7d}{d(6,5):10d}{d(7,6):10d}{d(8,7):6d}{d(8,0):8d}")
