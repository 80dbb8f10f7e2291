"""One short run of the tensor-core MFCC kernel for `ncu --set full --import-source on -k regex:mfcc_tc_kernel`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import synth
from mmla_audio_b200 import speaker_identification as si
n_clips, L = int(os.environ.get("CLIPS", "8192")), 40000
pcm = synth.synth_clips(0, n_clips, L)
cfg = si.MfccConfig(nfilt=int(os.environ.get("NFILT", "40")))
out = torch.empty((n_clips, cfg.num_frames(L), 13), dtype=torch.float32, device="cuda")
for _ in range(3):
    si.mfcc_batch(pcm, cfg, out=out)
torch.cuda.synchronize()
print("ok", float(out[0, 0, 0]))
