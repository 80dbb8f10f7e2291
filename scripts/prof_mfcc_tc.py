"""Per-role clock64 timeline of the tensor-core MFCC kernel (DBG build, CTA 0).  Development aid."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import _lib, synth
from mmla_audio_b200 import speaker_identification as si

lib = _lib.load()
n_clips, L, nfilt = 4736, 40000, int(os.environ.get("NFILT", "40"))
pcm = synth.synth_clips(0, n_clips, L)
cfg = si.MfccConfig(nfilt=nfilt)
T = cfg.num_frames(L)
gpc = (T + 15) // 16
n_tiles = (n_clips * gpc + 3) // 4
stamps = torch.zeros(64 * 32, dtype=torch.int64, device="cuda")
out = torch.empty((n_clips, T, 13), dtype=torch.float32, device="cuda")
si.mfcc_batch(pcm, cfg, out=out)          # warm (non-debug kernel)
lib.mmla_debug_mfcc_tc_dump(None, stamps.data_ptr())
si.mfcc_batch(pcm, cfg, out=out)
torch.cuda.synchronize()
lib.mmla_debug_mfcc_tc_dump(None, None)
prof = stamps.cpu().numpy().reshape(64, 32)
names = {0: "prod g0", 1: "mma s1 g0", 2: "mma s1 g3", 3: "mma s2 start", 4: "mma s2 issued",
         8: "S g0 start", 9: "S g0 end", 14: "S g3 start", 15: "S g3 end",
         16: "C g0 start", 17: "C g0 end", 22: "C g3 start", 23: "C g3 end", 24: "E start", 26: "E end w0", 27: "E end w3"}
t0 = prof[0][prof[0] > 0].min()
my_tiles = (n_tiles + 147) // 148
my_tiles = min(my_tiles, 64)
for it in range(20, 24):
    row = prof[it]
    print(f"tile {it}: " + "  ".join(f"{names[k]}={row[k] - t0}" for k in sorted(names) if row[k] > 0))
def med(x):
    x = np.asarray(x)[8:]
    return int(np.median(x))
P = prof[:my_tiles]
print("period            :", med(np.diff(P[:, 26])))
print("C tile (g0 start -> g3 end):", med(P[:, 23] - P[:, 16]), " per group:", [med(P[:, 17 + 2 * g] - P[:, 16 + 2 * g]) for g in range(4)])
print("C wait before g:", [med(P[:, 16 + 2 * g] - (P[:, 15 + 2 * g] if g else P[:, 16])) for g in range(4)])
print("E tile            :", med(P[:, 26] - P[:, 24]))
print("S tile            :", med(P[:, 15] - P[:, 8]), " per group:", [med(P[:, 9 + 2 * g] - P[:, 8 + 2 * g]) for g in range(4)])
print("mma s1 g0->g3     :", med(P[:, 2] - P[:, 1]), " s2 start->issued:", med(P[:, 4] - P[:, 3]))
print("C end -> s2 start :", med(P[:, 3] - P[:, 23]), " s2 issued -> E start:", med(P[:, 24] - P[:, 4]))
print("E end -> next s2 start:", med(P[1:, 3] - P[:-1, 26]))
print("s2 issued(it) -> C g0 start(it+1):", med(P[1:, 16] - P[:-1, 4]))
