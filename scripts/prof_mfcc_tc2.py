"""Per-role clock64 timeline of mfcc_tc2_kernel (CTA 0; stamps are taken when a stamp buffer is registered)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMLA_MFCC_TC"] = "2"
import torch
from mmla_audio_b200 import _lib, synth
from mmla_audio_b200 import speaker_identification as si

lib = _lib.load()
n_clips, L, nfilt = 4096, 24000, int(os.environ.get("NFILT", "26"))
pcm = synth.synth_clips(0, n_clips, L)
cfg = si.MfccConfig(nfilt=nfilt)
T = cfg.num_frames(L)
stamps = torch.zeros(64 * 32, dtype=torch.int64, device="cuda")
out = torch.empty((n_clips, T, 13), dtype=torch.float32, device="cuda")
si.mfcc_batch(pcm, cfg, out=out)
lib.mmla_debug_mfcc_tc_dump(None, stamps.data_ptr())
si.mfcc_batch(pcm, cfg, out=out)
torch.cuda.synchronize()
lib.mmla_debug_mfcc_tc_dump(None, None)
P = stamps.cpu().numpy().reshape(64, 32)
t0 = P[0][P[0] > 0].min()
names = {0: "tma g0", 1: "mma g0", 2: "mma g1", 3: "mma g2", 4: "mma g3", 8: "S0s", 9: "S0e", 10: "S1s", 11: "S1e", 12: "S2s", 13: "S2e",
         14: "S3s", 15: "S3e", 16: "C0s", 17: "C0e", 18: "C1s", 19: "C1e", 20: "C2s", 21: "C2e", 22: "C3s", 23: "C3e",
         24: "E0 start", 25: "E0 fft end", 26: "E0 end", 29: "E1 fft end", 30: "E1 end", 5: "barA", 6: "barB", 7: "barC", 28: "dct done"}
for it in range(20, 23):
    print(f"tile {it}: " + "  ".join(f"{names[k]}={P[it][k] - t0}" for k in sorted(names) if P[it][k] > 0))
med = lambda x: int(np.median(np.asarray(x)[8:40]))
print("period:", med(np.diff(P[:48, 26])))
print("E0: wait->fft end", med(P[:48, 25] - P[:48, 24]), " fft end->end", med(P[:48, 26] - P[:48, 25]))
print("finish: melend->barA", med(P[:48,5]-P[:48,25]), " A->B", med(P[:48,6]-P[:48,5]), " B->C", med(P[:48,7]-P[:48,6]), " C->dct done", med(P[:48,28]-P[:48,7]), " ->end", med(P[:48,26]-P[:48,28]))
print("C per group:", [med(P[:48, 17 + 2 * g] - P[:48, 16 + 2 * g]) for g in range(4)])
print("S per group:", [med(P[:48, 9 + 2 * g] - P[:48, 8 + 2 * g]) for g in range(4)])
print("mma g0->g3:", med(P[:48, 4] - P[:48, 1]))
print("E0 fft loop:", med(P[:48, 27] - P[:48, 24]), " mel pass:", med(P[:48, 25] - P[:48, 27]),)
