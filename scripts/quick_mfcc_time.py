"""Scratch timing of the fused MFCC kernel (CUDA events, inputs > L2). Not the bench contract."""
import sys, json, torch
sys.path.insert(0, '.')
from mmla_audio_b200 import speaker_identification as si, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
L = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
nfilt = int(sys.argv[3]) if len(sys.argv) > 3 else 40
deltas = len(sys.argv) > 4 and sys.argv[4] == '1'
pcm = synth.synth_clips(0, B, L)
cfg = si.MfccConfig(nfilt=nfilt)
pad = 256 if deltas else 0
out = si.mfcc_batch(pcm, cfg, with_deltas=deltas, pad_frames=pad)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); si.mfcc_batch(pcm, cfg, with_deltas=deltas, pad_frames=pad, out=out); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = min(ts)
T = cfg.num_frames(L)
bytes_ = B * (L * 2 + out.shape[1] * out.shape[2] * 4)
print(json.dumps(dict(B=B, L=L, nfilt=nfilt, deltas=deltas, ms=ms, all_ms=ts, clips_per_s=B / ms * 1e3,
                      frames_per_s=B * T / ms * 1e3, audio_s_per_s=B * L / 16000 / ms * 1e3,
                      GBps=bytes_ / ms / 1e6, frac_of_6441=bytes_ / ms / 1e6 / 6441)))
