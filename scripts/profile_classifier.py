"""Run the speaker (or overlap) classifier a few times on random features — target for ncu."""
import sys, torch
sys.path.insert(0, '.')
from mmla_audio_b200 import models, weights as W
kind = sys.argv[1] if len(sys.argv) > 1 else 'speaker'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
prec = sys.argv[3] if len(sys.argv) > 3 else 'tf32'
g = torch.Generator(device='cuda').manual_seed(0)
if kind == 'speaker':
    spec = W.speaker_spec(10, 'sigmoid')
    x = torch.randn((B, 256, 39), device='cuda', generator=g) * 10
else:
    spec = W.OVERLAP
    x = torch.randint(0, 256, (B, 128, 151, 3), dtype=torch.uint8, device='cuda', generator=g)
m = models.Model(spec, W.synthetic_weights(spec, 1), precision=prec)
for _ in range(3):
    p, l = m.predict_device(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    m.predict_device(x)
b.record(); torch.cuda.synchronize()
print('ms per forward', a.elapsed_time(b) / 5)
