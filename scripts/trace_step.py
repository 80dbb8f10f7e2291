import sys; sys.path.insert(0,'.')
import torch, collections
from mmla_audio_b200 import _lib, models, synth, weights as W
from mmla_audio_b200.pipeline import SpeakerPipeline
spec = W.speaker_spec(10, "sigmoid")
pipe = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision="tf32"))
pcm = synth.synth_clips(0, 4096, 24000)
for _ in range(3): pipe.run_device(pcm)
torch.cuda.synchronize()
tr = _lib.trace_launches(lambda: [pipe.run_device(pcm) for _ in range(5)], torch)
agg = collections.OrderedDict()
for n, ms in tr: agg.setdefault(n, []).append(ms)
tot = sum(ms for _, ms in tr)
for n, v in agg.items(): print(f"{n:28s} n/step={len(v)/5:4.1f} ms/step={sum(v)/5:8.4f} share={sum(v)/tot*100:5.1f}%")
print("total ms/step", tot/5)
for n, ms in tr[:25]: print(n, round(ms,4))
