"""One short run of the overlap feature kernel for `ncu --set full --import-source on -k regex:overlap_features`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmla_audio_b200 import synth
from mmla_audio_b200.overlap_features_generator import OverlapFeaturesGenerator
pcm = synth.synth_clips(0, int(os.environ.get("CLIPS", "512")), 24000)
ofg = OverlapFeaturesGenerator(25, 10)
for _ in range(3):
    img = ofg.classifier_input_batch(pcm)
torch.cuda.synchronize()
print("ok", int(img[0, 0, 0, 0]))
