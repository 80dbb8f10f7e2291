"""Regenerate the committed fixtures in tests/golden/.  Run HERE (the authoring container):

    python tests/golden/make_golden.py

* ``index_*.json`` — names / dtypes / shapes / offsets decoded from the reference's own
  ``variables.index`` files under /root/reference (read with mmla_audio_b200.tf_bundle); they pin
  the weight-layout spec in mmla_audio_b200/weights.py to the reference artefacts.  The GPU box
  has no /root/reference, so the decoded JSON is what travels.
* ``oracle_vectors.npz`` — outputs of the oracle on seeded synthetic clips.  PARITY UNPINNED:
  the reference ships no golden vectors and its libraries cannot run here, so these pin the
  oracle against regressions only, not against the reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from mmla_audio_b200 import tf_bundle  # noqa: E402
from oracle import librosa_mel as lm, psf, synth  # noqa: E402

REF = "/root/reference"
INDEXES = {
    "index_overlap_timit2.json": "OverlapDetection/timit/models/timit2.0/variables/variables.index",
    "index_overlap_timit1.json": "OverlapDetection/timit/models/timit1.0/variables/variables.index",
    "index_speaker_timit.json": "SpeakerIdentification/timit/model/variables/variables.index",
}


def main():
    for out, rel in INDEXES.items():
        header, entries = tf_bundle.read_index(os.path.join(REF, rel))
        doc = {"source": rel, "header": header,
               "entries": [{"key": e.key, "dtype": e.dtype, "shape": list(e.shape), "offset": e.offset,
                            "size": e.size} for e in entries if "OPTIMIZER_SLOT" not in e.key
                           and not e.key.startswith("optimizer/")]}
        with open(os.path.join(HERE, out), "w") as f:
            json.dump(doc, f, indent=0)
    clips = synth.synth_clips(0, 2, 8000)
    vec = {"pcm": clips}
    vec["mfcc13"] = psf.mfcc(clips[0], 16000, winlen=0.025, winstep=0.01, nfft=512)
    vec["mfcc13_nfilt40"] = psf.mfcc(clips[1], 16000, winlen=0.025, winstep=0.01, nfft=512, nfilt=40)
    vec["feat39"] = psf.input_feature_gen(clips[0])[0][:60]
    s_db, norm = lm.generate_mels(clips[1])
    vec["s_db_cols"] = s_db[:, ::10]
    vec["zcr"] = lm.generate_zcr(clips[1])
    vec["image_cols"] = lm.imsave_rgb_uint8(lm.generate_zcr_image(clips[1]))[:, ::10]
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **vec)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
