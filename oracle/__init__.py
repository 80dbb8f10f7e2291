"""CPU oracle for the mmla-audio hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy float64 / float32, torch-CPU fp32 for the
classifiers) of the arithmetic the reference's hot path delegates to third-party libraries
that are NOT vendored under /root/reference and are NOT installable here:

  * python_speech_features 0.6  (``mfcc``)        -> oracle/psf.py
  * librosa 0.8/0.9 + matplotlib ``imsave``       -> oracle/librosa_mel.py
  * TensorFlow/Keras 2.6 ``model.predict``        -> oracle/nets.py
  * the reference's own tallies                   -> oracle/tally.py
  * the synthetic PCM generator (bit-exact twin of the CUDA generator) -> oracle/synth.py

  * WebRTC VAD mode 3 (the C code ``webrtcvad`` wraps)  -> oracle/webrtc_vad.c (+ webrtc_vad.py: frame_generator /
    vad_collector / clip rewrite restated from the reference)
  * noisereduce 2.x stationary gate over scipy.signal  -> oracle/noisereduce_stationary.py
  * Keras categorical_crossentropy + RMSprop head fit  -> oracle/head_fit.py (torch autograd)

PARITY: LIBRARY ARITHMETIC UNPINNED, REFERENCE-OWN CODE PINNED.  None of the reference's libraries can be imported or
built in this image, so the restatements of THEIR arithmetic cannot be checked against reference outputs; they are
cross-checked end to end against independent implementations available here (scipy, torchaudio, torch.nn, torch
autograd) in tests/test_oracle_cpu.py.  Everything the reference computes in its OWN code (delta, binarizer,
frame_generator, vad_collector, segmentation, both visualization() tallies, normalize_matrix, and the composition
around the library calls in input_feature_gen / make_feature_experiment / OverlapFeaturesGenerator) is pinned by golden
vectors made by running the reference's modules themselves (tests/golden/make_reference_vectors.py ->
tests/golden/reference_vectors.{npz,json}; tests/test_reference_golden_cpu.py checks this package against them).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package.  The product (``mmla_audio_b200``) never does.
"""
