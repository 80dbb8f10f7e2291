"""CPU oracle for the mmla-audio hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy float64 / float32, torch-CPU fp32 for the
classifiers) of the arithmetic the reference's hot path delegates to third-party libraries
that are NOT vendored under /root/reference and are NOT installable here:

  * python_speech_features 0.6  (``mfcc``)        -> oracle/psf.py
  * librosa 0.8/0.9 + matplotlib ``imsave``       -> oracle/librosa_mel.py
  * TensorFlow/Keras 2.6 ``model.predict``        -> oracle/nets.py
  * the reference's own tallies                   -> oracle/tally.py
  * the synthetic PCM generator (bit-exact twin of the CUDA generator) -> oracle/synth.py

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 and none of its libraries can be imported or built in this image, so the
oracle cannot be checked against reference outputs.  It is instead cross-checked against
independent implementations available here (scipy.fft / scipy.fftpack, torchaudio's slaney
filterbank and ``compute_deltas``, torch.nn conv/LSTM) in tests/test_oracle_*.py, and pinned
by the reference's shape contracts (151 frames, [128,151,3], [256,39], weight shapes).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package.  The product (``mmla_audio_b200``) never does.
"""
