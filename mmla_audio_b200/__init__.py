"""mmla_audio_b200 — B200-native (sm_100a) drop-in for the analytics hot path of
lizaibeim/mmla-audio: fixed-window 16 kHz feature extraction (python_speech_features-style
MFCC+delta for speaker-ID, librosa-style log-mel+ZCR image for overlap detection), the
ResNet-BiLSTM classifier forward passes, arg-max labels and label tallies.

Python keeps the reference's call signatures (module names ``speaker_identification`` and
``overlap_features_generator``); all arithmetic runs in hand-written CUDA kernels reached through
the C-ABI in ``include/mmla_b200.h`` (``libmmla_b200.so``).  There is no CPU fallback: importing
a compute module without the built library, or calling it without a CUDA device, raises.
"""
__version__ = "0.1.0"
