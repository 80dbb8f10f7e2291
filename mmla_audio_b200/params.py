"""Feature parameters whose defaults equal the reference's module-level constants
(SURVEY.md §5 "Config / flags"): framerate 16000, winlen 0.025, winstep 0.01, nfft 512,
psf defaults nfilt 26 / numcep 13 / preemph 0.97 / ceplifter 22 / appendEnergy, rectangular
window; overlap path wl=25 ms, hl=10 ms, 150 hops, 128 mels."""
from __future__ import annotations

import decimal
from dataclasses import dataclass

WINDOW_IDS = {"rect": 0, "hann": 1, "hamming": 2}


def round_half_up(x) -> int:
    return int(decimal.Decimal(x).quantize(decimal.Decimal("1"), rounding=decimal.ROUND_HALF_UP))


@dataclass(frozen=True)
class MfccConfig:
    samplerate: int = 16000
    winlen: float = 0.025
    winstep: float = 0.01
    numcep: int = 13
    nfilt: int = 26
    nfft: int = 512
    lowfreq: float = 0.0
    highfreq: float | None = None
    preemph: float = 0.97
    ceplifter: int = 22
    appendEnergy: bool = True
    window: str = "rect"

    @property
    def frame_len(self) -> int:
        return round_half_up(self.winlen * self.samplerate)

    @property
    def frame_step(self) -> int:
        return round_half_up(self.winstep * self.samplerate)

    def num_frames(self, n_samples: int) -> int:
        fl, fs = self.frame_len, self.frame_step
        if n_samples <= fl:
            return 1
        return 1 + -(-(n_samples - fl) // fs)


SILENT_MIN_SAMPLES = 4000          # `len(sig) < 4000` => 'silent'  (speaker_identification.py:375)
SPEAKER_FRAMES = 256               # pad / truncate rows           (speaker_identification.py:391-395)
OVERLAP_CLIP_SAMPLES = 24000       # hop*150                       (overlap_features_generator.py:73-80)
OVERLAP_FRAMES = 151
OVERLAP_MELS = 128
