"""Drop-in for the reference's ``overlap_features_generator`` module
(OverlapDetection/scripts/overlap_features_generator.py:29-151): same class, methods, argument
meaning and return types, computed by the fused sm_100a kernel behind ``mmla_overlap_features``.

Every method accepts a WAV path (as the reference does) or an int16 array / tensor; the batched
``*_batch`` methods are what the offline pipelines and the benchmark use.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib
from .audio_io import as_int16_signal
from .params import OVERLAP_FRAMES


class OverlapFeaturesGenerator:
    def __init__(self, wl, hl, sr=16000):
        """wl / hl: window and hop length in milliseconds (overlap_features_generator.py:31-42)."""
        self.sr = sr
        self.window_length = int(sr * (wl / 1000))
        self.hop_length = int(sr * (hl / 1000))
        self.time_dim = 150
        self.mel_dim = 128
        if (self.window_length, self.hop_length, sr) != (400, 160, 16000):
            raise _lib.MmlaError("the device kernel is built for wl=25 ms, hl=10 ms, sr=16000 "
                                 "(the only configuration the reference instantiates)")

    def get_attributes(self):
        return self.window_length, self.hop_length, self.sr

    # ------------------------------------------------------------------ batched device path
    def features_batch(self, pcm, n_mels=128, want=("image",), lengths_host=None):
        """pcm: int16 [B, L] (numpy / torch).  Returns a dict of CUDA tensors for the names in
        ``want``: 's_db', 's_db_norm' float32 [B,n_mels,151]; 'zcr' float32 [B,151];
        'image' uint8 [B,n_mels,151,3] (rows flipped, trunc(v*255) — what the classifier eats).
        ``lengths_host``: optional per-clip sample counts (ragged clips, e.g. after silence removal): clip i is
        ``pcm[i, :lengths_host[i]]``, zero-padded / truncated to 24000 like every clip (…generator.py:73-80)."""
        from .speaker_identification import _to_device_pcm
        torch = _lib.require_cuda()
        lib = _lib.load()
        x = _to_device_pcm(torch, pcm)
        if x.dim() == 1:
            x = x[None, :]
        B, L = x.shape
        dev = x.device
        out = {}
        if "s_db" in want:
            out["s_db"] = torch.empty((B, n_mels, OVERLAP_FRAMES), dtype=torch.float32, device=dev)
        if "s_db_norm" in want:
            out["s_db_norm"] = torch.empty((B, n_mels, OVERLAP_FRAMES), dtype=torch.float32, device=dev)
        if "zcr" in want:
            out["zcr"] = torch.empty((B, OVERLAP_FRAMES), dtype=torch.float32, device=dev)
        if "image" in want:
            out["image"] = torch.empty((B, n_mels, OVERLAP_FRAMES, 3), dtype=torch.uint8, device=dev)
        ptr = lambda k: out[k].data_ptr() if k in out else None
        stride0 = x.stride(0) if B > 1 else L
        off_p = len_p = None
        if lengths_host is not None:
            ln = np.ascontiguousarray(lengths_host, dtype=np.int32)
            if ln.shape != (B,) or (ln > L).any() or (ln < 0).any():
                raise ValueError("lengths_host must hold one length in [0, L] per clip")
            off = np.arange(B, dtype=np.int64) * stride0
            off_p, len_p = off.ctypes.data, ln.ctypes.data
        _lib.check(lib.mmla_overlap_features(x.data_ptr(), (B - 1) * stride0 + L, off_p, len_p, B, L, stride0,
                                             n_mels, ptr("s_db"), ptr("s_db_norm"), ptr("zcr"), ptr("image"),
                                             _lib.stream_ptr(torch)), "mmla_overlap_features")
        return out

    # ------------------------------------------------------------------ reference signatures
    def generate_mels(self, wav_file_path, n_mels=128):
        """→ (s_db, s_db_norm), float32 ``[n_mels, 151]`` each (…generator.py:65-85)."""
        _, sig = as_int16_signal(wav_file_path)
        o = self.features_batch(np.asarray(sig), n_mels, want=("s_db", "s_db_norm"))
        return o["s_db"][0].cpu().numpy(), o["s_db_norm"][0].cpu().numpy()

    def generate_zcr(self, wav_file_path):
        """→ float64 ``[1, 151]`` zero-crossing rate (…generator.py:87-101)."""
        _, sig = as_int16_signal(wav_file_path)
        o = self.features_batch(np.asarray(sig), want=("zcr",))
        # the rate is k/400 with integer k: recover the exact float64 the reference returns
        k = np.rint(o["zcr"][0].cpu().numpy().astype(np.float64) * 400.0)
        return (k / 400.0)[None, :]

    @staticmethod
    def normalize_matrix(m):
        """(m - min) / (max - min), same dtype; NaN when max == min (…generator.py:103-117)."""
        torch = _lib.require_cuda()
        a = np.asarray(m)
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        mn, mx = t.min(), t.max()
        return ((t - mn) / (mx - mn)).cpu().numpy().astype(a.dtype)

    def generate_zcr_image(self, wav_file_path, out_dir, out_name=None):
        """``out_name is None`` → float64 ``[128,151,3]`` = (zcr, 1-norm, 1-norm); otherwise writes
        ``out_dir + out_name`` as a PNG exactly as ``plt.imsave(origin='lower')`` would
        (rows flipped, trunc(v*255), opaque alpha) and returns None (…generator.py:133-151)."""
        if not os.path.isdir(out_dir):
            os.mkdir(out_dir)
        _, sig = as_int16_signal(wav_file_path)
        want = ("s_db_norm", "zcr") if out_name is None else ("image",)
        o = self.features_batch(np.asarray(sig), want=want)
        if out_name is None:
            norm = o["s_db_norm"][0].cpu().numpy()
            zcr = np.rint(o["zcr"][0].cpu().numpy().astype(np.float64) * 400.0) / 400.0
            img = np.empty((norm.shape[0], norm.shape[1], 3))
            img[:, :, 0] = zcr[None, :]
            img[:, :, 1] = 1 - norm
            img[:, :, 2] = 1 - norm
            return img
        from PIL import Image
        rgb = o["image"][0].cpu().numpy()
        rgba = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2)
        Image.fromarray(rgba, "RGBA").save(out_dir + out_name, format="PNG")
        return None

    def classifier_input_batch(self, pcm, lengths_host=None):
        """uint8 CUDA ``[B,128,151,3]`` — the tensor ``decode_png(.,3)`` yields in the reference
        (record_on_pc.py:156-158), without the PNG round trip."""
        return self.features_batch(pcm, want=("image",), lengths_host=lengths_host)["image"]
