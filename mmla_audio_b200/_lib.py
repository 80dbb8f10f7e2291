"""ctypes binding of libmmla_b200.so (the C-ABI declared in include/mmla_b200.h).

Fails loudly: a missing library raises ImportError with the build command; a failing call
raises MmlaError carrying mmla_last_error().  No fallback path exists.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmmla_b200.so")

# every symbol include/mmla_b200.h declares (checked by tests/test_host_cpu.py::test_library_exports_every_declared_symbol)
SYMBOLS = (
    "mmla_last_error", "mmla_abi_version", "mmla_launch_count", "mmla_crc32c_host",
    "mmla_psf_num_frames", "mmla_psf_mfcc", "mmla_psf_mfcc_rows", "mmla_delta", "mmla_cmvn", "mmla_overlap_features",
    "mmla_net_create", "mmla_net_destroy", "mmla_net_set_precision", "mmla_net_workspace_bytes",
    "mmla_net_forward", "mmla_net_forward_cepstra", "mmla_net_embed", "mmla_head_fit",
    "mmla_tally", "mmla_synth_pcm", "mmla_vad_num_frames", "mmla_vad_trim", "mmla_noise_profile", "mmla_noise_gate", "mmla_debug_mfcc_tc_dump", "mmla_debug_resstage_stamps", "mmla_debug_lstm_stamps", "mmla_debug_conv2d", "mmla_debug_resblock2d", "mmla_debug_resblock2d_stamps", "mmla_debug_resblock2d_persist_stamps", "mmla_debug_conv_slab_stamps", "mmla_trace_begin", "mmla_trace_end",
)


class MmlaError(RuntimeError):
    pass


class MfccParams(C.Structure):
    """Mirror of ``MmlaMfccParams``; defaults = the reference call
    ``mfcc(sig, rate, winlen=0.025, winstep=0.01, nfft=512)``."""
    _fields_ = [
        ("samplerate", C.c_int32), ("frame_len", C.c_int32), ("frame_step", C.c_int32),
        ("nfft", C.c_int32), ("nfilt", C.c_int32), ("numcep", C.c_int32),
        ("ceplifter", C.c_int32), ("append_energy", C.c_int32), ("window", C.c_int32),
        ("with_deltas", C.c_int32), ("pad_frames", C.c_int32),
        ("preemph", C.c_float), ("lowfreq", C.c_float), ("highfreq", C.c_float),
    ]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m mmla_audio_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
    mp = C.POINTER(MfccParams)
    sigs = {
        "mmla_last_error": (C.c_char_p, []),
        "mmla_abi_version": (C.c_int, []),
        "mmla_launch_count": (i64, []),
        "mmla_crc32c_host": (u32, [vp, C.c_size_t]),
        "mmla_psf_num_frames": (i32, [i64, mp]),
        "mmla_psf_mfcc": (C.c_int, [vp, i64, vp, vp, i64, i32, i64, mp, vp, i64, vp]),
        "mmla_psf_mfcc_rows": (C.c_int, [vp, i64, vp, vp, i64, i32, i64, mp, vp, i64, i32, vp]),
        "mmla_delta": (C.c_int, [vp, i64, i32, i32, vp, vp]),
        "mmla_cmvn": (C.c_int, [vp, i64, i64, i32, i32, i32, vp, i32, vp]),
        "mmla_overlap_features": (C.c_int, [vp, i64, vp, vp, i64, i32, i64, i32, vp, vp, vp, vp, vp]),
        "mmla_net_create": (C.c_int, [i32, i32, i32, vp, i64, C.POINTER(vp)]),
        "mmla_net_destroy": (None, [vp]),
        "mmla_net_set_precision": (C.c_int, [vp, i32]),
        "mmla_net_workspace_bytes": (i64, [vp, i64]),
        "mmla_net_forward": (C.c_int, [vp, vp, i32, i64, vp, i64, vp, vp, vp]),
        "mmla_net_forward_cepstra": (C.c_int, [vp, vp, i64, i32, i64, vp, i64, vp, vp, vp]),
        "mmla_net_embed": (C.c_int, [vp, vp, i32, i64, vp, i64, vp, vp]),
        "mmla_head_fit": (C.c_int, [vp, vp, i64, i32, vp, i32, i32, C.c_float, C.c_float, C.c_float, vp, vp, vp, vp]),
        "mmla_tally": (C.c_int, [vp, i64, i32, vp, vp]),
        "mmla_noise_profile": (C.c_int, [vp, i64, C.c_float, vp, vp]),
        "mmla_noise_gate": (C.c_int, [vp, i64, i32, i64, vp, vp, vp, i64, vp]),
        "mmla_vad_num_frames": (i32, [i32]),
        "mmla_vad_trim": (C.c_int, [vp, i64, i32, i64, vp, i64, vp, vp, i32, vp, vp, i64, vp]),
        "mmla_synth_pcm": (C.c_int, [vp, i64, i64, i32, i64, u32, vp, vp]),
        "mmla_debug_mfcc_tc_dump": (None, [vp, vp]),
        "mmla_debug_resstage_stamps": (None, [vp, i32]),
        "mmla_debug_lstm_stamps": (None, [vp, i32]),
        "mmla_debug_conv_slab_stamps": (None, [vp, i32]),
        "mmla_debug_conv2d": (C.c_int, [vp, vp, vp, vp, vp, i32, vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, vp]),
        "mmla_debug_resblock2d": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
        "mmla_debug_resblock2d_stamps": (None, [vp, i32]),
        "mmla_debug_resblock2d_persist_stamps": (None, [vp]),
        "mmla_trace_begin": (C.c_int, [vp]),
        "mmla_trace_end": (C.c_int, [vp, i64, vp, i32]),
    }
    assert set(sigs) == set(SYMBOLS)
    for name, (res, args) in sigs.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:     # a stale / partial build: fail loudly, never fall back
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild with "
                              "`python -m mmla_audio_b200.build --force`") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mmla_last_error().decode("utf-8", "replace")
        raise MmlaError(f"{what} failed (code {rc}): {msg}")


def require_cuda():
    """Import torch and insist on a CUDA device — the product path never runs on the CPU."""
    import torch
    if not torch.cuda.is_available():
        raise MmlaError("mmla_audio_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def stream_ptr(torch) -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def trace_launches(fn, torch):
    """Run ``fn()`` with the library's launch trace open on the current stream; returns
    ``[(kernel_name, ms), ...]`` in launch order (device time per launch, CUDA events)."""
    lib = load()
    check(lib.mmla_trace_begin(stream_ptr(torch)), "mmla_trace_begin")
    try:
        fn()
    finally:
        names = C.create_string_buffer(1 << 16)
        ms = (C.c_float * 4096)()
        n = lib.mmla_trace_end(names, len(names), ms, 4096)
    if n < 0:
        check(n, "mmla_trace_end")
    nm = names.value.decode().split("\n")[:n]
    return [(nm[i] if i < len(nm) else "?", float(ms[i])) for i in range(n)]
