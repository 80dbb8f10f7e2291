"""Classifier loading and forward pass — the reference's ``tf.keras.models.load_model(dir)`` /
``model.predict(x)`` pair (OverlapDetection/scripts/record_on_pc.py:87-88,159;
SpeakerIdentification/scripts/record_on_pc.py:76-77,136;
overlap_detection_post_processing.py:154,208; speaker_identification_post_processing.py:206,272)
on hand-written sm_100a kernels via ``mmla_net_*``.

``load_model(dir)`` reads the SavedModel's ``variables/variables.{index,data-*}`` tensor bundle
(no TensorFlow needed).  The reference's own ``.data`` shards are stripped from its mount, so
``save_synthetic_model`` writes seeded weights with the reference's exact names and shapes in
the same format; real weights drop in unchanged once supplied.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

from . import _lib
from . import tf_bundle
from .weights import (NetSpec, OVERLAP, SPEAKER_BASE, speaker_spec, weight_shapes, check_weights,
                      synthetic_weights, dense_keys, lw, resolve_lstm_keys)

KIND_OVERLAP, KIND_SPEAKER = 0, 1
HEAD_IDS = {"softmax": 0, "sigmoid": 1}
PRECISION_IDS = {"fp32": 0, "tf32": 1, "fp16": 2}


def pack_weights(spec: NetSpec, w: Dict[str, np.ndarray]) -> np.ndarray:
    """Flatten weights in the traversal order ``mmla_net_create`` expects (csrc/nets.cu):
    stem(kernel,bias); per block: bn1(gamma,beta,mean,var), conv1, bn2, conv2, [shortcut];
    [final bn]; lstm fwd(kernel,recurrent,bias), lstm bwd; dense(kernel,bias)."""
    check_weights(spec, w)
    parts = []

    def conv(c):
        parts.append(w[lw(c.idx, "kernel")])
        parts.append(w[lw(c.idx, "bias")])

    def bn(i):
        for n in ("gamma", "beta", "moving_mean", "moving_variance"):
            parts.append(w[lw(i, n)])

    conv(spec.stem)
    for b in spec.blocks:
        bn(b.bn1)
        conv(b.conv1)
        bn(b.bn2)
        conv(b.conv2)
        if b.shortcut:
            conv(b.shortcut)
    if spec.final_bn is not None:
        bn(spec.final_bn)
    for k in spec.lstm_keys:
        parts.append(w[k])
    kk, bk = dense_keys(spec)
    parts.append(w[kk])
    parts.append(w[bk])
    return np.concatenate([np.ascontiguousarray(p, dtype=np.float32).reshape(-1) for p in parts])


class Model:
    """Object returned by ``load_model``; ``predict(x)`` mirrors Keras: numpy in, float32
    ``[B, n_classes]`` numpy out.  ``predict_device`` keeps everything on the GPU."""

    def __init__(self, spec: NetSpec, weights: Dict[str, np.ndarray], precision: str = "fp32"):
        """precision: 'fp32' (CUDA-core implicit GEMM, bit-faithful layer semantics), 'tf32' (tcgen05 tensor cores,
        TF32 operands, fp32 accumulation) or 'fp16' (as 'tf32', with fp16 operands — the same 11 significant bits, half
        the bytes — in the overlap net's conv pairs and in the LSTM recurrence of both nets)."""
        self.spec = spec
        self.weights = weights
        self._handle = None
        self._ws = None
        torch = _lib.require_cuda()
        lib = _lib.load()
        blob = pack_weights(spec, weights)
        handle = C.c_void_p()
        kind = KIND_OVERLAP if spec.ndim == 2 else KIND_SPEAKER
        _lib.check(lib.mmla_net_create(kind, spec.n_classes, HEAD_IDS[spec.head_activation],
                                       blob.ctypes.data, blob.size, C.byref(handle)), "mmla_net_create")
        self._handle = handle
        self._lib = lib
        self._torch = torch
        self.set_precision(precision)

    def set_precision(self, precision: str) -> None:
        if precision not in PRECISION_IDS:
            raise ValueError("precision must be 'fp32', 'tf32' or 'fp16'")
        _lib.check(self._lib.mmla_net_set_precision(self._handle, PRECISION_IDS[precision]),
                   "mmla_net_set_precision")
        self.precision = precision

    def __del__(self):
        try:
            if self._handle is not None:
                self._lib.mmla_net_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    @property
    def input_shape(self):
        return (128, 151, 3) if self.spec.ndim == 2 else (256, 39)

    def predict_device(self, x, return_labels: bool = True):
        """x: CUDA tensor [B,128,151,3] uint8|float32 (overlap) or [B,256,39] float32 (speaker).
        Returns (prob float32 CUDA [B,n], labels int32 CUDA [B])."""
        torch, lib = self._torch, self._lib
        pad40 = (self.spec.ndim == 1 and self.precision in ("tf32", "fp16") and tuple(x.shape[1:]) == (256, 40)
                 and x.dtype == torch.float32)      # channel-padded features (speaker_features_batch(row_stride=40))
        if tuple(x.shape[1:]) != self.input_shape and not pad40:
            raise ValueError(f"expected input [B, {self.input_shape}], got {tuple(x.shape)}")
        is_u8 = x.dtype == torch.uint8
        if not is_u8 and x.dtype != torch.float32:
            x = x.float()
        if is_u8 and self.spec.ndim != 2:
            x = x.float()
            is_u8 = False
        x = x.contiguous()
        B = x.shape[0]
        need = lib.mmla_net_workspace_bytes(self._handle, B)
        # one workspace per CUDA stream: forwards issued on different streams may run concurrently
        key = int(torch.cuda.current_stream().cuda_stream)
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need or ws.device != x.device:
            ws = self._ws[key] = torch.empty(need, dtype=torch.uint8, device=x.device)
        prob = torch.empty((B, self.spec.n_classes), dtype=torch.float32, device=x.device)
        labels = torch.empty((B,), dtype=torch.int32, device=x.device)
        _lib.check(lib.mmla_net_forward(self._handle, x.data_ptr(), 2 if pad40 else (1 if is_u8 else 0), B,
                                        ws.data_ptr(), ws.numel(), prob.data_ptr(),
                                        labels.data_ptr(), _lib.stream_ptr(torch)), "mmla_net_forward")
        return prob, labels

    def embed_device(self, x):
        """The frozen trunk's output — ``Model(base.input, base.layers[-2].output)`` of ``transfer_learning``
        (speaker_identification.py:402-406): float32 CUDA [B, 512] = [forward h | backward h] of the BiLSTM, the
        tensor the Dense head reads.  ``x`` as for :meth:`predict_device`."""
        torch, lib = self._torch, self._lib
        if tuple(x.shape[1:]) != self.input_shape:
            raise ValueError(f"expected input [B, {self.input_shape}], got {tuple(x.shape)}")
        is_u8 = x.dtype == torch.uint8 and self.spec.ndim == 2
        if not is_u8 and x.dtype != torch.float32:
            x = x.float()
        x = x.contiguous()
        B = x.shape[0]
        need = lib.mmla_net_workspace_bytes(self._handle, B)
        key = int(torch.cuda.current_stream().cuda_stream)
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need or ws.device != x.device:
            ws = self._ws[key] = torch.empty(need, dtype=torch.uint8, device=x.device)
        out = torch.empty((B, 512), dtype=torch.float32, device=x.device)
        _lib.check(lib.mmla_net_embed(self._handle, x.data_ptr(), 1 if is_u8 else 0, B, ws.data_ptr(), ws.numel(),
                                      out.data_ptr(), _lib.stream_ptr(torch)), "mmla_net_embed")
        return out

    def predict_device_cepstra(self, cep):
        """Speaker net, TF32 mode: ``cep`` float32 CUDA [B, T, 16] — the MFCC-13 rows of
        ``mfcc_batch(pcm, row_stride=16)`` (T = psf frame count <= 256).  Delta, delta-delta and the
        zero rows up to 256 are built inside the stem kernel; returns what ``predict_device`` returns
        on ``speaker_features_batch(pcm)``."""
        torch, lib = self._torch, self._lib
        if cep.dim() != 3 or cep.shape[2] != 16 or cep.dtype != torch.float32 or not cep.is_contiguous():
            raise ValueError("cepstra must be a contiguous float32 tensor [B, T, 16]")
        B, T = cep.shape[0], cep.shape[1]
        need = lib.mmla_net_workspace_bytes(self._handle, B)
        key = int(torch.cuda.current_stream().cuda_stream)
        if self._ws is None:
            self._ws = {}
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need or ws.device != cep.device:
            ws = self._ws[key] = torch.empty(need, dtype=torch.uint8, device=cep.device)
        prob = torch.empty((B, self.spec.n_classes), dtype=torch.float32, device=cep.device)
        labels = torch.empty((B,), dtype=torch.int32, device=cep.device)
        _lib.check(lib.mmla_net_forward_cepstra(self._handle, cep.data_ptr(), T * 16, T, B, ws.data_ptr(), ws.numel(),
                                                prob.data_ptr(), labels.data_ptr(), _lib.stream_ptr(torch)),
                   "mmla_net_forward_cepstra")
        return prob, labels

    def predict(self, x, batch_size=None, verbose=0):
        """Keras-style ``model.predict``: accepts numpy (any float dtype / uint8), returns numpy
        float32 probabilities.  Callers then do ``np.argmax(prob, axis=1)`` as in the reference."""
        torch = self._torch
        if isinstance(x, torch.Tensor):
            t = x if x.is_cuda else x.cuda()
        else:
            a = np.asarray(x)
            if a.dtype != np.uint8:
                a = a.astype(np.float32, copy=False)
            t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        prob, _ = self.predict_device(t)
        return prob.cpu().numpy()


def _read_model_weights(model_dir: str, spec: NetSpec) -> Dict[str, np.ndarray]:
    prefix = os.path.join(model_dir, "variables", "variables")
    want = set(weight_shapes(spec))
    got = tf_bundle.read_bundle(prefix, keys=want)
    return got


def load_model(model_dir: str, kind: Optional[str] = None, n_classes: Optional[int] = None,
               head: Optional[str] = None, precision: str = "fp32") -> Model:
    """``tf.keras.models.load_model(dir)`` stand-in.  The network family is inferred from the
    bundle's tensor names/shapes unless ``kind`` ('overlap' | 'speaker') is given."""
    prefix = os.path.join(model_dir, "variables", "variables")
    _, entries = tf_bundle.read_index(prefix + ".index")
    shapes = {e.key: e.shape for e in entries}
    if kind is None:
        k0 = shapes.get(lw(0, "kernel"))
        kind = "overlap" if k0 is not None and len(k0) == 4 else "speaker"
    if kind == "overlap":
        spec = OVERLAP
    else:
        custom = "customized_dense/kernel/.ATTRIBUTES/VARIABLE_VALUE"
        if custom in shapes:
            spec = speaker_spec(shapes[custom][1], head or "sigmoid")
        else:
            n = shapes[lw(42, "kernel")][1] if n_classes is None else n_classes
            spec = speaker_spec(n, head or "softmax")
    spec = resolve_lstm_keys(spec, shapes)
    return Model(spec, _read_model_weights(model_dir, spec), precision=precision)


def save_synthetic_model(model_dir: str, spec: NetSpec, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Write seeded synthetic weights as a TF tensor bundle under ``model_dir/variables/``."""
    w = synthetic_weights(spec, seed)
    tf_bundle.write_bundle(os.path.join(model_dir, "variables", "variables"), w, with_crc=True)
    return w
