"""Oracle: torch-CPU fp32 restatement of the reference's two Keras classifiers
(inference mode).  TEST INFRASTRUCTURE ONLY (see oracle/__init__).

PARITY UNPINNED — TensorFlow/Keras 2.6 is not installable here and the trained weight shards
are stripped from the mount; this follows the model-building code and the Keras layer
semantics it relies on:

  * overlap net: ``res_block`` / ``ResLSTM``  OverlapDetection/scripts/overlap_detector_temp.py:253-303
    (layer graph confirmed by ``timit2.0/keras_metadata.pb``: BN eps 1e-3, LeakyReLU 0.3)
  * speaker net: ``res_unit`` / ``res_model``  SpeakerIdentification/scripts/speaker_identification.py:168-218
    and the transfer-learned head ``Dense(n, sigmoid)``  :403-410

Keras semantics restated: 'same' padding puts the extra element at the END (k=4: 1 before,
2 after; MaxPool on width 151: 1 on the right, -inf); BN = gamma*(x-mean)/sqrt(var+1e-3)+beta;
LSTM gates i,f,c,o with sigmoid recurrent activation; Bidirectional(return_sequences=False)
concatenates the forward layer's last state with the backward layer's last state (the one
that has seen t=0); Dropout is identity.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from mmla_audio_b200.weights import NetSpec, ConvSpec, lw, dense_keys

BN_EPS = 1e-3


def _same_pads(n: int, k: int, s: int):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _t(w: Dict[str, np.ndarray], key: str) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(w[key], dtype=np.float32))


def _bn(x: torch.Tensor, w, idx: int) -> torch.Tensor:
    g, b = _t(w, lw(idx, "gamma")), _t(w, lw(idx, "beta"))
    m, v = _t(w, lw(idx, "moving_mean")), _t(w, lw(idx, "moving_variance"))
    shape = (1, -1) + (1,) * (x.dim() - 2)
    return g.view(shape) * (x - m.view(shape)) / torch.sqrt(v.view(shape) + BN_EPS) + b.view(shape)


def _conv(x: torch.Tensor, w, spec: NetSpec, c: ConvSpec) -> torch.Tensor:
    """Keras Conv{1,2}D(padding='same').  x is NCHW (overlap) or NCW (speaker)."""
    k = _t(w, lw(c.idx, "kernel"))
    b = _t(w, lw(c.idx, "bias"))
    if spec.ndim == 2:
        kt = k.permute(3, 2, 0, 1).contiguous()            # HWIO → OIHW
        pt, pb = _same_pads(x.shape[2], c.kh, c.stride)
        pl, pr = _same_pads(x.shape[3], c.kw, c.stride)
        x = F.pad(x, (pl, pr, pt, pb))
        return F.conv2d(x, kt, b, stride=c.stride)
    kt = k.permute(2, 1, 0).contiguous()                   # WIO → OIW
    pl, pr = _same_pads(x.shape[2], c.kw, c.stride)
    x = F.pad(x, (pl, pr))
    return F.conv1d(x, kt, b, stride=c.stride)


def _maxpool_same(x: torch.Tensor, ndim: int) -> torch.Tensor:
    if ndim == 2:
        pt, pb = _same_pads(x.shape[2], 2, 2)
        pl, pr = _same_pads(x.shape[3], 2, 2)
        x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
        return F.max_pool2d(x, 2, 2)
    pl, pr = _same_pads(x.shape[2], 2, 2)
    x = F.pad(x, (pl, pr), value=float("-inf"))
    return F.max_pool1d(x, 2, 2)


def lstm_last(x: torch.Tensor, kernel, recurrent, bias, reverse: bool) -> torch.Tensor:
    """Keras LSTM(256, return_sequences=False) over x [B,T,F]; returns h after the last
    processed step (t=T-1 forward, t=0 when ``reverse``)."""
    B, T, _ = x.shape
    u = recurrent.shape[0]
    h = torch.zeros(B, u)
    c = torch.zeros(B, u)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        z = x[:, t] @ kernel + h @ recurrent + bias
        i = torch.sigmoid(z[:, :u])
        f = torch.sigmoid(z[:, u:2 * u])
        g = torch.tanh(z[:, 2 * u:3 * u])
        o = torch.sigmoid(z[:, 3 * u:])
        c = f * c + i * g
        h = o * torch.tanh(c)
    return h


def _bilstm(x: torch.Tensor, w, spec: NetSpec) -> torch.Tensor:
    ks = [_t(w, k) for k in spec.lstm_keys]
    fwd = lstm_last(x, ks[0], ks[1], ks[2], reverse=False)
    bwd = lstm_last(x, ks[3], ks[4], ks[5], reverse=True)
    return torch.cat([fwd, bwd], dim=1)


def _head(z: torch.Tensor, w, spec: NetSpec) -> torch.Tensor:
    kk, bk = dense_keys(spec)
    logits = z @ _t(w, kk) + _t(w, bk)
    if spec.head_activation == "softmax":
        return torch.softmax(logits, dim=1)
    return torch.sigmoid(logits)


@torch.no_grad()
def overlap_forward(x_nhwc: np.ndarray, w: Dict[str, np.ndarray], spec: NetSpec,
                    return_logits: bool = False) -> np.ndarray:
    """x float32 [B,128,151,3] (pixel values 0..255) → prob float32 [B,2]."""
    x = torch.from_numpy(np.ascontiguousarray(x_nhwc, dtype=np.float32)).permute(0, 3, 1, 2)
    net = _conv(x, w, spec, spec.stem)
    for b in spec.blocks:
        res = net
        if b.pool:
            res = _conv(net, w, spec, b.shortcut)
        out = F.elu(_bn(net, w, b.bn1))
        out = _conv(out, w, spec, b.conv1)
        out = F.elu(_bn(out, w, b.bn2))
        out = _conv(out, w, spec, b.conv2)
        if b.pool:
            out = _maxpool_same(out, 2)
        net = res + out
    seq = net.mean(dim=2).permute(0, 2, 1).contiguous()      # mean over H → [B, W, C]
    z = _bilstm(seq, w, spec)
    z = F.leaky_relu(z, 0.3)
    if return_logits:
        kk, bk = dense_keys(spec)
        return (z @ _t(w, kk) + _t(w, bk)).numpy()
    return _head(z, w, spec).numpy()


@torch.no_grad()
def speaker_forward(x_btf: np.ndarray, w: Dict[str, np.ndarray], spec: NetSpec,
                    return_logits: bool = False, return_embedding: bool = False) -> np.ndarray:
    """x float [B,256,39] → prob float32 [B,n_classes] (softmax base / sigmoid transfer head).
    ``return_embedding``: the 512-d output of ``layers[-2]`` instead (speaker_identification.py:403)."""
    x = torch.from_numpy(np.ascontiguousarray(x_btf, dtype=np.float32)).permute(0, 2, 1)
    net = _conv(x, w, spec, spec.stem)
    for b in spec.blocks:
        res = net
        y = net
        if b.pool:
            y = _maxpool_same(net, 1)
            res = _conv(net, w, spec, b.shortcut)
        out = F.relu(_bn(y, w, b.bn1))
        out = _conv(out, w, spec, b.conv1)
        out = F.relu(_bn(out, w, b.bn2))
        out = _conv(out, w, spec, b.conv2)
        net = res + out
    net = F.relu(_bn(net, w, spec.final_bn))
    net = F.avg_pool1d(net, 4)
    seq = net.permute(0, 2, 1).contiguous()                  # [B, 8, 128]
    z = _bilstm(seq, w, spec)
    if return_embedding:
        return z.numpy()
    if return_logits:
        kk, bk = dense_keys(spec)
        return (z @ _t(w, kk) + _t(w, bk)).numpy()
    return _head(z, w, spec).numpy()


def predict_labels(prob: np.ndarray) -> np.ndarray:
    """``np.argmax(prob, axis=1)`` (first maximum wins) — record_on_pc.py:160."""
    return np.argmax(prob, axis=1).astype(np.int32)
