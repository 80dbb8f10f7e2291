"""Weight layout of the reference's two classifiers, keyed exactly as in their SavedModel
``variables.index`` files, plus the seeded synthetic-weight generator used while the real
``variables.data-*`` shards are unavailable (``/root/reference/.MISSING_LARGE_BLOBS``).

Topology sources:
  * overlap net  — ``OverlapDetection/scripts/overlap_detector_temp.py:253-303`` and
    ``OverlapDetection/timit/models/timit2.0/{keras_metadata.pb,variables/variables.index}``
  * speaker net  — ``SpeakerIdentification/scripts/speaker_identification.py:168-218,401-410`` and
    ``SpeakerIdentification/timit/model/variables/variables.index``

Kernel layouts are TensorFlow's: Conv2D ``[kh,kw,cin,cout]`` (HWIO), Conv1D ``[k,cin,cout]``,
LSTM ``kernel [in,4u]`` / ``recurrent [u,4u]`` / ``bias [4u]`` with gate order i,f,c,o,
Dense ``[in,out]``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def lw(i: int, name: str) -> str:
    return f"layer_with_weights-{i}/{name}{_SUFFIX}"


@dataclass(frozen=True)
class ConvSpec:
    idx: int                      # layer_with_weights index
    kh: int
    kw: int
    cin: int
    cout: int
    stride: int = 1


@dataclass(frozen=True)
class BlockSpec:
    """One residual block/unit.  bn1→act→conv1→bn2→act→conv2 (+ shortcut conv when pooled)."""
    bn1: int
    conv1: ConvSpec
    bn2: int
    conv2: ConvSpec
    shortcut: Optional[ConvSpec] = None

    @property
    def pool(self) -> bool:
        return self.shortcut is not None


@dataclass(frozen=True)
class NetSpec:
    name: str
    ndim: int                              # 2 = Conv2D (overlap), 1 = Conv1D (speaker)
    stem: ConvSpec
    blocks: Tuple[BlockSpec, ...]
    final_bn: Optional[int]                # speaker only
    lstm_keys: Tuple[str, ...]             # fwd kernel, fwd recurrent, fwd bias, bwd ...
    dense_idx: int
    n_classes: int
    head_activation: str                   # 'softmax' | 'sigmoid'
    dense_key_prefix: Optional[str] = None  # override for the transfer-learned head


def _overlap_spec() -> NetSpec:
    def blk(base, cin, cout, pool):
        # indices: BN, conv3x3, BN, conv(4,1)[, shortcut 1x1 stride 2]
        return BlockSpec(
            bn1=base, conv1=ConvSpec(base + 1, 3, 3, cin, cout),
            bn2=base + 2, conv2=ConvSpec(base + 3, 4, 1, cout, cout),
            shortcut=ConvSpec(base + 4, 1, 1, cin, cout, stride=2) if pool else None)
    blocks = (blk(1, 16, 32, True), blk(6, 32, 32, False), blk(10, 32, 32, False),
              blk(14, 32, 64, True), blk(19, 64, 64, False), blk(23, 64, 64, False),
              blk(27, 64, 128, True), blk(32, 128, 128, False), blk(36, 128, 128, False))
    lstm = tuple(f"variables/{i}{_SUFFIX}" for i in range(116, 122))
    return NetSpec("overlap", 2, ConvSpec(0, 1, 1, 3, 16), blocks, None, lstm, 41, 2, "softmax")


def _speaker_spec(n_classes: int = 630, head: str = "softmax") -> NetSpec:
    def unit(base, cin, cout, pool):
        if pool:   # BN, conv3, BN, shortcut(k1,s2), conv3
            return BlockSpec(bn1=base, conv1=ConvSpec(base + 1, 1, 3, cin, cout),
                             bn2=base + 2, conv2=ConvSpec(base + 4, 1, 3, cout, cout),
                             shortcut=ConvSpec(base + 3, 1, 1, cin, cout, stride=2))
        return BlockSpec(bn1=base, conv1=ConvSpec(base + 1, 1, 3, cin, cout),
                         bn2=base + 2, conv2=ConvSpec(base + 3, 1, 3, cout, cout))
    blocks = (unit(1, 32, 32, True), unit(6, 32, 32, False), unit(10, 32, 32, False),
              unit(14, 32, 64, True), unit(19, 64, 64, False), unit(23, 64, 64, False),
              unit(27, 64, 128, True), unit(32, 128, 128, False), unit(36, 128, 128, False))
    lstm = tuple(f"trainable_variables/{i}{_SUFFIX}" for i in range(82, 88))
    return NetSpec("speaker", 1, ConvSpec(0, 1, 4, 39, 32), blocks, 40, lstm, 42, n_classes,
                   head, None if head == "softmax" and n_classes == 630 else "customized_dense")


OVERLAP = _overlap_spec()
SPEAKER_BASE = _speaker_spec()


def speaker_spec(n_classes: int = 630, head: str = "softmax") -> NetSpec:
    """Base TIMIT model (630-way softmax) or the transfer-learned head
    (``Dense(n, sigmoid, name='customized_dense')``, speaker_identification.py:403-410)."""
    return _speaker_spec(n_classes, head)


def resolve_lstm_keys(spec: NetSpec, index_shapes: Dict[str, Tuple[int, ...]]) -> NetSpec:
    """The Bidirectional layer's six tensors are stored under checkpoint-dependent names
    (``variables/116..121`` in timit2.0, ``trainable_variables/80..85`` in timit1.0,
    ``trainable_variables/82..87`` in the speaker model).  Find them by shape: the first run of
    six consecutively numbered entries shaped [feat,1024],[256,1024],[1024] twice."""
    import re
    from dataclasses import replace
    if all(k in index_shapes for k in spec.lstm_keys):
        return spec
    feat = spec.blocks[-1].conv2.cout
    want = [(feat, 1024), (256, 1024), (1024,)] * 2
    groups: Dict[str, Dict[int, str]] = {}
    for k in index_shapes:
        m = re.match(r"^(trainable_variables|variables)/(\d+)/\.ATTRIBUTES/VARIABLE_VALUE$", k)
        if m:
            groups.setdefault(m.group(1), {})[int(m.group(2))] = k
    for prefix in ("variables", "trainable_variables"):
        nums = groups.get(prefix, {})
        for n0 in sorted(nums):
            keys = [nums.get(n0 + i) for i in range(6)]
            if all(keys) and [tuple(index_shapes[k]) for k in keys] == want:
                return replace(spec, lstm_keys=tuple(keys))
    raise KeyError(f"{spec.name}: no Bidirectional(LSTM(256)) tensors found in the checkpoint index")


def dense_keys(spec: NetSpec) -> Tuple[str, str]:
    if spec.dense_key_prefix:
        return (f"{spec.dense_key_prefix}/kernel{_SUFFIX}", f"{spec.dense_key_prefix}/bias{_SUFFIX}")
    return lw(spec.dense_idx, "kernel"), lw(spec.dense_idx, "bias")


def conv_kernel_shape(spec: NetSpec, c: ConvSpec) -> Tuple[int, ...]:
    return (c.kh, c.kw, c.cin, c.cout) if spec.ndim == 2 else (c.kw, c.cin, c.cout)


def weight_shapes(spec: NetSpec) -> Dict[str, Tuple[int, ...]]:
    """{TF key: shape} for every inference weight of ``spec`` (optimizer slots excluded)."""
    out: Dict[str, Tuple[int, ...]] = {}

    def conv(c: ConvSpec):
        out[lw(c.idx, "kernel")] = conv_kernel_shape(spec, c)
        out[lw(c.idx, "bias")] = (c.cout,)

    def bn(i: int, ch: int):
        for n in ("gamma", "beta", "moving_mean", "moving_variance"):
            out[lw(i, n)] = (ch,)

    conv(spec.stem)
    for b in spec.blocks:
        bn(b.bn1, b.conv1.cin)
        conv(b.conv1)
        bn(b.bn2, b.conv1.cout)
        conv(b.conv2)
        if b.shortcut:
            conv(b.shortcut)
    feat = spec.blocks[-1].conv2.cout
    if spec.final_bn is not None:
        bn(spec.final_bn, feat)
    for d in range(2):
        out[spec.lstm_keys[3 * d + 0]] = (feat, 1024)
        out[spec.lstm_keys[3 * d + 1]] = (256, 1024)
        out[spec.lstm_keys[3 * d + 2]] = (1024,)
    kk, bk = dense_keys(spec)
    out[kk] = (512, spec.n_classes)
    out[bk] = (spec.n_classes,)
    return out


def synthetic_weights(spec: NetSpec, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Seeded synthetic weights with exactly the reference's shapes:
    He-normal conv kernels (stem scaled for 0..255 / int16-scale inputs), BN gamma~U[.5,1.5],
    beta~N(0,.1), mean~N(0,.1), var~U[.5,1.5], Glorot LSTM / dense kernels, N(0,.05) biases.
    Residual-branch output convs are damped (x0.5) so nine stacked blocks stay O(1)."""
    rng = np.random.default_rng(seed)
    shapes = weight_shapes(spec)
    damp = {lw(b.conv2.idx, "kernel") for b in spec.blocks}
    stem_key = lw(spec.stem.idx, "kernel")
    w: Dict[str, np.ndarray] = {}
    for key in sorted(shapes):
        shp = shapes[key]
        if key.endswith("kernel" + _SUFFIX) and len(shp) >= 3:
            fan_in = int(np.prod(shp[:-1]))
            a = rng.normal(0.0, np.sqrt(2.0 / fan_in), shp)
            if key in damp:
                a *= 0.5
            if key == stem_key:
                # overlap stem sees pixels 0..255, speaker stem sees MFCCs of O(10..100)
                a *= (1.0 / 128.0) if spec.ndim == 2 else (1.0 / 16.0)
        elif len(shp) == 2:                                   # LSTM / dense kernels: Glorot
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            a = rng.uniform(-lim, lim, shp)
        elif key.endswith("gamma" + _SUFFIX) or key.endswith("moving_variance" + _SUFFIX):
            a = rng.uniform(0.5, 1.5, shp)
        elif key.endswith("beta" + _SUFFIX) or key.endswith("moving_mean" + _SUFFIX):
            a = rng.normal(0.0, 0.1, shp)
        else:                                                 # biases
            a = rng.normal(0.0, 0.05, shp)
        w[key] = a.astype(np.float32)
    return w


def check_weights(spec: NetSpec, w: Dict[str, np.ndarray]) -> None:
    """Raise KeyError/ValueError unless ``w`` holds every tensor of ``spec`` with its shape."""
    for key, shp in weight_shapes(spec).items():
        if key not in w:
            raise KeyError(f"{spec.name}: missing weight {key}")
        if tuple(w[key].shape) != tuple(shp):
            raise ValueError(f"{spec.name}: {key} has shape {w[key].shape}, expected {shp}")
