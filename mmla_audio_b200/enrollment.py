"""Speaker enrollment: frozen-trunk embeddings + fit of the transfer head — the first phase of
``transfer_learning`` (SpeakerIdentification/scripts/speaker_identification.py:401-432) on the device.

    sliced_base_model = Model(base_model.input, base_model.layers[-2].output)   → ``Model.embed_device``
    Dense(dim, activation='sigmoid', name='customized_dense')                    → ``fit_head``
    compile(loss="categorical_crossentropy", optimizer=RMSprop(lr=0.0001)); fit(batch_size=16, epochs=500)

The second phase of the reference (:438-447: un-freeze everything, 20 epochs at lr 1e-6) trains the trunk itself and
is training code outside the hot path; ``transfer_learning`` here stops after the head fit and says so.
"""
from __future__ import annotations

import os
from dataclasses import replace
from typing import Optional, Tuple

import numpy as np

from . import _lib, tf_bundle
from .weights import dense_keys, speaker_spec

CUSTOM_K = "customized_dense/kernel/.ATTRIBUTES/VARIABLE_VALUE"
CUSTOM_B = "customized_dense/bias/.ATTRIBUTES/VARIABLE_VALUE"


def glorot_uniform(n_in: int, n_out: int, seed: int) -> np.ndarray:
    """Keras' default Dense initialiser: U(-l, l), l = sqrt(6 / (fan_in + fan_out))."""
    lim = np.sqrt(6.0 / (n_in + n_out))
    return np.random.default_rng(seed).uniform(-lim, lim, size=(n_in, n_out)).astype(np.float32)


def epoch_orders(n_samples: int, epochs: int, seed: int) -> np.ndarray:
    """int32 [epochs, n_samples]: one permutation per epoch (Keras ``shuffle=True``)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.permutation(n_samples) for _ in range(epochs)]).astype(np.int32)


def fit_head(embed, y_onehot, epochs: int = 500, batch_size: int = 16, lr: float = 1e-4, rho: float = 0.9,
             eps: float = 1e-7, seed: int = 1, kernel0=None, bias0=None, order=None):
    """embed: float32 [M, 512] (CUDA tensor or numpy); y_onehot: [M, n].  Returns (kernel [512, n], bias [n],
    loss per epoch [epochs]) as numpy float32."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    e = embed if isinstance(embed, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(embed, np.float32))
    e = e.to("cuda", torch.float32).contiguous()
    y = torch.from_numpy(np.ascontiguousarray(y_onehot, np.float32)).cuda()
    M, n = y.shape
    if e.shape != (M, 512):
        raise ValueError(f"embed must be [{M}, 512]")
    k0 = glorot_uniform(512, n, seed) if kernel0 is None else np.ascontiguousarray(kernel0, np.float32)
    b0 = np.zeros(n, np.float32) if bias0 is None else np.ascontiguousarray(bias0, np.float32)
    order = epoch_orders(M, epochs, seed) if order is None else np.ascontiguousarray(order, np.int32)
    if order.shape != (epochs, M):
        raise ValueError("order must be [epochs, n_samples]")
    k = torch.from_numpy(k0.copy()).cuda()
    b = torch.from_numpy(b0.copy()).cuda()
    od = torch.from_numpy(order).cuda()
    loss = torch.zeros((max(epochs, 1),), dtype=torch.float32, device="cuda")
    _lib.check(lib.mmla_head_fit(e.data_ptr(), y.data_ptr(), M, n, od.data_ptr(), epochs, batch_size, lr, rho, eps,
                                 k.data_ptr(), b.data_ptr(), loss.data_ptr(), _lib.stream_ptr(torch)), "mmla_head_fit")
    return k.cpu().numpy(), b.cpu().numpy(), loss[:epochs].cpu().numpy()


def with_transfer_head(base_model, kernel: np.ndarray, bias: np.ndarray, precision: Optional[str] = None):
    """The base model's trunk with a new ``Dense(n, sigmoid)`` head (``customized_dense``)."""
    from .models import Model
    n = kernel.shape[1]
    spec = speaker_spec(n, "sigmoid")
    w = dict(base_model.weights)
    old_k, old_b = dense_keys(base_model.spec)
    w.pop(old_k, None)
    w.pop(old_b, None)
    spec = replace(spec, lstm_keys=base_model.spec.lstm_keys)      # checkpoint-dependent tensor names of the BiLSTM
    kk, bk = dense_keys(spec)
    w[kk], w[bk] = np.ascontiguousarray(kernel, np.float32), np.ascontiguousarray(bias, np.float32)
    return Model(spec, w, precision=precision or base_model.precision)


def transfer_learning(_x, _y, seed, _test_split_ratio, base_model, final_model_path: Optional[str] = None,
                      epochs: int = 500, batch_size: int = 16) -> Tuple[float, object]:
    """``transfer_learning(_x, _y, seed, _test_split_ratio, base_model_path, final_model_path)`` (:401-477), head-fit
    phase: stratified 70/30 train/validation split (:421-422, sklearn, ``random_state=seed``), embeddings of the frozen
    trunk, head fit, validation accuracy (``categorical_accuracy``).  ``base_model``: a loaded ``models.Model`` or a
    SavedModel directory.  Returns (accuracy, fitted model); the model is also saved to ``final_model_path``."""
    from sklearn.model_selection import train_test_split
    from .models import load_model
    torch = _lib.require_cuda()
    if isinstance(base_model, (str, os.PathLike)):
        base_model = load_model(os.fspath(base_model), kind="speaker")
    x, y = np.asarray(_x, np.float32), np.asarray(_y, np.float32)
    x_train, y_train = x, y
    x_test = y_test = None
    if _test_split_ratio > 0:
        x_train, x_test, y_train, y_test = train_test_split(x_train, y_train, test_size=_test_split_ratio, stratify=y, random_state=seed)
    x_train, x_val, y_train, y_val = train_test_split(x_train, y_train, test_size=0.3, stratify=y_train, random_state=seed)
    emb = base_model.embed_device(torch.from_numpy(np.ascontiguousarray(x_train)).cuda())
    k, b, _loss = fit_head(emb, y_train, epochs=epochs, batch_size=batch_size, seed=seed)
    model = with_transfer_head(base_model, k, b)
    xe, ye = (x_test, y_test) if _test_split_ratio > 0 else (x_val, y_val)
    prob = model.predict(xe)
    accuracy = float((prob.argmax(1) == ye.argmax(1)).mean())
    if final_model_path:
        save_model(model, final_model_path)
    return accuracy, model


def save_model(model, model_dir: str) -> None:
    """Write the model as a TF tensor bundle ``model_dir/variables/variables.*`` that ``load_model`` reads back (the
    head under the reference's layer name ``customized_dense``)."""
    w = dict(model.weights)
    kk, bk = dense_keys(model.spec)
    if model.spec.head_activation == "sigmoid" and kk != CUSTOM_K:
        w[CUSTOM_K], w[CUSTOM_B] = w.pop(kk), w.pop(bk)
    tf_bundle.write_bundle(os.path.join(model_dir, "variables", "variables"), w)
