"""Oracle: the head-fit phase of the reference's ``transfer_learning``
(SpeakerIdentification/scripts/speaker_identification.py:401-432) on precomputed trunk embeddings —
TEST INFRASTRUCTURE ONLY (see oracle/__init__).  PARITY UNPINNED (Keras is not installable here).

Restates Keras 2.6 semantics with torch-CPU autograd:
  * ``Dense(dim, activation='sigmoid')``;
  * ``loss="categorical_crossentropy"`` on probabilities (from_logits=False): outputs divided by their sum over the class
    axis, clipped to [1e-7, 1 - 1e-7], ``-sum(y * log(q))``, mean over the mini-batch;
  * ``RMSprop(lr=1e-4)``: rho 0.9, momentum 0, epsilon 1e-7, ``var -= lr * g / (sqrt(rms) + eps)`` — torch.optim.RMSprop
    with alpha = rho computes exactly that update;
  * ``fit(batch_size=16, epochs=500)`` visiting the samples in a given per-epoch order, last mini-batch short.
The gradient comes from autograd, so it is independent of the hand-derived expression in csrc/head_fit.cu.
"""
from __future__ import annotations

import numpy as np
import torch


def keras_categorical_crossentropy(y, p):
    q = p / p.sum(dim=-1, keepdim=True)
    q = torch.clamp(q, 1e-7, 1.0 - 1e-7)
    return -(y * torch.log(q)).sum(dim=-1)


def fit_head(embed, y_onehot, kernel0, bias0, order, batch_size=16, lr=1e-4, rho=0.9, eps=1e-7):
    """→ (kernel [512,n], bias [n], mean training loss per epoch [epochs]); float32 throughout."""
    torch.manual_seed(0)
    e = torch.from_numpy(np.ascontiguousarray(embed, np.float32))
    y = torch.from_numpy(np.ascontiguousarray(y_onehot, np.float32))
    k = torch.tensor(np.asarray(kernel0, np.float32), requires_grad=True)
    b = torch.tensor(np.asarray(bias0, np.float32), requires_grad=True)
    opt = torch.optim.RMSprop([k, b], lr=lr, alpha=rho, eps=eps, momentum=0.0, centered=False)
    losses = []
    M = e.shape[0]
    for ep in range(order.shape[0]):
        tot = 0.0
        for s0 in range(0, M, batch_size):
            idx = torch.from_numpy(np.asarray(order[ep, s0:s0 + batch_size], np.int64))
            p = torch.sigmoid(e[idx] @ k + b)
            per = keras_categorical_crossentropy(y[idx], p)
            loss = per.mean()
            opt.zero_grad()
            loss.backward()
            opt.step()
            tot += float(per.sum())
        losses.append(tot / M)
    return k.detach().numpy(), b.detach().numpy(), np.asarray(losses, np.float32)
