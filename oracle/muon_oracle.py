"""CPU restatement of the reference's Muon / AuxAdam update rules (stockeh/swift ``training/optimizers/muon.py``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/`` (and nothing on the product path) as the checker of
``swb200_muon_step`` / ``swb200_adam_step``.  Pinned to the real reference by ``tests/golden/make_muon_golden.py`` ->
``tests/golden/muon.npz`` (the reference's own functions run on seeded matrices in this container).
"""
from __future__ import annotations

import torch


def newton_schulz5(G: torch.Tensor, steps: int = 5) -> torch.Tensor:
    """muon.py:5-33: quintic Newton-Schulz orthogonalisation in bf16 (transposed if tall, Frobenius-normalised first)."""
    a, b, c = 3.4445, -4.7750, 2.0315
    X = G.bfloat16()
    tall = G.size(-2) > G.size(-1)
    if tall:
        X = X.mT
    X = X / (X.norm(dim=(-2, -1), keepdim=True) + 1e-7)
    for _ in range(steps):
        A = X @ X.mT
        B = b * A + c * A @ A
        X = a * X + B @ X
    return X.mT if tall else X


def muon_update(grad: torch.Tensor, momentum: torch.Tensor, beta: float = 0.95, ns_steps: int = 5, nesterov: bool = True):
    """muon.py:36-45.  Returns (update, new momentum); the inputs are not modified (the reference lerps in place)."""
    momentum = torch.lerp(momentum, grad, 1 - beta)
    update = torch.lerp(grad, momentum, beta) if nesterov else momentum
    shape = update.shape
    if update.ndim == 4:
        update = update.view(len(update), -1)
    update = newton_schulz5(update, ns_steps).to(torch.float32)
    update = update * max(1, grad.size(-2) / grad.size(-1)) ** 0.5
    return update.reshape(shape), momentum


def muon_step(p, grad, momentum, lr: float, weight_decay: float, beta: float = 0.95):
    """muon.py:228-233: p <- p (1 - lr wd) - lr update.  Returns (new p, new momentum)."""
    update, momentum = muon_update(grad, momentum, beta)
    return p * (1 - lr * weight_decay) - lr * update.reshape(p.shape), momentum


def adam_step(p, grad, buf1, buf2, step: int, lr: float, betas=(0.9, 0.95), eps: float = 1e-10, weight_decay: float = 0.0):
    """muon.py:147-152 + :261-266.  Returns (new p, new buf1, new buf2)."""
    buf1 = torch.lerp(buf1, grad, 1 - betas[0])
    buf2 = torch.lerp(buf2, grad.square(), 1 - betas[1])
    update = (buf1 / (1 - betas[0] ** step)) / ((buf2 / (1 - betas[1] ** step)).sqrt() + eps)
    return p * (1 - lr * weight_decay) - lr * update, buf1, buf2
