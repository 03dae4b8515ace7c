"""``MuonWithAuxAdam`` on the device: the optimiser of the reference's sCM training configuration
(stockeh/swift ``training/optimizers/muon.py:155-266``, selected by ``configs/optimizer/muon.yaml``), same constructor
contract (param groups carrying ``use_muon``), same update rules, same distribution of the Muon matrices over ranks
(sorted by size, round-robin, ``all_gather`` of the updated parameters: muon.py:218-241).

Each Muon matrix is one ``swb200_muon_step`` call (momentum, Nesterov blend, five bf16 Newton-Schulz steps on the tcgen05
GEMM, parameter update), each AuxAdam tensor one ``swb200_adam_step``.  PyTorch holds the optimiser state and does the
collective.  Parameters must be contiguous fp32 CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from .engine import _aligned_buffer


class MuonWithAuxAdam(torch.optim.Optimizer):
    def __init__(self, param_groups, group=None, **kwargs):
        param_groups = [dict(g) for g in param_groups]
        for g in param_groups:
            assert "use_muon" in g
            g["params"] = list(g["params"])
            if g["use_muon"]:
                g["params"] = sorted(g["params"], key=lambda x: x.size(), reverse=True)        # muon.py:183-185
                g["lr"] = g.get("lr", 0.02)
                g["momentum"] = g.get("momentum", 0.95)
                g["weight_decay"] = g.get("weight_decay", 0)
                assert set(g.keys()) == {"params", "lr", "momentum", "weight_decay", "use_muon"}
            else:
                g["lr"] = g.get("lr", 3e-4)
                g["betas"] = g.get("betas", (0.9, 0.95))
                g["eps"] = g.get("eps", 1e-10)
                g["weight_decay"] = g.get("weight_decay", 0)
                assert set(g.keys()) == {"params", "lr", "betas", "eps", "weight_decay", "use_muon"}
        super().__init__(param_groups, dict())
        self._group = group
        self._ws = None
        self.lib = _lib.lib()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _check(p: torch.Tensor) -> None:
        if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
            raise RuntimeError("swift_b200.optim: parameters and gradients must be contiguous float32 CUDA tensors "
                               "(there is no CPU path)")

    def _workspace(self, rows: int, cols: int, batch: int, device):
        need = self.lib.swb200_muon_workspace_bytes(rows, cols, batch)
        if self._ws is None or self._ws[2] < need or self._ws[0].device != device:
            buf, base = _aligned_buffer(need, device)
            self._ws = (buf, base, need)
        return self._ws

    MAX_STACK_BYTES = 4 << 30       # workspace cap of one batched call (a stack of 12 Swift-B w1 matrices needs ~1.3 GB)

    def _muon_many(self, items, lr: float, wd: float, beta: float) -> None:
        """items: [(param, grad, momentum)] owned by this rank.  Matrices of one shape go through ONE batched
        ``swb200_muon_step`` (the Newton-Schulz products of the whole stack are single batched GEMM launches)."""
        import ctypes as C
        groups = {}
        for p, g, m in items:
            self._check(p), self._check(g)
            rows = p.shape[0]
            cols = p.numel() // rows                               # 4-D filters are viewed as [len, -1] (muon.py:39-40)
            if min(rows, cols) < 8 or rows % 8 or cols % 8:
                self._muon_small(p, g, m, rows, cols, lr, wd, beta)
            else:
                groups.setdefault((rows, cols, p.device), []).append((p, g, m))
        stream = torch.cuda.current_stream().cuda_stream
        for (rows, cols, dev), lst in groups.items():
            per = max(1, self.lib.swb200_muon_workspace_bytes(rows, cols, 1))
            step = max(1, min(len(lst), self.MAX_STACK_BYTES // per))
            for i in range(0, len(lst), step):
                part = lst[i:i + step]
                n = len(part)
                ws = self._workspace(rows, cols, n, dev)
                arr = lambda k: (C.c_void_p * n)(*[t[k].data_ptr() for t in part])
                _lib.check(self.lib.swb200_muon_step(arr(0), arr(1), arr(2), n, rows, cols, float(lr), float(wd), float(beta), 1, 5,
                                                     ws[1], ws[2], stream), "muon_step")

    def _muon_small(self, p, grad, mom, rows, cols, lr, wd, beta) -> None:
        """Shapes the GEMM kernel does not tile.  Vectors (Swift-B: the [1, heads, 1, 1] logit scales, which train.py:289
        hands to Muon as well) have their own kernel: with one row the Newton-Schulz products are scalars."""
        if min(rows, cols) != 1 or rows * cols > 4096:
            raise NotImplementedError(f"swift_b200 Muon handles matrices whose extents are multiples of 8 and vectors of at "
                                      f"most 4096 elements, got {rows} x {cols}")
        _lib.check(self.lib.swb200_muon_vector_step(p.data_ptr(), grad.data_ptr(), mom.data_ptr(), rows, cols, float(lr),
                                                    float(wd), float(beta), 1, 5, torch.cuda.current_stream().cuda_stream),
                   "muon_vector_step")

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        import torch.distributed as dist
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(self._group) > 1
        world = dist.get_world_size(self._group) if distributed else 1
        rank = dist.get_rank(self._group) if distributed else 0
        for group in self.param_groups:
            if group["use_muon"]:
                params = group["params"]
                pad = params + [torch.empty_like(params[-1])] * (world - len(params) % world) if distributed else params
                mine = []                      # the matrices of this rank: params[base + rank] for every base (muon.py:218-221)
                for base_i in range(0, len(params), world):
                    if base_i + rank < len(params):
                        p = params[base_i + rank]
                        if p.grad is None:
                            p.grad = torch.zeros_like(p)                                   # muon.py:223-225
                        st = self.state[p]
                        if len(st) == 0:
                            st["momentum_buffer"] = torch.zeros_like(p)
                        mine.append((p, p.grad, st["momentum_buffer"]))
                self._muon_many(mine, group["lr"], group["weight_decay"], group["momentum"])
                if distributed:                # same collectives as the reference, after this rank's updates instead of between them
                    for base_i in range(0, len(params), world):
                        dist.all_gather(pad[base_i:base_i + world], pad[base_i + rank], group=self._group)
            else:
                stream = torch.cuda.current_stream().cuda_stream
                for p in group["params"]:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                    self._check(p), self._check(p.grad)
                    st = self.state[p]
                    if len(st) == 0:
                        st["exp_avg"], st["exp_avg_sq"], st["step"] = torch.zeros_like(p), torch.zeros_like(p), 0
                    st["step"] += 1
                    b1, b2 = group["betas"]
                    _lib.check(self.lib.swb200_adam_step(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                                         st["exp_avg_sq"].data_ptr(), p.numel(), float(group["lr"]), float(b1),
                                                         float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                         int(st["step"]), stream), "adam_step")
        return loss


def swinv2_param_groups(net: torch.nn.Module, lr: float = 0.02, weight_decay: float = 0.01, adam_lr: float = 3e-4,
                        adam_betas=(0.9, 0.95), adam_weight_decay: float = 0.01, adam_eps: float = 1e-10):
    """The grouping train.py:286-313 applies for ``MuonWithAuxAdam`` + SwinV2: >= 2-D tensors under ``transformer`` go to
    Muon, everything else to AuxAdam (defaults: configs/optimizer/muon.yaml)."""
    muon, adam = [], []
    for name, p in net.named_parameters():
        (muon if p.ndim >= 2 and "transformer" in name else adam).append(p)
    return [dict(params=muon, use_muon=True, lr=lr, weight_decay=weight_decay),
            dict(params=adam, use_muon=False, lr=adam_lr, betas=tuple(adam_betas), eps=adam_eps, weight_decay=adam_weight_decay)]
