"""CPU/fp32 oracle for the Swift forecast hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional, plain-PyTorch fp32 restatement of the reference
algorithm (stockeh/swift, ``src/swift``).  It is the checker that the CUDA path
is compared against; it is never the thing that is shipped or measured.  Only
``tests/``, ``__graft_entry__.smoke()`` and the baseline legs of ``bench.py``
(``cpu_baseline`` / ``--impl reference`` on the host cores, and the
``gpu_eager_baseline`` SURVEY.md section 8d asks for: the same plain-PyTorch ops
run eagerly on the GPU, reported beside the product's number) may import it.  The product package
``swift_b200`` must never import anything from ``oracle/``.

Pinning: ``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` is mounted) executes the *real* reference modules on the
seeded fixtures of ``swift_b200.synthetic`` and stores their outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this restatement
against those files.  The reference itself ships no tests or golden vectors
(SURVEY.md section 4), so those generated vectors are the pin.

Every function cites the reference file:line it follows (paths relative to
``/root/reference/src/swift``).  The code is written as free functions over a
flat ``{name: tensor}`` state dict (the reference's own ``state_dict`` schema),
not as a copy of the reference's ``nn.Module`` tree.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------
# helpers


def _pair(v) -> Tuple[int, int]:
    """models/abstract.py:42-62 (Shape2D): int -> (v, v); list/tuple -> tuple."""
    if isinstance(v, int):
        return (v, v)
    v = tuple(int(a) for a in v)
    assert len(v) == 2
    return v


def _linear(x: torch.Tensor, p: Params, name: str) -> torch.Tensor:
    w = p[name + ".weight"]
    b = p.get(name + ".bias")
    return F.linear(x, w, b)


# ----------------------------------------------------------------------------
# conditioning  (models/swinv2.py:44-60, :67-74, :316-321)


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10_000.0) -> torch.Tensor:
    """Sinusoidal embedding, returned as [sin | cos].

    models/swinv2.py:44-60 builds cat([cos, sin]) and then swaps the two halves
    ("flip sin/cos as done with edm"), i.e. the result is [sin(args) | cos(args)].
    Odd ``dim`` appends one zero column *before* the swap (:52-53); the swap
    reshapes to (B, 2, -1) so for odd dim it would fail in the reference too.
    """
    half = dim // 2
    k = torch.arange(half, dtype=t.dtype, device=t.device)
    freqs = torch.exp(-math.log(max_period) * k / half)
    args = t[:, None] * freqs[None, :]
    return torch.cat([torch.sin(args), torch.cos(args)], dim=-1)


def conditioning_vector(p: Params, t: torch.Tensor, auxiliary: Optional[torch.Tensor], dim: int,
                        auxiliary_dim: int, timestep_weight: float, batch: int) -> torch.Tensor:
    """models/swinv2.py:316-321: t -> repeat to batch -> sinusoid -> (+aux embed) -> 2x(Linear+SiLU)."""
    if t.dim() == 0 or (t.dim() == 1 and t.shape[0] == 1):
        t = t.reshape(-1).repeat(batch)
    emb = timestep_embedding(t * timestep_weight, dim)
    if auxiliary_dim and auxiliary is not None and ("auxiliary_embed.weight" in p):
        emb = emb + _linear(auxiliary * math.sqrt(auxiliary_dim), p, "auxiliary_embed")
    h = F.silu(_linear(emb, p, "latent_embed.l1"))      # models/swinv2.py:73-74
    return F.silu(_linear(h, p, "latent_embed.l2"))


# ----------------------------------------------------------------------------
# blocks


def modulated_norm(x: torch.Tensor, c: torch.Tensor, p: Params, prefix: str, eps: float = 1e-6) -> torch.Tensor:
    """models/swinv2.py:77-86: LN(x; affine, eps=1e-6) * (1 + scale(c)) + shift(c).

    ``x`` is [B, n, D]; ``c`` is the per-sample conditioning vector [B, D].
    """
    d = x.shape[-1]
    y = F.layer_norm(x, (d,), p[prefix + ".norm.weight"], p[prefix + ".norm.bias"], eps)
    mod = _linear(c, p, prefix + ".modulation")
    scale, shift = mod[:, :d], mod[:, d:]
    return y * (1.0 + scale[:, None, :]) + shift[:, None, :]


def cosine_window_attention(xw: torch.Tensor, p: Params, prefix: str, heads: int, sdpa: bool = False) -> torch.Tensor:
    """Scaled-cosine attention inside each window, before ``wo``.

    models/swinv2.py:118-135.  ``xw`` is [nWin*B, n, D].  The fused qkv output is
    split per head first ("b n (h d) -> b h n d" with d = 3*head_dim) and only
    then chunked into q, k, v, so column h*3*hd + {0..hd-1: q, hd..2hd-1: k,
    2hd..3hd-1: v}.  q and k are L2-normalised (F.normalize eps 1e-12); q is
    multiplied by exp(min(scale_h, ln 100)); softmax logits use scale 1.0; no
    mask and no position bias.  ``sdpa`` selects the reference's inference
    branch (F.scaled_dot_product_attention); the explicit softmax is its
    ``jvp`` branch and the default here.
    """
    nb, n, d = xw.shape
    hd = d // heads
    qkv = F.linear(xw, p[prefix + ".to_qkv.weight"])                 # [nb, n, 3D]
    qkv = qkv.reshape(nb, n, heads, 3 * hd).permute(0, 2, 1, 3)      # [nb, h, n, 3hd]
    q, k, v = qkv[..., :hd], qkv[..., hd:2 * hd], qkv[..., 2 * hd:]
    logit_scale = torch.clamp(p[prefix + ".scale"], max=math.log(1.0 / 0.01)).exp()  # [1,h,1,1]
    q = q / q.norm(dim=-1, keepdim=True).clamp_min(1e-12) * logit_scale
    k = k / k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    if sdpa:                                                          # :128-129 (flash and not jvp)
        o = F.scaled_dot_product_attention(q, k, v, scale=1.0)
    else:                                                             # :130-133
        attn = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
        o = attn @ v                                                  # [nb, h, n, hd]
    return o.permute(0, 2, 1, 3).reshape(nb, n, d)


def to_windows(x: torch.Tensor, grid: Tuple[int, int], win: Tuple[int, int]) -> torch.Tensor:
    """models/swinv2.py:17-28 applied to tokens: [B, gh*gw, D] -> [B*nWin, wh*ww, D]."""
    b, _, d = x.shape
    gh, gw = grid
    wh, ww = win
    x = x.reshape(b, gh // wh, wh, gw // ww, ww, d).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, wh * ww, d)


def from_windows(xw: torch.Tensor, grid: Tuple[int, int], win: Tuple[int, int]) -> torch.Tensor:
    """models/swinv2.py:31-41: inverse of ``to_windows``."""
    gh, gw = grid
    wh, ww = win
    d = xw.shape[-1]
    x = xw.reshape(-1, gh // wh, gw // ww, wh, ww, d).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, gh * gw, d)


def swin_block(x: torch.Tensor, c: torch.Tensor, p: Params, i: int, cfg: dict) -> torch.Tensor:
    """One layer of models/swinv2.py:186-212 (res-post-norm, no attention mask)."""
    grid, win, shift = cfg["grid"], cfg["window"], cfg["shift"]
    b, n, d = x.shape
    pa, pf = f"transformer.layers.{i}.0", f"transformer.layers.{i}.1"
    shifted = any(shift) and (i % 2 != 0)

    y = x.reshape(b, grid[0], grid[1], d)
    if shifted:                                                       # :192-193
        y = torch.roll(y, shifts=(-shift[0], -shift[1]), dims=(1, 2))
    yw = to_windows(y.reshape(b, n, d), grid, win)                    # :196-197
    nwin = yw.shape[0] // b
    o = cosine_window_attention(yw, p, pa, cfg["heads"], cfg.get("sdpa", False))
    o = F.linear(o, p[pa + ".wo.weight"])                             # :137
    o = modulated_norm(o, c.repeat_interleave(nwin, dim=0), p, pa + ".norm")  # :138, :184
    o = from_windows(o, grid, win).reshape(b, grid[0], grid[1], d)    # :202-203
    if shifted:                                                       # :206-207
        o = torch.roll(o, shifts=(shift[0], shift[1]), dims=(1, 2))
    x = x + o.reshape(b, n, d)                                        # :211

    h = F.linear(x, p[pf + ".w1.weight"])                             # :99
    dff = h.shape[-1] // 2
    h = F.silu(h[..., :dff]) * h[..., dff:]                           # :99-100 (gate first)
    h = F.linear(h, p[pf + ".w2.weight"])
    return x + modulated_norm(h, c, p, pf + ".norm")                  # :101, :212


def patchify(x: torch.Tensor, patch: Tuple[int, int]) -> torch.Tensor:
    """models/swinv2.py:224-229: "b c (h p1) (w p2) -> b (h w) (p1 p2 c)"."""
    b, c, hh, ww = x.shape
    p1, p2 = patch
    x = x.reshape(b, c, hh // p1, p1, ww // p2, p2).permute(0, 2, 4, 3, 5, 1)
    return x.reshape(b, (hh // p1) * (ww // p2), p1 * p2 * c)


def unpatchify(y: torch.Tensor, grid: Tuple[int, int], patch: Tuple[int, int]) -> torch.Tensor:
    """models/swinv2.py:241-243: "b (h w) (c p1 p2) -> b c (h p1) (w p2)"  (note: c-major, unlike patchify)."""
    b = y.shape[0]
    gh, gw = grid
    p1, p2 = patch
    c = y.shape[-1] // (p1 * p2)
    y = y.reshape(b, gh, gw, c, p1, p2).permute(0, 3, 1, 4, 2, 5)
    return y.reshape(b, c, gh * p1, gw * p2)


# ----------------------------------------------------------------------------
# whole network  (models/swinv2.py:305-330)


def make_cfg(img_resolution, in_channels: int, out_channels: int, window_size, shift_size, patch_size,
             depth: int, dim: int, heads: int, auxiliary_dim: int = 0, logvar: bool = False,
             timestep_weight: float = 1.0) -> dict:
    res, win, shift, patch = _pair(img_resolution), _pair(window_size), _pair(shift_size), _pair(patch_size)
    grid = (res[0] // patch[0], res[1] // patch[1])                   # models/swinv2.py:274
    return dict(res=res, window=win, shift=shift, patch=patch, grid=grid, in_channels=in_channels,
                out_channels=out_channels, depth=depth, dim=dim, heads=heads,
                auxiliary_dim=auxiliary_dim, logvar=logvar, timestep_weight=timestep_weight)


def swinv2_forward(p: Params, cfg: dict, x: torch.Tensor, t: torch.Tensor,
                   auxiliary: Optional[torch.Tensor] = None, return_logvar: bool = False,
                   taps: Optional[dict] = None):
    """models/swinv2.py:305-330.  ``taps`` (optional dict) receives intermediate tensors."""
    tok = F.linear(patchify(x, cfg["patch"]), p["patch_embed.emb.weight"], p["patch_embed.emb.bias"])
    tok = tok + p["pos_embed"]                                        # :314
    c = conditioning_vector(p, t, auxiliary, cfg["dim"], cfg["auxiliary_dim"], cfg["timestep_weight"],
                            x.shape[0])
    if taps is not None:
        taps["embed"] = tok
        taps["cond"] = c
    for i in range(cfg["depth"]):
        tok = swin_block(tok, c, p, i, cfg)
        if taps is not None:
            taps[f"block{i}"] = tok
    y = unpatchify(F.linear(tok, p["head.head.0.weight"]), cfg["grid"], cfg["patch"])
    if cfg["logvar"] and return_logvar:                               # :326-328
        return y, _linear(c, p, "logvar_embed").squeeze(-1)
    return y


# ----------------------------------------------------------------------------
# preconditioner + samplers


def pass_precond(p: Params, cfg: dict, x: torch.Tensor, t: torch.Tensor, condition: Optional[torch.Tensor],
                 auxiliary) -> torch.Tensor:
    """models/precond.py:133-148 (+ :21-31): aux -> [B, aux_dim]; cat([x, condition], 1); model(arg, t.flatten()).

    ``p`` holds *model* keys (no ``model.`` prefix).
    """
    b = x.shape[0]
    aux = None
    adim = cfg["auxiliary_dim"]
    if adim:
        if auxiliary is None:
            aux = torch.zeros(1, adim, device=x.device, dtype=x.dtype)
        else:
            aux = auxiliary if isinstance(auxiliary, torch.Tensor) else torch.tensor(auxiliary, device=x.device)
            if aux.dim() == 0 or (aux.dim() == 1 and aux.shape[0] == 1):
                aux = aux.reshape(-1).repeat(b)
            aux = aux.reshape(-1, adim).to(x.dtype)
    arg = x if condition is None else torch.cat([x, condition], dim=1)
    return swinv2_forward(p, cfg, arg, t.flatten(), aux)


def scm_time_grid(num_steps: int, sigma_min: float, sigma_max: float, sigma_data: float,
                  intermediates: Optional[Sequence[float]] = None) -> torch.Tensor:
    """generating/diffusion.py:435-450: the t grid of the sCM sampler (last entry 0)."""
    if num_steps == 1:
        ts = torch.tensor([math.pi / 2], dtype=torch.float32)
    else:
        lo, hi = torch.log(torch.tensor(sigma_min)), torch.log(torch.tensor(sigma_max))
        u = torch.linspace(1, 0, num_steps)
        ts = torch.atan(torch.exp(lo + u * (hi - lo)) / sigma_data)
    ts = torch.cat([ts, torch.zeros(1)])
    if num_steps == 2 and intermediates is None:
        ts = torch.tensor([float(ts[0]), 1.1, 0.0])
    elif intermediates:
        ts = torch.cat([ts[:1], torch.as_tensor(list(intermediates), dtype=torch.float32), ts[-1:]])
    return ts


def scm_solver(net: Callable, latents: torch.Tensor, condition: torch.Tensor, auxiliary,
               num_steps: int = 1, sigma_min: float = 0.02, sigma_max: float = 200.0, sigma_data: float = 1.0,
               intermediates=None, noise_fn: Optional[Callable] = None) -> torch.Tensor:
    """generating/diffusion.py:417-461.  ``net(x, t[B], condition, auxiliary) -> F``.

    x_t = latents*sigma_d; for each t (except the trailing 0):
        (i>0) x_t = sin(t)*sigma_d*z + cos(t)*x_t
        F = net(x_t/sigma_d, t, cond, aux);  x_t = cos(t)*x_t - sin(t)*sigma_d*F
    """
    ts = scm_time_grid(num_steps, sigma_min, sigma_max, sigma_data, intermediates).to(latents.device)
    b = latents.shape[0]
    x = latents * sigma_data
    for i, t in enumerate(ts[:-1]):
        if i > 0:
            z = noise_fn(x) if noise_fn is not None else torch.randn_like(x)
            x = torch.sin(t) * (sigma_data * z) + torch.cos(t) * x
        f = net(x / sigma_data, t.expand(b), condition, auxiliary)
        x = torch.cos(t) * x - torch.sin(t) * sigma_data * f
    return x


def dpm_solver_2s(net: Callable, latents: torch.Tensor, condition: torch.Tensor, auxiliary,
                  num_steps: int = 20, sigma_min: float = 0.02, sigma_max: float = 200.0,
                  sigma_data: float = 1.0) -> torch.Tensor:
    """generating/diffusion.py:355-415: TrigFlow 2nd-order (Heun) solver, 2*num_steps-1 net calls."""
    dev = latents.device
    lo, hi = torch.log(torch.tensor(sigma_min, device=dev)), torch.log(torch.tensor(sigma_max, device=dev))
    u = torch.linspace(1, 0, num_steps, device=dev)
    ts = torch.atan(torch.exp(lo + u * (hi - lo)) / sigma_data)
    ts = torch.cat([ts, torch.zeros(1, device=dev)])
    b = latents.shape[0]
    x = latents * sigma_data
    for k in range(num_steps):
        s, t = ts[k], ts[k + 1]
        delta = t - s
        f_s = net(x / sigma_data, s.repeat(b), condition, auxiliary)
        x_e = x + delta * sigma_data * f_s
        if k < num_steps - 1:
            f_t = net(x_e / sigma_data, t.repeat(b), condition, auxiliary)
            x = x + delta * sigma_data * 0.5 * (f_s + f_t)
        else:
            x = x_e
    return x


def rollout_step(sample: Callable, x_std: torch.Tensor, forcings_std: torch.Tensor, x_mean, x_std_dev, diff_std,
                 n_var: int):
    """One iteration of generate.py:97-136 (residual=True branch) on already-standardised forcings.

    X <- cat([X, forcings]);  Y = sampler(X);  X_phys = unstd_x(X)[:, :n_var] + Y*diff_std  (t_means are 0,
    data/era5.py:95-100);  X <- std_x(X_phys).  Returns (next standardised state, physical state).
    """
    xin = torch.cat([x_std, forcings_std], dim=1)
    y = sample(xin)
    x_phys = (x_std * x_std_dev + x_mean) + y * diff_std
    return (x_phys - x_mean) / x_std_dev, x_phys
