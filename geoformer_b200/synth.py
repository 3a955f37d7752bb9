"""Deterministic synthetic checkpoints and image pairs.

There is no network (no released ``geoformer.ckpt``, no HPatches), so every
test and benchmark runs on random-init weights of the reference architecture
and synthetic 640x480 pairs.  The generator is independent of
the reference's constructor so that the *same* state dict can be produced on
the GPU box (where ``/root/reference`` does not exist) and loaded into the
reference model in the build container when golden vectors are made.

Key names / shapes follow the reference checkpoint schema:
253 tensors, 14 187 504 parameters.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

BLOCK_DIMS = (128, 196, 256)
D_COARSE, D_FINE = 256, 128


def _spec() -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every checkpoint entry, in state-dict order."""
    s: List[Tuple[str, Tuple[int, ...], str]] = []

    def conv(name, co, ci, k):
        s.append((name + ".weight", (co, ci, k, k), "conv"))

    def bn(name, c):
        s.extend([(name + ".weight", (c,), "bn_w"), (name + ".bias", (c,), "bn_b"),
                  (name + ".running_mean", (c,), "bn_m"), (name + ".running_var", (c,), "bn_v"),
                  (name + ".num_batches_tracked", (), "count")])

    b = "backbone."
    conv(b + "conv1", 128, 1, 7)
    bn(b + "bn1", 128)
    cin = 128
    for li, dim in enumerate(BLOCK_DIMS, start=1):
        for bi in range(2):
            p = f"{b}layer{li}.{bi}"
            conv(p + ".conv1", dim, cin if bi == 0 else dim, 3)
            conv(p + ".conv2", dim, dim, 3)
            bn(p + ".bn1", dim)
            bn(p + ".bn2", dim)
            if bi == 0 and li > 1:
                conv(p + ".downsample.0", dim, cin, 1)
                bn(p + ".downsample.1", dim)
        cin = dim
    conv(b + "layer3_outconv", 256, 256, 1)
    conv(b + "layer2_outconv", 256, 196, 1)
    conv(b + "layer2_outconv2.0", 256, 256, 3)
    bn(b + "layer2_outconv2.1", 256)
    conv(b + "layer2_outconv2.3", 196, 256, 3)
    conv(b + "layer1_outconv", 196, 128, 1)
    conv(b + "layer1_outconv2.0", 196, 196, 3)
    bn(b + "layer1_outconv2.1", 196)
    conv(b + "layer1_outconv2.3", 128, 196, 3)

    def encoder(prefix, n_layers, d):
        for i in range(n_layers):
            p = f"{prefix}.layers.{i}"
            for nm in ("q_proj", "k_proj", "v_proj", "merge"):
                s.append((f"{p}.{nm}.weight", (d, d), "xavier"))
            s.append((f"{p}.mlp.0.weight", (2 * d, 2 * d), "xavier"))
            s.append((f"{p}.mlp.2.weight", (d, 2 * d), "xavier"))
            for nm in ("norm1", "norm2"):
                s.append((f"{p}.{nm}.weight", (d,), "ln_w"))
                s.append((f"{p}.{nm}.bias", (d,), "ln_b"))

    encoder("loftr_coarse", 8, D_COARSE)
    s.append(("fine_preprocess.down_proj.weight", (D_FINE, D_COARSE), "kaiming_lin"))
    s.append(("fine_preprocess.down_proj.bias", (D_FINE,), "lin_bias_256"))
    s.append(("fine_preprocess.merge_feat.weight", (D_FINE, 2 * D_FINE), "kaiming_lin"))
    s.append(("fine_preprocess.merge_feat.bias", (D_FINE,), "lin_bias_256"))
    encoder("loftr_fine", 2, D_FINE)
    encoder("geo_module.des_transformer", 4, D_COARSE)
    s.append(("geo_module.des_transformer.norm.weight", (D_COARSE,), "ln_w"))   # unused in forward
    s.append(("geo_module.des_transformer.norm.bias", (D_COARSE,), "ln_b"))
    return s


def make_state_dict(seed: int = 0, randomize_norm: bool = False) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's init *distributions*
    (kaiming fan_out for convs / fine_preprocess, xavier-uniform for
    transformer linears; resnet_fpn.py:85-90, transformer.py:77-80,
    fine_preprocess.py:25-28).  ``randomize_norm`` perturbs BN/LN affine
    parameters and BN running stats so that folding bugs cannot hide behind
    the identity defaults (used by parity tests; benchmarks keep defaults)."""
    out: Dict[str, torch.Tensor] = {}
    for idx, (key, shape, kind) in enumerate(_spec()):
        g = torch.Generator().manual_seed(seed * 100003 + idx)
        if kind == "conv":
            std = math.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))
            t = torch.randn(shape, generator=g) * std
        elif kind == "xavier":
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "kaiming_lin":
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / shape[0])
        elif kind == "lin_bias_256":
            t = (torch.rand(shape, generator=g) * 2 - 1) / 16.0
        elif kind in ("bn_w", "ln_w"):
            t = 0.75 + 0.5 * torch.rand(shape, generator=g) if randomize_norm else torch.ones(shape)
        elif kind in ("bn_b", "ln_b", "bn_m"):
            t = 0.1 * torch.randn(shape, generator=g) if randomize_norm else torch.zeros(shape)
        elif kind == "bn_v":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g) if randomize_norm else torch.ones(shape)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.long)
        else:  # pragma: no cover
            raise KeyError(kind)
        out[key] = t
    return out


def make_image(h: int, w: int, seed: int) -> torch.Tensor:
    """Uniform-noise grayscale image [1,1,h,w] in [0,1)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(1, 1, h, w, generator=g)


def make_pairs(n: int, h: int, w: int, regime: str = "dense", seed0: int = 0):
    """``n`` synthetic pairs -> (image0 [n,1,h,w], image1 [n,1,h,w]).

    regimes: 'dense' image1 == image0 (M_c ~ 0.79 L, the heavy case);
    'shift' image1 = roll(image0, (8a, 8b)) (valid translation homography, few hundred matches);
    'mixed' sample i: i % 3 == 0 dense, == 1 an UNRELATED image (a handful of noise matches: the no-homography /
    garbage-homography branches of geo_module.py:45-94 inside one batch), == 2 shift."""
    im0 = torch.cat([make_image(h, w, seed0 + i) for i in range(n)], 0)
    shift = lambda i: torch.roll(im0[i], (8 * (1 + i % 2), 8 * (2 - i % 2)), (1, 2))
    if regime == "dense":
        im1 = im0.clone()
    elif regime == "shift":
        im1 = torch.stack([shift(i) for i in range(n)], 0)
    elif regime == "mixed":
        im1 = torch.stack([im0[i].clone() if i % 3 == 0 else (make_image(h, w, seed0 + 1000 + i)[0] if i % 3 == 1 else shift(i))
                           for i in range(n)], 0)
    else:
        raise ValueError(regime)
    return im0, im1
