"""Seeded synthetic checkpoints and clips.

There are no pretrained weights in the reference tree (README.md:29 links an unreachable
share) and no network, so parity tests and benchmarks run on checkpoints that are a pure
function of (constructor kwargs, seed), written in the reference's ``.amlt`` state-dict
layout (``srcs/utils.py:91``): every tensor is drawn from a generator seeded by
``(seed, crc32(key))`` so values do not depend on key order.

Gotchas honoured: RVQ ``inited`` = 1 and non-zero codebooks (otherwise
``core_vq.py:209`` runs k-means in eval mode); ``diffusion.model.*`` aliases ``diff_model.*``.
"""
import zlib
from collections import OrderedDict

import torch

from .layout import state_dict_spec, SCHEDULE_BUFFERS
from .schedule import make_buffers


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _weight_v(seed, key, shape):
    # Conv1d [Cout,Cin,k]: fan_in = Cin*k.  ConvTranspose1d [Cin,Cout,k=2s]: two taps reach an output.
    fan_in = shape[1] * shape[2] if "convtr" not in key else shape[0] * 2
    return _uniform(shape, (3.0 / fan_in) ** 0.5, _gen(seed, key))


def make_state_dict(seed=0, **model_kwargs):
    """state_dict for DiffAudioRep(**model_kwargs), deterministic in `seed`."""
    spec = state_dict_spec(**model_kwargs)
    sd = OrderedDict()
    sched = None
    for key, shape in spec.items():
        if key.startswith("diffusion.model."):
            sd[key] = sd["diff_model." + key[len("diffusion.model."):]]      # aliased storage
            continue
        g = _gen(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if key.startswith("diffusion."):
            if sched is None:
                sched = make_buffers()
            assert leaf in SCHEDULE_BUFFERS
            t = sched[leaf]
        elif leaf == "weight_v":
            t = _weight_v(seed, key, shape)
        elif leaf == "weight_g":                                              # registered before weight_v
            kv = key[:-1] + "v"
            v = _weight_v(seed, kv, spec[kv])
            t = v.flatten(1).norm(dim=1).reshape(shape) * (0.8 + 0.4 * torch.rand(shape, generator=g))
        elif leaf == "inited":
            t = torch.ones(shape)
        elif leaf == "cluster_size":
            t = torch.ones(shape)
        elif leaf == "embed":
            q = int(key.split(".")[3])
            t = torch.randn(shape, generator=g) * (0.35 * 0.7 ** q)
        elif leaf == "embed_avg":
            t = sd[key[:-4]].clone()
        elif leaf == "g":                                                     # channel LayerNorm gain
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".norm." in key and leaf == "weight":                            # GroupNorm affine
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".norm." in key and leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif "lstm" in key:
            H = shape[0] // 4
            t = _uniform(shape, H ** -0.5, g)
        elif leaf == "weight":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if "upsampling_layers" in key:                                    # ConvTranspose1d [Cin,Cout,k], k=2s
                fan_in = shape[0] * 2
            t = _uniform(shape, (3.0 / fan_in) ** 0.5, g)
        elif leaf == "bias":
            t = _uniform(shape, 0.05, g)
        else:
            raise KeyError(f"synthetic init has no rule for {key}")
        sd[key] = t.to(torch.float32).contiguous()
    return sd


def make_clips(B, T=38400, seed=1234):
    """'LibriSpeech-shaped' clips: AR(2)-coloured Gaussian noise under a 3-6 Hz syllabic envelope,
    peak-normalised per clip like dataset_libri.py:48-52.  Returns fp32 [B,1,T] on CPU."""
    out = torch.empty(B, 1, T)
    n = torch.arange(T, dtype=torch.float32) / 16000.0
    for b in range(B):
        g = torch.Generator(device="cpu"); g.manual_seed(seed + b)
        e = torch.randn(T + 2, generator=g)
        # resonant AR(2) (pole radius .97, ~500-1500 Hz) via FFT-domain filtering
        f0 = 500.0 + 1000.0 * torch.rand(1, generator=g).item()
        r = 0.97
        a1, a2 = -2 * r * torch.cos(torch.tensor(2 * torch.pi * f0 / 16000.0)).item(), r * r
        E = torch.fft.rfft(e, n=2 * (T + 2))
        w = torch.arange(E.shape[0], dtype=torch.float32) * (2 * torch.pi / (2 * (T + 2)))
        Hden = 1 + a1 * torch.exp(-1j * w) + a2 * torch.exp(-2j * w)
        x = torch.fft.irfft(E / Hden, n=2 * (T + 2))[2:T + 2]
        rate = 3.0 + 3.0 * torch.rand(1, generator=g).item()
        ph = 2 * torch.pi * torch.rand(1, generator=g).item()
        env = 0.55 + 0.45 * torch.sin(2 * torch.pi * rate * n + ph)
        x = x * env
        out[b, 0] = x / (x.abs().max() + 1e-8)
    return out


def make_noise(n_steps, B, C, L, seed=0):
    """Pre-drawn DDPM noise, [n_steps-1, B, C, L], in the order the reference's
    ``torch.randn_like`` draws it inside halfway_sampling (ddpm_loss.py:249)."""
    g = torch.Generator(device="cpu"); g.manual_seed(seed)
    return torch.randn(max(n_steps - 1, 0), B, C, L, generator=g)
