"""Checkpoint layout of the reference (``torch.save(model.state_dict())`` → ``*.amlt``,
``srcs/utils.py:85-108``) restated as a pure function of the constructor arguments.

``state_dict_spec(cfg)`` enumerates every key and shape a ``DiffAudioRep`` built with ``cfg``
(``srcs/model.py:34-106``) stores, in registration order, without importing the reference.
It is what the strict loader checks against and what the synthetic-checkpoint writer fills.
tests/test_layout.py pins it against the key lists dumped from the real reference.
"""
from collections import OrderedDict

import numpy as np

from .config import MODEL_DEFAULTS, UNET_DIM_MULTS, NUM_TIMESTEPS, CODEBOOK_BINS, ATTN_HEADS, ATTN_DIM_HEAD, \
    num_quantizers

SCHEDULE_BUFFERS = (
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
    "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
    "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
    "p2_loss_weight",
)  # ddpm_loss.py:140-168, registration order


def model_cfg(**kwargs):
    """DiffAudioRep's effective constructor arguments (unknown keys are swallowed by **kwargs,
    model.py:34 — e.g. sample.py:63 passes ``ratios=`` which is ignored)."""
    cfg = {k: (list(v) if isinstance(v, list) else v) for k, v in MODEL_DEFAULTS.items()}
    for k, v in kwargs.items():
        if k in cfg:
            cfg[k] = v
    return cfg


def _wn_conv(sd, prefix, cout, cin, k):          # NormConv1d + weight_norm: conv.py:137, g per OUT channel
    sd[prefix + ".conv.conv.bias"] = (cout,)
    sd[prefix + ".conv.conv.weight_g"] = (cout, 1, 1)
    sd[prefix + ".conv.conv.weight_v"] = (cout, cin, k)


def _wn_convtr(sd, prefix, cin, cout, k):        # NormConvTranspose1d: conv.py:171, g per IN channel
    sd[prefix + ".convtr.convtr.bias"] = (cout,)
    sd[prefix + ".convtr.convtr.weight_g"] = (cin, 1, 1)
    sd[prefix + ".convtr.convtr.weight_v"] = (cin, cout, k)


def _resblock(sd, prefix, dim):                  # seanet.py:36-60 (kernel_sizes [3,1], compress 2, true_skip False)
    _wn_conv(sd, prefix + ".block.1", dim // 2, dim, 3)
    _wn_conv(sd, prefix + ".block.3", dim, dim // 2, 1)
    _wn_conv(sd, prefix + ".shortcut", dim, dim, 1)


def _lstm(sd, prefix, dim, layers):              # lstm.py:20 — nn.LSTM(dim, dim, layers)
    for l in range(layers):
        sd[f"{prefix}.lstm.weight_ih_l{l}"] = (4 * dim, dim)
        sd[f"{prefix}.lstm.weight_hh_l{l}"] = (4 * dim, dim)
        sd[f"{prefix}.lstm.bias_ih_l{l}"] = (4 * dim,)
        sd[f"{prefix}.lstm.bias_hh_l{l}"] = (4 * dim,)


def encoder_spec(cfg, sd=None, prefix="encoder"):
    """seanet.py:91-151."""
    sd = OrderedDict() if sd is None else sd
    nf, dim, nres = cfg["n_filters"], cfg["rep_dims"], cfg["n_residual_layers"]
    assert nres == 1, "n_residual_layers != 1 is outside the hot path"
    ratios = list(reversed(cfg["enc_ratios"]))
    i, mult = 0, 1
    _wn_conv(sd, f"{prefix}.model.{i}", nf, 1, 7); i += 1
    for r in ratios:
        _resblock(sd, f"{prefix}.model.{i}", mult * nf); i += 1
        i += 1                                                          # ELU
        _wn_conv(sd, f"{prefix}.model.{i}", mult * nf * 2, mult * nf, 2 * r); i += 1
        mult *= 2
    if cfg["lstm"]:
        _lstm(sd, f"{prefix}.model.{i}", mult * nf, cfg["lstm"]); i += 1
    i += 1                                                              # ELU
    _wn_conv(sd, f"{prefix}.model.{i}", dim, mult * nf, 7)
    return sd


def decoder_spec(cfg, sd=None, prefix="decoder"):
    """seanet.py:184-244."""
    sd = OrderedDict() if sd is None else sd
    nf, dim = cfg["n_filters"], cfg["rep_dims"]
    ratios = list(cfg["enc_ratios"])
    mult = 2 ** len(ratios)
    i = 0
    _wn_conv(sd, f"{prefix}.model.{i}", mult * nf, dim, 7); i += 1
    if cfg["lstm"]:
        _lstm(sd, f"{prefix}.model.{i}", mult * nf, cfg["lstm"]); i += 1
    for r in ratios:
        i += 1                                                          # ELU
        _wn_convtr(sd, f"{prefix}.model.{i}", mult * nf, mult * nf // 2, 2 * r); i += 1
        _resblock(sd, f"{prefix}.model.{i}", mult * nf // 2); i += 1
        mult //= 2
    i += 1                                                              # ELU
    _wn_conv(sd, f"{prefix}.model.{i}", 1, nf, 7)
    return sd


def quantizer_spec(cfg, sd=None, prefix="quantizer"):
    """vq.py:59-67, core_vq.py:125-134."""
    sd = OrderedDict() if sd is None else sd
    hop = int(np.prod(cfg["enc_ratios"]))
    n_q = num_quantizers(cfg["bandwidth"], hop, cfg["sample_rate"])
    for q in range(n_q):
        p = f"{prefix}.vq.layers.{q}._codebook"
        sd[p + ".inited"] = (1,)
        sd[p + ".cluster_size"] = (CODEBOOK_BINS,)
        sd[p + ".embed"] = (CODEBOOK_BINS, cfg["rep_dims"])
        sd[p + ".embed_avg"] = (CODEBOOK_BINS, cfg["rep_dims"])
    return sd


def unet_dims(dim):
    dims = [dim] + [dim * m for m in UNET_DIM_MULTS]
    return dims, list(zip(dims[:-1], dims[1:]))


def _unet_resnet(sd, p, cin, cout, time_dim):    # unet.py:156-174
    sd[p + ".mlp.1.weight"] = (2 * cout, time_dim)
    sd[p + ".mlp.1.bias"] = (2 * cout,)
    for b, ci in (("block1", cin), ("block2", cout)):
        sd[f"{p}.{b}.proj.weight"] = (cout, ci, 3)
        sd[f"{p}.{b}.proj.bias"] = (cout,)
        sd[f"{p}.{b}.norm.weight"] = (cout,)
        sd[f"{p}.{b}.norm.bias"] = (cout,)
    if cin != cout:
        sd[p + ".res_conv.weight"] = (cout, cin, 1)
        sd[p + ".res_conv.bias"] = (cout,)


def _unet_linattn(sd, p, dim):                   # unet.py:194-206 inside Residual(PreNorm(.))
    hid = ATTN_HEADS * ATTN_DIM_HEAD
    sd[p + ".fn.fn.to_qkv.weight"] = (3 * hid, dim, 1)
    sd[p + ".fn.fn.to_out.0.weight"] = (dim, hid, 1)
    sd[p + ".fn.fn.to_out.0.bias"] = (dim,)
    sd[p + ".fn.fn.to_out.1.g"] = (1, dim, 1)
    sd[p + ".fn.norm.g"] = (1, dim, 1)


def unet_spec(cfg, sd=None, prefix="diff_model"):
    """unet.py:251-377 with the arguments model.py:74 passes."""
    sd = OrderedDict() if sd is None else sd
    dim, inp = cfg["diff_dims"], cfg["rep_dims"]
    cc = cfg["cond_channels"]
    in_ch = inp + cc if cfg["other_cond"] else inp * (2 if (cfg["self_condition"] or cfg["qtz_condition"]) else 1)
    time_dim = dim * 4
    dims, in_out = unet_dims(dim)
    sd[f"{prefix}.init_conv.weight"] = (dim, in_ch, 7)
    sd[f"{prefix}.init_conv.bias"] = (dim,)
    sd[f"{prefix}.time_mlp.1.weight"] = (time_dim, dim)
    sd[f"{prefix}.time_mlp.1.bias"] = (time_dim,)
    sd[f"{prefix}.time_mlp.3.weight"] = (time_dim, time_dim)
    sd[f"{prefix}.time_mlp.3.bias"] = (time_dim,)
    n = len(in_out)
    for i, (di, do) in enumerate(in_out):
        p = f"{prefix}.downs.{i}"
        _unet_resnet(sd, p + ".0", di, di, time_dim)
        _unet_resnet(sd, p + ".1", di, di, time_dim)
        _unet_linattn(sd, p + ".2", di)
        sd[p + ".3.weight"] = (do, di, 4 if i < n - 1 else 3)
        sd[p + ".3.bias"] = (do,)
    for i, (di, do) in enumerate(reversed(in_out)):
        p = f"{prefix}.ups.{i}"
        _unet_resnet(sd, p + ".0", do + di, do, time_dim)
        _unet_resnet(sd, p + ".1", do + di, do, time_dim)
        _unet_linattn(sd, p + ".2", do)
        q = p + (".3.1" if i < n - 1 else ".3")
        sd[q + ".weight"] = (di, do, 3)
        sd[q + ".bias"] = (di,)
    mid = dims[-1]
    hid = ATTN_HEADS * ATTN_DIM_HEAD
    _unet_resnet(sd, f"{prefix}.mid_block1", mid, mid, time_dim)
    sd[f"{prefix}.mid_attn.fn.fn.to_qkv.weight"] = (3 * hid, mid, 1)
    sd[f"{prefix}.mid_attn.fn.fn.to_out.weight"] = (mid, hid, 1)
    sd[f"{prefix}.mid_attn.fn.fn.to_out.bias"] = (mid,)
    sd[f"{prefix}.mid_attn.fn.norm.g"] = (1, mid, 1)
    _unet_resnet(sd, f"{prefix}.mid_block2", mid, mid, time_dim)
    _unet_resnet(sd, f"{prefix}.final_res_block", dim * 2, dim, time_dim)
    sd[f"{prefix}.final_conv.weight"] = (inp, dim, 1)
    sd[f"{prefix}.final_conv.bias"] = (inp,)
    if cfg["other_cond"] and cfg["upsampling_ratios"] is not None:
        for j, r in enumerate(cfg["upsampling_ratios"]):       # unet.py:372-377 — no weight-norm
            sd[f"{prefix}.upsampling_layers.{j}.convtr.convtr.weight"] = (cc, cc, 2 * r)
            sd[f"{prefix}.upsampling_layers.{j}.convtr.convtr.bias"] = (cc,)
    return sd


def state_dict_spec(**kwargs):
    """Ordered {key: shape} for DiffAudioRep(**kwargs).state_dict()."""
    cfg = model_cfg(**kwargs)
    sd = OrderedDict()
    encoder_spec(cfg, sd)
    decoder_spec(cfg, sd)
    if cfg["quantization"]:
        quantizer_spec(cfg, sd)
    if cfg["run_diff"]:
        if cfg["model_type"] != "unet":
            raise NotImplementedError("model_type must be 'unet' on the sampling path (model.py:73)")
        unet_spec(cfg, sd, "diff_model")
        for name in SCHEDULE_BUFFERS:                          # GaussianDiffusion1D buffers come first …
            sd["diffusion." + name] = (NUM_TIMESTEPS,)
        alias = OrderedDict()
        unet_spec(cfg, alias, "diffusion.model")               # … then the aliased UNet (same storage)
        # nn.Module.state_dict emits a module's own buffers after its parameters but before
        # child modules; GaussianDiffusion1D has no parameters, so buffers precede `model.*`.
        sd.update(alias)
    return sd


def cond_model_kwargs(args):
    """sample.py:63 — the kwargs of the conditioning codec (``ratios=`` is swallowed)."""
    a = vars(args) if not isinstance(args, dict) else args
    return dict(rep_dims=a["rep_dims"], emb_dims=a["emb_dims"], n_residual_layers=a["n_residual_layers"],
                n_filters=a["n_filters"], lstm=a["lstm"], quantization=True, bandwidth=a["cond_bandwidth"],
                ratios=a["cond_enc_ratios"], final_activation=a["final_activation"])


def ladiff_model_kwargs(args):
    """sample.py:52-56."""
    a = dict(vars(args) if not isinstance(args, dict) else args)
    a["other_cond"] = bool(a.get("model_for_cond"))
    return a
