"""Shared helpers for the GPU parity tests: build the CUDA models and the oracle inputs from the same
seeded synthetic checkpoints / clips / noise that produced tests/golden/*.pt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ladiffcodec_b200.config import sample_args                                   # noqa: E402
from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs        # noqa: E402
from ladiffcodec_b200.synthetic import make_state_dict, make_clips                # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# Stated tolerances.  fp32 SIMT codec stages: a few 1e-5 of the tensor's scale.
# UNet: bf16 operands / bf16 activations with fp32 accumulation -> relative L2 per evaluation.
TOL = dict(codec_abs=5e-5, upsample_abs=2e-5, unet_rel_l2=3e-2, unet_simt_vs_tc_rel_l2=3e-2, latent_rel_l2=2e-2,
           wav_snr_db=25.0,
           # long trajectories (N = 50 / 200 / 1000, other samplers): PROVISIONAL until measured on the B200 — see DESIGN.md §2
           latent_rel_l2_long=0.2, wav_snr_db_long=15.0, latent_rel_l2_1000=0.5, wav_snr_db_1000=5.0, loss_abs=5e-3)


def ptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def stream():
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def snr_db(x, ref):
    x, ref = x.double().cpu(), ref.double().cpu()
    return (10 * torch.log10(ref.pow(2).sum() / ((x - ref).pow(2).sum() + 1e-30))).item()


def case_setup(name):
    fx = load_golden(name)
    args = sample_args(**fx["flags"])
    sdm = make_state_dict(seed=fx["seeds"]["model"], **ladiff_model_kwargs(args))
    sdc = make_state_dict(seed=fx["seeds"]["cond"], **cond_model_kwargs(args))
    wav = make_clips(fx["B"], fx["T"], seed=fx["seeds"]["wav"])
    L = fx["T"]
    for r in args.enc_ratios:
        L //= r
    torch.manual_seed(fx["seeds"]["noise"])
    noise = torch.randn(max(fx["n_steps"] - 1, 0), fx["B"], 128, L)
    return fx, args, sdm, sdc, wav, noise


def cuda_models(args, sdm, sdc):
    from ladiffcodec_b200.model import DiffAudioRep
    from ladiffcodec_b200.utils import load_model
    m = DiffAudioRep(**ladiff_model_kwargs(args)).to("cuda")
    load_model(m, sdm, strict=True)
    c = DiffAudioRep(**cond_model_kwargs(args)).to("cuda")
    load_model(c, sdc)
    return m.eval(), c.eval()


def unet_kwargs(args):
    return dict(dim=args.diff_dims, upsampling_ratios=tuple(args.upsampling_ratios), unet_scale_cond=args.unet_scale_cond)


# ---------------------------------------------------------------------------------------------- round-2 cases
def r2_draws(case, B, L, seed):
    """The tensors the reference draws from torch's global generator for a round-2 golden case, in its order
    (tests/golden/make_golden_r2.py:draws — kept in sync by test_oracle_golden_r2.py, which fails if the streams differ)."""
    kind = case["kind"]
    torch.manual_seed(seed)
    shape = (B, 128, L)
    if kind == "halfway":
        return dict(noise=torch.randn(case["n_steps"] - 1, *shape))
    if kind == "loop":
        return dict(init=torch.randn(shape), noise=torch.randn(case["n_steps"], *shape))
    if kind == "sample":
        return dict(init=torch.randn(shape), noise=torch.randn(case["n_steps"] - 1, *shape))
    if kind == "ddim":
        return dict(init=torch.randn(shape), noise=torch.randn(case["sampling_timesteps"] - 1, *shape))
    if kind == "infill":
        return dict(init=torch.rand(shape), noise=torch.randn(2 * (case["midway_t"] - 1), *shape))
    if kind == "loss":
        return dict(noise=torch.randn(shape))
    raise KeyError(kind)


def r2_setup(name):
    """Checkpoints, clips, condition-independent inputs and the reference's random draws of round-2 golden case `name`."""
    fx = load_golden("r2_" + name)
    args = sample_args(**fx["flags"])
    sdm = make_state_dict(seed=fx["seeds"]["model"], **ladiff_model_kwargs(args))
    sdc = make_state_dict(seed=fx["seeds"]["cond"], **cond_model_kwargs(args))
    wav = make_clips(fx["B"], fx["T"], seed=fx["seeds"]["wav"])
    L = fx["T"]
    for r in args.enc_ratios:
        L //= r
    d = r2_draws(fx["case"], fx["B"], L, fx["seeds"]["noise"])
    return fx, args, sdm, sdc, wav, L, d


def normalized_img(cond, sdm, args):
    """sample.py:125-129 on the oracle: upsampled condition, per-clip max-normalised."""
    from oracle import ladiff_oracle as O
    B = cond.shape[0]
    with torch.no_grad():
        img = O.cond_upsample(cond, sdm, args.upsampling_ratios)
    return img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
