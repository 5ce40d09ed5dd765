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
           wav_snr_db=25.0)


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
