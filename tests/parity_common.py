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

# Stated tolerances: what was measured on the B200 (profiles/r2a/parity_report_*.json) times ~2.
# fp32 SIMT codec stages: a few 1e-5 of the tensor's scale.  UNet: 16-bit tensor-core operands and stored activations with fp32
# accumulation -> relative L2 per evaluation, and per trajectory at the step counts the benchmarks run.
_TOL_F16 = dict(
    codec_abs=5e-5, upsample_abs=2e-5,
    unet_rel_l2=4e-3,               # one UNet evaluation vs the fp32 oracle: measured 1.4e-3 (config 2 full size), 1.3-1.6e-3 (goldens)
    unet_simt_vs_tc_rel_l2=4e-3,    # tensor-core path vs SIMT check kernel on the same operands (summation order + rounding flips)
    latent_rel_l2=3e-4, wav_snr_db=70.0,                 # 2-3 step trajectories (round-1 goldens, edge shapes, 5 s utterance): measured <= 8.3e-5 / >= 78.7 dB
    latent_rel_l2_long=2e-3, wav_snr_db_long=54.0,       # N = 50 / 200, first 20 steps from noise, infilling: measured <= 7.6e-4 / >= 60.5 dB
    latent_rel_l2_1000=2.5e-3, wav_snr_db_1000=50.0,     # p_sample_loop, all 1000 steps: measured 1.2e-3 / 56.5 dB
    latent_rel_l2_ddim=5e-2, wav_snr_db_ddim=26.0,       # ddim_sample 20 steps eta 0 / 10 steps eta .5: measured 2.5e-2 / 32.2 dB (no fresh
                                                         # noise to wash errors out, 50-100 timestep jumps, x0 clamp flips at t ~ 999)
    loss_abs=2e-3, pred_x0_rel_l2=1e-2)
_TOL_BF16 = dict(_TOL_F16, unet_rel_l2=3e-2, unet_simt_vs_tc_rel_l2=3e-2, latent_rel_l2=2e-2, wav_snr_db=25.0, latent_rel_l2_long=1.5e-2,
                 wav_snr_db_long=38.0, latent_rel_l2_1000=2e-2, wav_snr_db_1000=32.0, latent_rel_l2_ddim=0.3, wav_snr_db_ddim=12.0,
                 loss_abs=1e-2, pred_x0_rel_l2=5e-2)


class _Tol:
    """Tolerances of the loaded build: fp16 (default) or the -DLADIFF_USE_BF16 A/B build."""
    def _table(self):
        from ladiffcodec_b200 import _lib
        import torch as _t
        return _TOL_F16 if _lib.act_dtype() == _t.float16 else _TOL_BF16

    def __getitem__(self, k):
        return self._table()[k]


TOL = _Tol()
MEASURED = {}          # test name -> figures; dumped to gpurun_out/parity_report.json by tests/conftest.py at session end


def record(key, **vals):
    MEASURED.setdefault(key, {}).update(vals)


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
