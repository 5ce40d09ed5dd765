"""Round-2 golden vectors: long trajectories and the other samplers, produced by the REAL reference
(imported from /root/reference via oracle/ref_import.py) on seeded synthetic checkpoints.

Run in the build container only:
    python tests/golden/make_golden_r2.py [case ...]
Writes tests/golden/r2_<case>.pt (+ r2_index.json).  tests/test_oracle_golden.py pins the oracle to them on CPU;
tests/test_parity_long_gpu.py compares the CUDA path with the oracle AND with these vectors.

Cases (all through the reference's own methods; nothing here restates arithmetic):
  B_N50     halfway_sampling(t=50), Layout-B (enc_ratios 8 4), B=2            ddpm_loss.py:370-385  (BASELINE config 2's N)
  A_N200    halfway_sampling(t=200), Layout-A 1.5 kbps + unet_scale_cond, B=1 (BASELINE config 3's N)
  A_loop20  the first 20 steps (t = 999 … 980) of p_sample_loop from seeded noise: the reference's own p_sample called
            in the loop of ddpm_loss.py:258-263, stopped after 20 iterations
  A_full1000  diffusion.sample(batch_size=1, condition) = p_sample_loop, all 1000 steps, minimal T   :253-266, 305-309
  A_ddim20  ddim_sample with sampling_timesteps=20, eta=0                                           :268-303
  A_ddim10_eta  ddim_sample with sampling_timesteps=10, eta=0.5 (noise enters)
  A_infill  infilling(infill_img, condition, midway_t=6, lam=0.8)                                   :331-367
  A_loss    GaussianDiffusion1D.forward(x, cond, t, noise) = p_losses (training-loss forward)       :404-450
RNG: the global generator is seeded right before each reference call; the tensors the reference then draws
(randn(shape), randn_like per step, rand for infilling) are re-drawn here in the same order from the same seed and
stored implicitly by their seed (tests regenerate them with `draws()`); equality of the two streams is asserted.
"""
import json
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from oracle.ref_import import build_reference_models            # noqa: E402
from ladiffcodec_b200.config import sample_args                  # noqa: E402
from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs  # noqa: E402
from ladiffcodec_b200.synthetic import make_state_dict, make_clips  # noqa: E402

FLAGS_A3 = dict(run_diff=True, scaling_global=True, cond_bandwidth=3.0, unet_scale_cond=True, model_for_cond="c", model_path="m")
FLAGS_A15 = dict(run_diff=True, scaling_global=True, cond_bandwidth=1.5, unet_scale_cond=True, model_for_cond="c", model_path="m")
FLAGS_B3 = dict(run_diff=True, cond_bandwidth=3.0, enc_ratios=[8, 4], upsampling_ratios=[5, 2], model_for_cond="c", model_path="m")

CASES = {
    "B_N50": dict(flags=FLAGS_B3, T=5120, B=2, kind="halfway", n_steps=50),
    "A_N200": dict(flags=FLAGS_A15, T=2560, B=1, kind="halfway", n_steps=200),
    "A_loop20": dict(flags=FLAGS_A3, T=1280, B=2, kind="loop", n_steps=20),
    "A_full1000": dict(flags=FLAGS_A3, T=640, B=1, kind="sample", n_steps=1000),
    "A_ddim20": dict(flags=FLAGS_A3, T=1280, B=2, kind="ddim", sampling_timesteps=20, eta=0.0),
    "A_ddim10_eta": dict(flags=FLAGS_A3, T=1280, B=1, kind="ddim", sampling_timesteps=10, eta=0.5),
    "A_infill": dict(flags=FLAGS_A3, T=1280, B=1, kind="infill", midway_t=6, lam=0.8),
    "A_loss": dict(flags=FLAGS_A3, T=1280, B=2, kind="loss", t=[17, 640]),
}
SEED_MODEL, SEED_COND, SEED_WAV, SEED_NOISE = 11, 12, 13, 14
TRACE_EVERY = {"B_N50": 1, "A_N200": 4, "A_loop20": 1, "A_full1000": 20, "A_ddim20": 1, "A_ddim10_eta": 1}


def latent_len(args, T):
    L = T
    for r in args.enc_ratios:
        L //= r
    return L


def draws(case, L, seed=SEED_NOISE):
    """The tensors the reference draws from the global generator for this case, in its order."""
    B, kind = case["B"], case["kind"]
    torch.manual_seed(seed)
    shape = (B, 128, L)
    if kind == "halfway":
        return dict(noise=torch.randn(case["n_steps"] - 1, *shape))
    if kind == "loop":
        init = torch.randn(shape)
        return dict(init=init, noise=torch.randn(case["n_steps"], *shape))
    if kind == "sample":
        init = torch.randn(shape)
        return dict(init=init, noise=torch.randn(case["n_steps"] - 1, *shape))
    if kind == "ddim":
        init = torch.randn(shape)
        return dict(init=init, noise=torch.randn(case["sampling_timesteps"] - 1, *shape))
    if kind == "infill":
        init = torch.rand(shape)
        return dict(init=init, noise=torch.randn(2 * (case["midway_t"] - 1), *shape))
    if kind == "loss":
        return dict(noise=torch.randn(shape))
    raise KeyError(kind)


def summary(x):
    x = x.double()
    return dict(shape=list(x.shape), mean=x.mean().item(), std=x.std().item(), l2=x.norm().item(), absmax=x.abs().max().item())


def run_case(name, case):
    args = sample_args(**case["flags"])
    model, cond_model = build_reference_models(vars(args))
    sdm = make_state_dict(seed=SEED_MODEL, **ladiff_model_kwargs(args))
    sdc = make_state_dict(seed=SEED_COND, **cond_model_kwargs(args))
    model.load_state_dict(sdm, strict=True)
    cond_model.load_state_dict(sdc, strict=True)
    B, T, kind = case["B"], case["T"], case["kind"]
    L = latent_len(args, T)
    wav = make_clips(B, T, seed=SEED_WAV)
    d = draws(case, L)
    diff = model.diffusion
    trace, every = [], TRACE_EVERY.get(name, 0)
    orig_p_sample = diff.p_sample

    def traced_p_sample(x, t, condition=None, clip_denoised=True):
        out = orig_p_sample(x, t, condition, clip_denoised)
        trace.append(out[0].clone())
        return out

    fx = dict(flags=case["flags"], T=T, B=B, kind=kind, seeds=dict(model=SEED_MODEL, cond=SEED_COND, wav=SEED_WAV, noise=SEED_NOISE),
              case={k: v for k, v in case.items() if k != "flags"})
    with torch.no_grad():
        cond = cond_model.get_cond(wav)
        img = cond
        for layer in model.diff_model.upsampling_layers:
            img = layer(img)
        img = img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)         # sample.py:129 per clip
        fx["cond"] = cond.clone()
        if kind == "halfway":
            diff.p_sample = traced_p_sample
            torch.manual_seed(SEED_NOISE)
            latent = diff.halfway_sampling(img=img.clone(), condition=cond, t=case["n_steps"])
        elif kind == "loop":
            # ddpm_loss.py:253-263 verbatim, stopped after n_steps iterations
            torch.manual_seed(SEED_NOISE)
            x = torch.randn((B, 128, L))
            assert torch.equal(x, d["init"])
            for k, t in enumerate(reversed(range(0, diff.num_timesteps))):
                if k == case["n_steps"]:
                    break
                x, _ = diff.p_sample(x, t, cond)
                trace.append(x.clone())
            latent = x
        elif kind == "sample":
            diff.p_sample = traced_p_sample
            diff.seq_length = L                                                              # sample.py:90
            torch.manual_seed(SEED_NOISE)
            latent = diff.sample(batch_size=B, condition=cond)
        elif kind == "ddim":
            diff.sampling_timesteps = case["sampling_timesteps"]
            diff.ddim_sampling_eta = case["eta"]
            torch.manual_seed(SEED_NOISE)
            latent = diff.ddim_sample((B, 128, L), condition=cond)
        elif kind == "infill":
            diff.seq_length = L
            torch.manual_seed(SEED_NOISE)
            # `noise` is passed so that the reference's unused default(noise, randn_like) draw (ddpm_loss.py:352) does not
            # consume the generator; it plays no part in the arithmetic
            latent = diff.infilling(img.clone(), cond, midway_t=case["midway_t"], noise=torch.zeros(1), lam=case["lam"])
        elif kind == "loss":
            diff.seq_length = L
            t = torch.tensor(case["t"], dtype=torch.long)
            x_start = img.clone()
            torch.manual_seed(SEED_NOISE)
            loss, pred_x0, x_t, t_out = diff(x_start, cond, t=t)                             # forward → p_losses draws randn_like
            fx.update(loss=loss.clone(), pred_x_start=pred_x0.clone(), x_t=x_t.clone(), t=t_out.clone())
            latent = None
        diff.p_sample = orig_p_sample
        if latent is not None:
            fx["latent"] = latent.clone()
            fx["latent_sum"] = summary(latent)
            x = model.decoder(latent)
            x = x / (x.reshape(B, -1).std(1).reshape(B, 1, 1) + 1e-8)
            x = x / (x.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
            fx["wav_hat"] = x.clone()
    if kind == "sample":      # full intermediate states: the CPU test re-runs only segments of the 1000-step chain
        fx["states"] = {k: trace[k - 1].clone() for k in (20, 500, 980)}
    if trace and every:
        fx["trace_every"] = every
        fx["trace_l2"] = torch.tensor([t.double().norm().item() for t in trace])
        fx["trace_sub"] = torch.stack([t[:, ::8, ::8] for t in trace[every - 1::every]]).clone()
    path = os.path.join(HERE, f"r2_{name}.pt")
    torch.save(fx, path)
    info = dict(bytes=os.path.getsize(path), kind=kind, L=L)
    if latent is not None:
        info["latent"] = fx["latent_sum"]
    else:
        info["loss"] = float(fx["loss"])
    print(name, info, flush=True)
    return info


def main():
    torch.set_num_threads(os.cpu_count())
    names = sys.argv[1:] or list(CASES)
    ipath = os.path.join(HERE, "r2_index.json")
    index = json.load(open(ipath)) if os.path.exists(ipath) else {}
    for name in names:
        index[name] = run_case(name, CASES[name])
        with open(ipath, "w") as f:
            json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
