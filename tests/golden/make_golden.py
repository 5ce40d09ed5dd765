"""Generates tests/golden/*.pt by running the REAL reference (imported from /root/reference via
oracle/ref_import.py) on seeded synthetic checkpoints, clips and pre-drawn noise.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
The fixtures it writes are committed; tests/test_oracle_golden.py (CPU) pins the oracle to them
and tests/test_parity_gpu.py (GPU) compares the CUDA path with both.

Every case drives exactly the calls of srcs/sample.py:94-134: get_cond → upsampling_layers →
max-normalise → halfway_sampling(t=n_steps) → decoder → std/max normalise; plus one bare
Unet1D.forward at a fixed t.  torch.manual_seed + randn_like inside the reference is replaced
by nothing: we seed the global generator so that the reference's own draws equal the
pre-drawn tensor (verified below).
"""
import hashlib
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
from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs, state_dict_spec  # noqa: E402
from ladiffcodec_b200.synthetic import make_state_dict, make_clips  # noqa: E402

CASES = {
    # README layout ("Layout-A"): enc_ratios [8], upsampling [5,4,2], 3 kbps, unet_scale_cond
    "A_3kbps": dict(flags=dict(run_diff=True, scaling_global=True, cond_bandwidth=3.0, unet_scale_cond=True,
                               model_for_cond="c", model_path="m"), T=5120, B=1, n_steps=3, t_probe=37),
    # BASELINE config 2 layout ("Layout-B"): enc_ratios [8,4], upsampling [5,2], no cond scaling
    "B_3kbps": dict(flags=dict(run_diff=True, cond_bandwidth=3.0, enc_ratios=[8, 4], upsampling_ratios=[5, 2],
                               model_for_cond="c", model_path="m"), T=5120, B=2, n_steps=3, t_probe=99),
    # 1.5 kbps (n_q = 3), Layout-A
    "A_1p5kbps": dict(flags=dict(run_diff=True, scaling_global=True, cond_bandwidth=1.5, unet_scale_cond=True,
                                 model_for_cond="c", model_path="m"), T=2560, B=1, n_steps=2, t_probe=0),
}
SEED_MODEL, SEED_COND, SEED_WAV, SEED_NOISE = 11, 12, 13, 14


def summary(x):
    x = x.double()
    return dict(shape=list(x.shape), mean=x.mean().item(), std=x.std().item(), l2=x.norm().item(),
                absmax=x.abs().max().item())


def key_hash(spec):
    s = "\n".join(f"{k}:{tuple(v)}" for k, v in spec.items())
    return hashlib.sha256(s.encode()).hexdigest()


def main():
    torch.set_num_threads(os.cpu_count())
    index = {}
    for name, case in CASES.items():
        args = sample_args(**case["flags"])
        model, cond_model = build_reference_models(vars(args))
        sdm = make_state_dict(seed=SEED_MODEL, **ladiff_model_kwargs(args))
        sdc = make_state_dict(seed=SEED_COND, **cond_model_kwargs(args))
        # layout pin: the real reference's own state_dict keys/shapes
        ref_keys_m = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        ref_keys_c = {k: tuple(v.shape) for k, v in cond_model.state_dict().items()}
        model.load_state_dict(sdm, strict=True)
        cond_model.load_state_dict(sdc, strict=True)
        B, T, N = case["B"], case["T"], case["n_steps"]
        wav = make_clips(B, T, seed=SEED_WAV)
        with torch.no_grad():
            z = cond_model.encoder(wav)
            qr = cond_model.quantizer(z, sample_rate=cond_model.frame_rate, bandwidth=cond_model.bandwidth)
            cond = cond_model.get_cond(wav)
            assert torch.equal(cond, qr.quantized)
            img = cond
            for layer in model.diff_model.upsampling_layers:
                img = layer(img)
            img_raw = img.clone()
            # per-clip normalisation == sample.py:129 with B=1
            img = img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
            tp = torch.full((B,), case["t_probe"], dtype=torch.long)
            eps = model.diff_model(img, tp, cond)
            L = img.shape[-1]
            torch.manual_seed(SEED_NOISE)
            noise = torch.randn(max(N - 1, 0), B, 128, L)
            torch.manual_seed(SEED_NOISE)   # the reference draws randn_like(img) per step: same stream
            latent = model.diffusion.halfway_sampling(img=img.clone(), condition=cond, t=N)
            x = model.decoder(latent)
            dec_raw = x.clone()
            x = x / (x.reshape(B, -1).std(1).reshape(B, 1, 1) + 1e-8)
            x = x / (x.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
            if B == 1:   # cross-check against the literal sample.py:129-134 expressions
                y = dec_raw.clone()
                y /= torch.std(y.flatten()) + 1e-8
                y /= torch.max(torch.abs(y.flatten())) + 1e-8
                assert torch.equal(x, y)
        fx = dict(
            flags=case["flags"], T=T, B=B, n_steps=N, t_probe=case["t_probe"],
            seeds=dict(model=SEED_MODEL, cond=SEED_COND, wav=SEED_WAV, noise=SEED_NOISE),
            codes=qr.codes.to(torch.int16), cond=cond.clone(), enc_z=z.clone(),
            img_raw_sub=img_raw[:, ::4, ::4].clone(), img_raw_sum=summary(img_raw),
            eps_sub=eps[:, ::4, ::4].clone(), eps_sum=summary(eps),
            latent_sub=latent[:, ::4, ::4].clone(), latent_sum=summary(latent),
            dec_raw_sum=summary(dec_raw), wav_hat=x.clone(),
            keys_model_sha=key_hash(ref_keys_m), keys_cond_sha=key_hash(ref_keys_c),
            n_keys_model=len(ref_keys_m), n_keys_cond=len(ref_keys_c),
        )
        # sanity: our spec reproduces the reference's key list (order included)
        assert key_hash(state_dict_spec(**ladiff_model_kwargs(args))) == fx["keys_model_sha"], name
        assert key_hash(state_dict_spec(**cond_model_kwargs(args))) == fx["keys_cond_sha"], name
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(fx, path)
        index[name] = dict(bytes=os.path.getsize(path), eps=fx["eps_sum"], latent=fx["latent_sum"],
                           n_q=int(qr.codes.shape[0]))
        print(name, index[name])
        if name == "A_3kbps":   # the cond codec key list in clear, for humans
            with open(os.path.join(HERE, "cond_codec_keys.json"), "w") as f:
                json.dump({k: list(v) for k, v in ref_keys_c.items()}, f, indent=0)
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
