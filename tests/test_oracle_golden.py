"""CPU: pins the oracle restatement (oracle/ladiff_oracle.py) to golden vectors produced by the
REAL reference (tests/golden/make_golden.py) — stage by stage, same seeded checkpoints, clips
and pre-drawn noise.  Tolerances: integer codes and the RVQ lookup are bit-exact; float stages
agree to a few fp32 ulps of their scale (same ATen primitives, same order)."""
import pytest
import torch

from conftest import load_golden
from oracle import ladiff_oracle as O
from ladiffcodec_b200.config import sample_args
from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs
from ladiffcodec_b200.synthetic import make_state_dict, make_clips, make_noise

CASES = ["A_3kbps", "B_3kbps", "A_1p5kbps"]


def _setup(name):
    fx = load_golden(name)
    args = sample_args(**fx["flags"])
    sdm = make_state_dict(seed=fx["seeds"]["model"], **ladiff_model_kwargs(args))
    sdc = make_state_dict(seed=fx["seeds"]["cond"], **cond_model_kwargs(args))
    wav = make_clips(fx["B"], fx["T"], seed=fx["seeds"]["wav"])
    return fx, args, sdm, sdc, wav


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    fx, args, sdm, sdc, wav = _setup(name)
    torch.manual_seed(0)
    with torch.no_grad():
        cond, codes, z = O.get_cond(wav, sdc, args.cond_bandwidth, return_codes=True)
        assert torch.allclose(z, fx["enc_z"], atol=2e-6, rtol=0)
        assert torch.equal(codes.to(torch.int16), fx["codes"])                 # integer work: bit-exact
        assert torch.equal(cond, fx["cond"])                                   # lookup + sum: bit-exact
        img = O.cond_upsample(cond, sdm, args.upsampling_ratios)
        assert torch.allclose(img[:, ::4, ::4], fx["img_raw_sub"], atol=1e-6, rtol=0)
        B = wav.shape[0]
        img = img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
        uk = dict(dim=args.diff_dims, upsampling_ratios=tuple(args.upsampling_ratios),
                  unet_scale_cond=args.unet_scale_cond)
        tp = torch.full((B,), fx["t_probe"], dtype=torch.long)
        eps = O.unet_forward(img, tp, cond, sdm, **uk)
        assert torch.allclose(eps[:, ::4, ::4], fx["eps_sub"], atol=1e-5, rtol=0)
        assert abs(eps.double().norm().item() - fx["eps_sum"]["l2"]) < 1e-4 * fx["eps_sum"]["l2"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_synthesize_matches_reference_golden(name):
    """Whole sample.py:94-134 body incl. halfway_sampling with pre-drawn noise and both
    output normalisations; compared with the reference's waveform."""
    fx, args, sdm, sdc, wav = _setup(name)
    L = fx["T"] // int(torch.tensor(args.enc_ratios).prod())
    noise = make_noise_like_reference(fx, L)
    st = {}
    with torch.no_grad():
        x = O.synthesize(wav, sdm, sdc, n_steps=fx["n_steps"], noise=noise, cond_bandwidth=args.cond_bandwidth,
                         enc_ratios=args.enc_ratios, upsampling_ratios=args.upsampling_ratios,
                         diff_dims=args.diff_dims, unet_scale_cond=args.unet_scale_cond, stages=st)
    assert torch.allclose(st["latent"][:, ::4, ::4], fx["latent_sub"], atol=2e-5, rtol=0)
    assert torch.allclose(x, fx["wav_hat"], atol=2e-4, rtol=0)
    assert x.abs().max().item() == pytest.approx(1.0, abs=1e-6)


def make_noise_like_reference(fx, L):
    """make_golden.py seeds the global generator and the reference draws randn_like per step;
    drawing [N-1,B,128,L] in one call from the same seed yields the same stream."""
    torch.manual_seed(fx["seeds"]["noise"])
    return torch.randn(max(fx["n_steps"] - 1, 0), fx["B"], 128, L)


def test_explicit_lstm_equals_aten_lstm():
    torch.manual_seed(3)
    H, T, B = 64, 50, 3
    sd = {}
    for l in range(2):
        for n, shp in (("weight_ih", (4 * H, H)), ("weight_hh", (4 * H, H)), ("bias_ih", (4 * H,)), ("bias_hh", (4 * H,))):
            sd[f"p.lstm.{n}_l{l}"] = torch.randn(shp) * 0.1
    x = torch.randn(B, H, T)
    a = O.slstm(x, sd, "p", 2, fast=False)
    b = O.slstm(x, sd, "p", 2, fast=True)
    assert torch.allclose(a, b, atol=1e-6)


def test_rvq_encode_decode_roundtrip_is_bit_exact():
    g = torch.Generator().manual_seed(5)
    embeds = [torch.randn(1024, 128, generator=g) * 0.3 * 0.7 ** q for q in range(6)]
    x = torch.randn(2, 128, 40, generator=g) * 0.3
    q, codes = O.rvq_forward(x, embeds, 6)
    assert torch.equal(O.rvq_decode(codes, embeds), q)
    # ragged / edge: a single frame, n_q = 1
    q1, c1 = O.rvq_forward(x[:, :, :1], embeds, 1)
    assert c1.shape == (1, 2, 1) and torch.equal(q1, embeds[0][c1[0]].permute(0, 2, 1))


def test_schedule_matches_package_schedule():
    from ladiffcodec_b200.schedule import make_buffers
    a, b = O.schedule_buffers(), make_buffers()
    for k in a:
        assert torch.equal(a[k], b[k]), k
