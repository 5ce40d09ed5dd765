"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C-ABI, against
 (a) the oracle on the same seeded inputs, (b) golden vectors produced by the REAL reference (tests/golden),
 (c) size-independent properties at BASELINE sizes.
Tolerances are the ones DESIGN.md states: integer/lookup work bit-exact; fp32 SIMT codec stages <= 5e-5 abs at O(1)
scale; UNet (bf16 operands/activations, fp32 accumulate) rel-L2 <= 3e-2 per evaluation; final waveform SNR >= 25 dB."""
import ctypes

import pytest
import torch

import parity_common as pc
from oracle import ladiff_oracle as O

pytestmark = pytest.mark.gpu
CASES = ["A_3kbps", "B_3kbps", "A_1p5kbps"]


@pytest.fixture(scope="module", params=CASES)
def case(request):
    fx, args, sdm, sdc, wav, noise = pc.case_setup(request.param)
    m, c = pc.cuda_models(args, sdm, sdc)
    yield dict(name=request.param, fx=fx, args=args, sdm=sdm, sdc=sdc, wav=wav, noise=noise, m=m, c=c)
    del m, c
    torch.cuda.empty_cache()


def test_library_loaded_is_in_tree():
    from ladiffcodec_b200 import _lib
    _lib.get_lib()
    assert any("libladiff_b200.so" in l for l in open("/proc/self/maps"))


def test_cond_codec_codes_bit_exact(case):
    fx, c, wav = case["fx"], case["c"], case["wav"]
    with torch.no_grad():
        cond_o, codes_o, z_o = O.get_cond(wav, case["sdc"], case["args"].cond_bandwidth, return_codes=True)
    z = c.encoder(wav.cuda()).cpu()
    pc.record("cond_encoder_" + case["name"], max_abs_vs_oracle=(z - z_o).abs().max().item(), max_abs_vs_reference=(z - fx["enc_z"]).abs().max().item(),
              z_absmax=z_o.abs().max().item())
    assert (z - z_o).abs().max().item() <= pc.TOL["codec_abs"]
    assert (z - fx["enc_z"]).abs().max().item() <= pc.TOL["codec_abs"]
    cond, codes = c.get_cond(wav.cuda(), return_codes=True)
    mism = (codes.cpu() != codes_o)
    if mism.any():   # a flip is only admissible at an fp32 near-tie: report the margin
        zf = z_o.permute(0, 2, 1).reshape(-1, 128)
        raise AssertionError(f"{int(mism.sum())} code mismatches of {codes_o.numel()}")
    assert torch.equal(codes.cpu().to(torch.int16), fx["codes"])            # vs the real reference
    assert torch.equal(cond.cpu(), fx["cond"])                              # lookup + residual sum: bit-exact


def test_rvq_lookup_bit_exact_and_roundtrip(case):
    c, fx = case["c"], case["fx"]
    codes = fx["codes"].to(torch.int64)
    q = c.quantizer.decode(codes.cuda()).cpu()
    assert torch.equal(q, fx["cond"])
    embeds = [case["sdc"][f"quantizer.vq.layers.{i}._codebook.embed"] for i in range(codes.shape[0])]
    assert torch.equal(q, O.rvq_decode(codes, embeds))
    # encode(decode(codes)) is idempotent on the quantised representation
    q2, c2 = c.quantizer._run(q.cuda(), codes.shape[0], True, True)
    q3, c3 = c.quantizer._run(q2, codes.shape[0], True, True)
    assert torch.equal(c2, c3) and torch.equal(q2, q3)


def test_cond_upsamplers(case):
    m, fx, args = case["m"], case["fx"], case["args"]
    img = fx["cond"].cuda()
    for layer in m.diff_model.upsampling_layers:
        img = layer(img)
    with torch.no_grad():
        img_o = O.cond_upsample(fx["cond"], case["sdm"], args.upsampling_ratios)
    pc.record("cond_upsample_" + case["name"], max_abs_vs_oracle=(img.cpu() - img_o).abs().max().item(), absmax=img_o.abs().max().item())
    assert (img.cpu() - img_o).abs().max().item() <= pc.TOL["upsample_abs"]
    assert (img.cpu()[:, ::4, ::4] - fx["img_raw_sub"]).abs().max().item() <= pc.TOL["upsample_abs"]


def _unet_inputs(case):
    fx, args = case["fx"], case["args"]
    B = fx["B"]
    with torch.no_grad():
        img = O.cond_upsample(fx["cond"], case["sdm"], args.upsampling_ratios)
    img = img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
    return img, torch.full((B,), fx["t_probe"], dtype=torch.long)


def test_unet_forward_tcgen05(case):
    m, fx, args = case["m"], case["fx"], case["args"]
    img, tp = _unet_inputs(case)
    with torch.no_grad():
        eps_o = O.unet_forward(img, tp, fx["cond"], case["sdm"], **pc.unet_kwargs(args))
    m.set_conv_impl(0)
    eps = m.diff_model(img.cuda(), tp.cuda(), fx["cond"].cuda()).cpu()
    assert torch.isfinite(eps).all()
    pc.record("unet_eval_" + case["name"], rel_l2_vs_oracle=pc.rel_l2(eps, eps_o), rel_l2_vs_reference=pc.rel_l2(eps[:, ::4, ::4], fx["eps_sub"]))
    assert pc.rel_l2(eps, eps_o) <= pc.TOL["unet_rel_l2"]
    assert pc.rel_l2(eps[:, ::4, ::4], fx["eps_sub"]) <= pc.TOL["unet_rel_l2"]        # vs the real reference
    # the tcgen05 kernel and the SIMT check kernel consume identical packed operands
    m.set_conv_impl(1)
    eps_s = m.diff_model(img.cuda(), tp.cuda(), fx["cond"].cuda()).cpu()
    m.set_conv_impl(2)          # tcgen05 with one activation tile per tap (no row-shifted descriptors)
    eps_u = m.diff_model(img.cuda(), tp.cuda(), fx["cond"].cuda()).cpu()
    m.set_conv_impl(0)
    pc.record("unet_eval_" + case["name"], simt_vs_tc=pc.rel_l2(eps, eps_s), per_tap_vs_shared=pc.rel_l2(eps, eps_u))
    assert pc.rel_l2(eps, eps_s) <= pc.TOL["unet_simt_vs_tc_rel_l2"]
    assert pc.rel_l2(eps, eps_u) <= pc.TOL["unet_simt_vs_tc_rel_l2"]
    # per-sample time indices: a batch with mixed t equals the per-t evaluations
    if fx["B"] > 1:
        tmix = tp.clone(); tmix[0] = 3
        e_mix = m.diff_model(img.cuda(), tmix.cuda(), fx["cond"].cuda()).cpu()
        e_3 = m.diff_model(img.cuda(), torch.full_like(tp, 3).cuda(), fx["cond"].cuda()).cpu()
        assert torch.equal(e_mix[0], e_3[0]) and torch.equal(e_mix[1], eps[1])


def test_ddpm_trajectory(case):
    m, fx, args, noise = case["m"], case["fx"], case["args"], case["noise"]
    img, _ = _unet_inputs(case)
    with torch.no_grad():
        lat_o = O.halfway_sampling(img.clone(), fx["n_steps"], fx["cond"], case["sdm"], noise, pc.unet_kwargs(args))
    lat = m.diffusion.halfway_sampling(img=img.cuda(), t=fx["n_steps"], condition=fx["cond"].cuda(), noise=noise.cuda()).cpu()
    pc.record("short_trajectory_" + case["name"], n_steps=fx["n_steps"], latent_rel_l2_vs_oracle=pc.rel_l2(lat, lat_o),
              latent_rel_l2_vs_reference=pc.rel_l2(lat[:, ::4, ::4], fx["latent_sub"]))
    assert pc.rel_l2(lat, lat_o) <= pc.TOL["latent_rel_l2"]
    assert pc.rel_l2(lat[:, ::4, ::4], fx["latent_sub"]) <= pc.TOL["latent_rel_l2"]
    assert lat.abs().max().item() <= 1.0 + 1e-5           # last step is a clamp to [-1,1] scaled by coef1+coef2 = 1
    # single steps through p_sample compose to the same trajectory
    x = img.cuda()
    k = 0
    for i in reversed(range(fx["n_steps"])):
        nz = noise[k:k + 1].cuda() if i > 0 else None
        x, _ = m.diffusion.p_sample(x, i, fx["cond"].cuda(), noise=nz if nz is not None else None)
        k += 1 if i > 0 else 0
    assert torch.equal(x.cpu(), lat)


def test_in_kernel_noise_is_standard_normal_and_seeded(case):
    """Throughput mode (noise=None): the posterior step draws z from the in-kernel counter-based generator.  With the same
    x and eps, x_noisy - x_zero_noise = sigma_t * z, so z is recovered exactly enough to check its first moments, that it is a
    function of the seed only, and that different seeds / steps give different draws."""
    m, fx = case["m"], case["fx"]
    img, _ = _unet_inputs(case)
    cond = fx["cond"].cuda()
    t = 40
    B, C, L = img.shape
    base, _ = m.diffusion.p_sample(img.cuda(), t, cond, noise=torch.zeros(1, B, C, L))
    a, _ = m.diffusion.p_sample(img.cuda(), t, cond, noise=None, seed=7)
    b, _ = m.diffusion.p_sample(img.cuda(), t, cond, noise=None, seed=7)
    c, _ = m.diffusion.p_sample(img.cuda(), t, cond, noise=None, seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c)
    sigma = float(torch.exp(0.5 * case["sdm"]["diffusion.posterior_log_variance_clipped"][t]))
    z = ((a - base) / sigma).double().flatten().cpu()
    n = z.numel()
    assert abs(z.mean().item()) < 6.0 / n ** 0.5
    assert abs(z.var().item() - 1.0) < 6.0 * (2.0 / n) ** 0.5 + 1e-3
    assert abs((z ** 3).mean().item()) < 6.0 * (15.0 / n) ** 0.5 + 1e-3          # skewness of N(0,1) is 0
    assert abs((z ** 4).mean().item() - 3.0) < 6.0 * (96.0 / n) ** 0.5 + 1e-2    # kurtosis 3
    zc = ((c - base) / sigma).double().flatten().cpu()
    assert abs(float((z * zc).mean())) < 6.0 / n ** 0.5                           # seeds are uncorrelated
    zz = z.reshape(B, C, L)
    assert abs(float((zz[:, :, 1:] * zz[:, :, :-1]).mean())) < 6.0 / n ** 0.5     # neighbours along time
    assert abs(float((zz[:, 8:, :] * zz[:, :-8, :]).mean())) < 6.0 / n ** 0.5     # the four elements of one generator call


def test_decoder(case):
    m, args = case["m"], case["args"]
    img, _ = _unet_inputs(case)
    with torch.no_grad():
        d_o = O.seanet_decoder(img, case["sdm"], list(args.enc_ratios))
    d = m.decoder(img.cuda()).cpu()
    pc.record("decoder_" + case["name"], max_abs_vs_oracle=(d - d_o).abs().max().item(), absmax=d_o.abs().max().item())
    assert (d - d_o).abs().max().item() <= pc.TOL["codec_abs"]


def test_synthesize_matches_reference_waveform(case):
    from ladiffcodec_b200.sample import synthesize
    m, c, fx = case["m"], case["c"], case["fx"]
    out, lat = synthesize(m, c, case["wav"].cuda(), n_steps=fx["n_steps"], noise=case["noise"], return_latent=True)
    pc.record("short_trajectory_" + case["name"], wav_snr_db_vs_reference=pc.snr_db(out, fx["wav_hat"]))
    assert pc.snr_db(out, fx["wav_hat"]) >= pc.TOL["wav_snr_db"]
    assert pc.rel_l2(lat.cpu()[:, ::4, ::4], fx["latent_sub"]) <= pc.TOL["latent_rel_l2"]
    assert out.abs().reshape(fx["B"], -1).max(1).values.sub(1.0).abs().max().item() < 1e-5   # sample.py:134
    # host-buffer entry (pinned staging) gives the same samples; batch invariance: clip 0 alone == clip 0 in the batch
    out_h = synthesize(m, c, case["wav"], n_steps=fx["n_steps"], noise=case["noise"])
    assert out_h.device.type == "cpu" and torch.equal(out_h, out.cpu())
    if fx["B"] > 1:
        o0 = synthesize(m, c, case["wav"][:1].cuda(), n_steps=fx["n_steps"], noise=case["noise"][:, :1])
        assert pc.snr_db(o0, out[:1]) > 60.0
    # Philox mode: deterministic in the seed, different across seeds
    a = synthesize(m, c, case["wav"].cuda(), n_steps=fx["n_steps"], noise=None, seed=7)
    b = synthesize(m, c, case["wav"].cuda(), n_steps=fx["n_steps"], noise=None, seed=7)
    d = synthesize(m, c, case["wav"].cuda(), n_steps=fx["n_steps"], noise=None, seed=8)
    assert torch.equal(a, b)
    assert fx["n_steps"] < 2 or not torch.equal(a, d)


def test_codes_in_decode_is_bit_identical(case):
    """decoding from the RVQ indices (the codec's wire payload) gives exactly the samples of the
    waveform-in path, and the indices are the oracle's."""
    from ladiffcodec_b200.sample import synthesize, synthesize_from_codes
    m, c, fx = case["m"], case["c"], case["fx"]
    _, codes = c.get_cond(case["wav"].cuda(), return_codes=True)
    with torch.no_grad():
        codes_o = O.get_cond(case["wav"], case["sdc"], case["args"].cond_bandwidth, return_codes=True)[1]
    assert torch.equal(codes.cpu(), codes_o)
    a, la = synthesize(m, c, case["wav"].cuda(), n_steps=fx["n_steps"], noise=case["noise"], return_latent=True)
    b, lb = synthesize_from_codes(m, c, codes, n_steps=fx["n_steps"], noise=case["noise"], return_latent=True)
    assert torch.equal(a, b) and torch.equal(la, lb)
    with pytest.raises(Exception):
        synthesize_from_codes(m, c, codes[:, :, :-1], n_steps=1)      # odd frame count: not a multiple of 640 samples


def test_reference_script_surface(case):
    """The literal statement sequence of sample.py:94-134 runs on top of the mirrored objects."""
    m, c, fx, wav = case["m"], case["c"], case["fx"], case["wav"]
    if fx["B"] != 1:
        pytest.skip("sample.py is B=1")
    device = m.device
    w = wav.to(device)
    m.diffusion.seq_length = int(w.shape[-1] / case["args"].enc_ratios[0])
    cond = c.get_cond(w)
    img = cond
    for layer in m.diff_model.upsampling_layers:
        img = layer(img)
    img /= torch.max(torch.abs(img.flatten())) + 1e-8
    torch.manual_seed(fx["seeds"]["noise"])
    sample = m.diffusion.halfway_sampling(img=img, condition=cond, t=fx["n_steps"], noise=case["noise"].to(device))
    x = m.decoder(sample)
    x /= torch.std(x.flatten()) + 1e-8
    x /= torch.max(torch.abs(x.flatten())) + 1e-8
    assert pc.snr_db(x, fx["wav_hat"]) >= pc.TOL["wav_snr_db"]


def test_strict_loader_errors():
    from ladiffcodec_b200.model import DiffAudioRep
    from ladiffcodec_b200.layout import cond_model_kwargs
    from ladiffcodec_b200.config import readme_args
    from ladiffcodec_b200.synthetic import make_state_dict
    args = readme_args()
    sd = make_state_dict(seed=1, **cond_model_kwargs(args))
    m = DiffAudioRep(**cond_model_kwargs(args)).to("cuda")
    bad = dict(sd); bad.pop("encoder.model.0.conv.conv.bias")
    with pytest.raises(RuntimeError, match="Missing key"):
        m.load_state_dict(bad, strict=True)
    bad = dict(sd); bad["bogus.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        m.load_state_dict(bad, strict=True)
    bad = dict(sd); bad["encoder.model.0.conv.conv.bias"] = torch.zeros(3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(bad, strict=True)
    ddp = {"module." + k: v for k, v in sd.items()}          # utils.py:100-107
    from ladiffcodec_b200.utils import load_model
    load_model(m, ddp, strict=True)
    with pytest.raises(Exception):
        m.get_cond(torch.zeros(1, 1, 321).cuda())             # not a multiple of the hop


def test_tc_conv_operator_shapes():
    """tcgen05 conv operator vs torch on 16-bit-rounded operands: ragged L (not a tile multiple), k in {1,3,7}, Cin up to 2048,
    minimum size, several short clips packed per tile; fp32 (direct epilogue) and 16-bit (smem-staged TMA-store epilogue)
    outputs; tap-shared (impl 0) and per-tap (impl 2) activation tiles; the positions-on-M kernel with one CTA (impl 3) and a cta_group::2 CTA pair (impl 4) per tile."""
    import torch.nn.functional as F
    from ladiffcodec_b200 import _lib
    lib = _lib.get_lib()
    h16 = _lib.act_dtype()                       # torch.float16 (default build) or torch.bfloat16
    ulp = 2.0 ** -11 if h16 == torch.float16 else 2.0 ** -8
    P = ctypes.c_void_p
    for (B, L, Cin, Cout, k) in [(2, 75, 128, 128, 1), (1, 300, 1024, 1024, 3), (2, 640, 256, 256, 7), (5, 37, 2048, 1024, 3),
                                 (1, 4800, 512, 256, 3), (2, 16, 64, 128, 3), (1, 1201, 256, 384, 1), (7, 75, 1024, 1024, 3),
                                 (3, 150, 512, 512, 3)]:
        g = torch.Generator().manual_seed(L + Cin)
        x = torch.randn(B, L, Cin, generator=g).to(h16)
        w = torch.randn(Cout, Cin, k, generator=g) * (Cin * k) ** -0.5
        bias = torch.randn(Cout, generator=g) * 0.1
        ref = F.conv1d(x.float().permute(0, 2, 1), w.to(h16).float(), bias, padding=(k - 1) // 2).permute(0, 2, 1)
        s_ref = ref.reshape(B, L, Cout // 32, 32).sum(dim=(1, 3))
        xd, wd, bd = x.cuda(), w.cuda(), bias.cuda()
        for impl in (0, 2, 3, 4, 5, 6):
            if impl == 5 and (k != 1 or L <= 120):      # two-CTAs-per-SM shape: small-K single-clip tiles only
                continue

            for f32 in (1, 0):
                y = torch.full((B, L, Cout), float("nan"), device="cuda", dtype=torch.float32 if f32 else h16)
                st = torch.zeros(B, Cout // 32, 2, device="cuda")
                rc = lib.ladiff_op_conv1d_cl(P(xd.data_ptr()), P(wd.data_ptr()), P(bd.data_ptr()), B, L, Cin, Cout, k, P(y.data_ptr()),
                                             f32, impl, P(st.data_ptr()))
                if impl == 6 and rc != 0 and b"pair mode needs" in lib.ladiff_last_error():
                    continue                            # a single position tile: nothing to pair
                assert rc == 0, lib.ladiff_last_error()
                tol = 2e-4 if f32 else 2e-4 + ref.abs().max().item() * ulp          # 16-bit output rounding
                assert (y.float().cpu() - ref).abs().max().item() < tol, (B, L, Cin, Cout, k, impl, f32)
                assert (st[:, :, 0].cpu() - s_ref).abs().max().item() < 1e-2 * max(1.0, s_ref.abs().max().item())


def test_full_size_properties_config2():
    """BASELINE config 2 at its full size (B = 32 clips of 2.4 s, enc_ratios [8,4], L = 1200), through properties that do not
    need the oracle at that size: RVQ round trip bit-exact, tcgen05 vs the SIMT check kernel on one UNet evaluation,
    clip-permutation equivariance and sub-batch invariance of the whole path, output normalisation."""
    import bench
    from ladiffcodec_b200.sample import synthesize
    cfg = bench.CONFIGS[2]
    args, sdm, sdc = bench.build_state(cfg)
    m, c = pc.cuda_models(args, sdm, sdc)
    B, T, n_steps = cfg["batch"], bench.T_SAMPLES, 3
    wav = make_clips_cached(B, T)
    L = T // 32
    g = torch.Generator().manual_seed(99)
    noise = torch.randn(n_steps - 1, B, 128, L, generator=g)
    # RVQ: codes -> decode reproduces the quantized tensor bit for bit at full size
    cond, codes = c.get_cond(wav.cuda(), return_codes=True)
    assert codes.shape == (6, B, T // 320) and int(codes.min()) >= 0 and int(codes.max()) < 1024
    assert torch.equal(c.quantizer.decode(codes), cond)
    # one UNet evaluation: tensor-core path vs SIMT check kernel (same bf16 operands, different accumulation order)
    img = cond
    for layer in m.diff_model.upsampling_layers:
        img = layer(img)
    img = img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
    tt = torch.full((B,), 17, dtype=torch.long, device="cuda")
    e_tc = m.diff_model(img, tt, cond)
    m.set_conv_impl(1)
    e_simt = m.diff_model(img, tt, cond)
    m.set_conv_impl(0)
    pc.record("config2_full_size", simt_vs_tc=pc.rel_l2(e_tc, e_simt))
    assert pc.rel_l2(e_tc, e_simt) <= pc.TOL["unet_simt_vs_tc_rel_l2"]     # 16-bit activations: rounding flips, not a bias
    # whole path: permuting the clips permutes the outputs; a sub-batch reproduces its clips
    out = synthesize(m, c, wav.cuda(), n_steps=n_steps, noise=noise)
    assert out.shape == (B, 1, T) and bool(torch.isfinite(out).all())
    assert out.abs().reshape(B, -1).max(1).values.sub(1.0).abs().max().item() < 1e-5
    perm = torch.randperm(B, generator=g)
    out_p = synthesize(m, c, wav[perm].cuda(), n_steps=n_steps, noise=noise[:, perm])
    assert pc.snr_db(out_p, out[perm]) > 55.0
    sub = synthesize(m, c, wav[:4].cuda(), n_steps=n_steps, noise=noise[:, :4])
    assert pc.snr_db(sub, out[:4]) > 55.0
    del m, c
    torch.cuda.empty_cache()


def make_clips_cached(B, T):
    from ladiffcodec_b200.synthetic import make_clips
    return make_clips(B, T, seed=4242)


def test_long_utterance_against_oracle():
    """a whole utterance (5.12 s, Layout-A: latent L = 10240, bottleneck attention over n = 640 positions,
    encoder LSTM over 256 frames, decoder LSTM over 10240) goes through the same entry point; checked against the oracle."""
    from ladiffcodec_b200.config import readme_args
    from ladiffcodec_b200.sample import synthesize
    args = readme_args()
    sdm = pc.make_state_dict(seed=11, **pc.ladiff_model_kwargs(args))
    sdc = pc.make_state_dict(seed=12, **pc.cond_model_kwargs(args))
    m, c = pc.cuda_models(args, sdm, sdc)
    T, n_steps = 640 * 128, 2
    wav = pc.make_clips(1, T, seed=5)
    L = T // 8
    noise = torch.randn(n_steps - 1, 1, 128, L, generator=torch.Generator().manual_seed(3))
    out, lat = synthesize(m, c, wav.cuda(), n_steps=n_steps, noise=noise, return_latent=True)
    stages = {}
    with torch.no_grad():
        ref = O.synthesize(wav, sdm, sdc, n_steps=n_steps, noise=noise, cond_bandwidth=args.cond_bandwidth,
                           enc_ratios=args.enc_ratios, upsampling_ratios=args.upsampling_ratios, diff_dims=args.diff_dims,
                           unet_scale_cond=args.unet_scale_cond, fast_lstm=True, stages=stages)
    pc.record("long_utterance_5s", latent_rel_l2=pc.rel_l2(lat, stages["latent"]), wav_snr_db=pc.snr_db(out, ref))
    assert pc.rel_l2(lat, stages["latent"]) <= pc.TOL["latent_rel_l2"]
    assert pc.snr_db(out, ref) >= pc.TOL["wav_snr_db"]
    del m, c
    torch.cuda.empty_cache()


@pytest.mark.parametrize("B,T,layout", [(1, 640, "A"), (3, 1280, "A"), (5, 1920, "B"), (2, 640 * 7, "B")])
def test_edge_shapes_against_oracle(B, T, layout):
    """Smallest legal clip (one 640-sample block: 2 conditioning frames, latent L = 80 / 20 → bottleneck of 5 / 1.25 → L must
    be a multiple of 16 for the UNet: Layout-B needs T % 512 == 0, so those cases use multiples that satisfy both), odd batch
    sizes, lengths that are not multiples of any tile size — all through the one-call entry, against the oracle."""
    from ladiffcodec_b200.config import readme_args, sample_args
    from ladiffcodec_b200.sample import synthesize
    if layout == "A":
        args = readme_args()
        hop = 8
    else:
        args = sample_args(run_diff=True, cond_bandwidth=3.0, enc_ratios=[8, 4], upsampling_ratios=[5, 2], diff_dims=256,
                           model_for_cond="c", model_path="m")
        hop = 32
    if (T // hop) % 16 != 0:
        T = (T // (hop * 16) + 1) * hop * 16
        T = (T + 639) // 640 * 640
        while (T // hop) % 16 != 0 or T % 640 != 0:
            T += 640
    sdm = pc.make_state_dict(seed=21, **pc.ladiff_model_kwargs(args))
    sdc = pc.make_state_dict(seed=22, **pc.cond_model_kwargs(args))
    m, c = pc.cuda_models(args, sdm, sdc)
    n_steps = 2
    wav = pc.make_clips(B, T, seed=B * 7 + T)
    L = T // hop
    noise = torch.randn(n_steps - 1, B, 128, L, generator=torch.Generator().manual_seed(B + T))
    out, lat = synthesize(m, c, wav.cuda(), n_steps=n_steps, noise=noise, return_latent=True)
    stages = {}
    with torch.no_grad():
        ref = O.synthesize(wav, sdm, sdc, n_steps=n_steps, noise=noise, cond_bandwidth=args.cond_bandwidth,
                           enc_ratios=args.enc_ratios, upsampling_ratios=args.upsampling_ratios, diff_dims=args.diff_dims,
                           unet_scale_cond=args.unet_scale_cond, fast_lstm=True, stages=stages)
    assert out.shape == (B, 1, T)
    pc.record(f"edge_B{B}_T{T}_{layout}", latent_rel_l2=pc.rel_l2(lat, stages["latent"]), wav_snr_db=pc.snr_db(out, ref))
    assert pc.rel_l2(lat, stages["latent"]) <= pc.TOL["latent_rel_l2"]
    assert pc.snr_db(out, ref) >= pc.TOL["wav_snr_db"]
    del m, c
    torch.cuda.empty_cache()


def test_pipeline_two_batches_in_flight(case):
    """SynthesisPipeline (depth 2: two batches on two streams, own workspaces) returns, ticket by ticket, what the one-call
    entry returns for the same clips and seeds — device and host inputs, more submissions than slots."""
    from ladiffcodec_b200.sample import synthesize, SynthesisPipeline
    m, c, fx = case["m"], case["c"], case["fx"]
    wavs = [case["wav"], case["wav"].flip(0) * 0.5, case["wav"] * 0.25, case["wav"].roll(7, -1)]
    ref = [synthesize(m, c, w.cuda(), n_steps=fx["n_steps"], noise=None, seed=11 + i).cpu() for i, w in enumerate(wavs)]
    pipe = SynthesisPipeline(m, c, depth=2)
    tickets = [pipe.submit(w.cuda() if i % 2 == 0 else w.pin_memory(), n_steps=fx["n_steps"], seed=11 + i) for i, w in enumerate(wavs)]
    outs = [pipe.result(t) for t in tickets]
    torch.cuda.synchronize()
    for i, (o, r) in enumerate(zip(outs, ref)):
        assert (o.device.type == "cuda") == (i % 2 == 0)
        assert pc.snr_db(o, r) > 55.0, i


@pytest.mark.parametrize("L", [240, 480, 960, 1920])
def test_unet_lengths_around_the_multi_clip_tile_boundary(L):
    """Latent lengths whose resolution levels straddle the 120-row limit of the several-clips-per-tile mode (a level with exactly
    120 rows used to make the tap-shared and the per-tap tilings of one conv disagree): one UNet evaluation against the oracle, and
    the SIMT / per-tap check variants on the same plan."""
    from ladiffcodec_b200.config import readme_args
    args = readme_args()
    sdm = pc.make_state_dict(seed=11, **pc.ladiff_model_kwargs(args))
    sdc = pc.make_state_dict(seed=12, **pc.cond_model_kwargs(args))
    m, c = pc.cuda_models(args, sdm, sdc)
    F = L // 40                      # upsampling ratios 5, 4, 2; level j of the UNet has L / 2^j rows -> one of them has exactly 120
    g = torch.Generator().manual_seed(L)
    cond = torch.randn(2, 128, F, generator=g) * 0.3
    x = torch.randn(2, 128, L, generator=g)
    t = torch.tensor([5, 700])
    with torch.no_grad():
        eo = O.unet_forward(x, t, cond, sdm, **pc.unet_kwargs(args))
    e = m.diff_model(x.cuda(), t.cuda(), cond.cuda()).cpu()
    assert pc.rel_l2(e, eo) <= pc.TOL["unet_rel_l2"]
    for impl in (1, 2):
        m.set_conv_impl(impl)
        assert pc.rel_l2(m.diff_model(x.cuda(), t.cuda(), cond.cuda()).cpu(), e) <= pc.TOL["unet_simt_vs_tc_rel_l2"]
    m.set_conv_impl(0)
    del m, c
    torch.cuda.empty_cache()


@pytest.mark.parametrize("B,L", [(2, 75), (1, 300), (3, 129), (1, 1000), (1, 4375), (2, 17)])
def test_full_attention_operator_tcgen05_and_simt(B, L):
    """Mid-block attention core (unet.py:238-245) on the tensor cores (tcgen05 QK^T and PV) and on the two SIMT kernels against torch
    on the same 16-bit-rounded q, k, v: ragged lengths, several clips, the 35 s utterance's bottleneck (n = 4375)."""
    from ladiffcodec_b200 import _lib
    lib = _lib.get_lib()
    h16 = _lib.act_dtype()
    g = torch.Generator().manual_seed(B * 1000 + L)
    qkv = (torch.randn(B, L, 384, generator=g) * 1.5).to(h16)
    q, k, v = [t.float().reshape(B, L, 4, 32).permute(0, 2, 1, 3) for t in qkv.split(128, dim=2)]      # [B, h, L, d]
    sim = torch.einsum("bhid,bhjd->bhij", q * 32 ** -0.5, k)
    ref = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v).permute(0, 2, 1, 3).reshape(B, L, 128)
    qd = qkv.cuda()
    P = ctypes.c_void_p
    for impl in (3, 1, 2, 0):
        if impl == 2 and L > 512:
            continue
        out = torch.full((B, L, 128), float("nan"), dtype=h16, device="cuda")
        rc = lib.ladiff_op_fullattn(P(qd.data_ptr()), P(out.data_ptr()), B, L, impl)
        assert rc == 0, lib.ladiff_last_error()
        err = (out.float().cpu() - ref).abs().max().item()
        tol = (4e-3 if h16 == torch.float16 else 3e-2) * max(1.0, ref.abs().max().item())
        assert err < tol, (B, L, impl, err)


def test_35s_utterance_against_oracle():
    """SURVEY §8f rank 1 at the length the reference's script meets on real LibriSpeech files: a 35 s utterance (T = 560 000 samples,
    Layout-A latent L = 70 000, bottleneck attention over n = 4375 positions -> the tcgen05 attention kernel; encoder LSTM over 1750
    frames, decoder LSTM over 70 000 steps) through the one-call entry, against the oracle."""
    from ladiffcodec_b200.config import readme_args
    from ladiffcodec_b200.sample import synthesize
    args = readme_args()
    sdm = pc.make_state_dict(seed=11, **pc.ladiff_model_kwargs(args))
    sdc = pc.make_state_dict(seed=12, **pc.cond_model_kwargs(args))
    m, c = pc.cuda_models(args, sdm, sdc)
    T, n_steps = 640 * 875, 2
    wav = pc.make_clips(1, T, seed=35)
    L = T // 8
    noise = torch.randn(n_steps - 1, 1, 128, L, generator=torch.Generator().manual_seed(4))
    out, lat = synthesize(m, c, wav.cuda(), n_steps=n_steps, noise=noise, return_latent=True)
    ws_gb = m._lib.ladiff_synthesize_workspace_bytes(m._h, c._h, 1, T) / 2 ** 30
    stages = {}
    with torch.no_grad():
        ref = O.synthesize(wav, sdm, sdc, n_steps=n_steps, noise=noise, cond_bandwidth=args.cond_bandwidth,
                           enc_ratios=args.enc_ratios, upsampling_ratios=args.upsampling_ratios, diff_dims=args.diff_dims,
                           unet_scale_cond=args.unet_scale_cond, fast_lstm=True, stages=stages)
    pc.record("long_utterance_35s", latent_rel_l2=pc.rel_l2(lat, stages["latent"]), wav_snr_db=pc.snr_db(out, ref), workspace_gib=ws_gb,
              bottleneck_n=L // 16)
    assert out.shape == (1, 1, T)
    assert pc.rel_l2(lat, stages["latent"]) <= pc.TOL["latent_rel_l2"]
    assert pc.snr_db(out, ref) >= pc.TOL["wav_snr_db"]
    del m, c
    torch.cuda.empty_cache()


def test_fused_linear_attention_tail_operator_level():
    """The opt-in fused tail of the linear-attention block (ctx^T softmax(q) -> to_out 1x1 on tcgen05 -> LayerNorm -> + x in one kernel,
    LADIFF_ATTN_TAIL=1) against the default three-kernel form: run in a subprocess so that the environment switch is seen at plan time."""
    import os, subprocess, sys
    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import parity_common as pc\n"
        "fx, args, sdm, sdc, wav, noise = pc.case_setup('B_3kbps')\n"
        "m, c = pc.cuda_models(args, sdm, sdc)\n"
        "from oracle import ladiff_oracle as O\n"
        "B = fx['B']\n"
        "img = pc.normalized_img(fx['cond'], sdm, args)\n"
        "tp = torch.full((B,), fx['t_probe'], dtype=torch.long)\n"
        "with torch.no_grad(): eo = O.unet_forward(img, tp, fx['cond'], sdm, **pc.unet_kwargs(args))\n"
        "e = m.diff_model(img.cuda(), tp.cuda(), fx['cond'].cuda()).cpu()\n"
        "print('REL', pc.rel_l2(e, eo))\n"
    ) % (pc.ROOT, os.path.join(pc.ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, LADIFF_ATTN_TAIL="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    rel = float([l for l in r.stdout.splitlines() if l.startswith("REL")][0].split()[1])
    assert rel <= pc.TOL["unet_rel_l2"], rel


@pytest.mark.gpu
def test_codec_fma_kernels_against_tensor_core_path(case):
    """The codec's two conv implementations against each other and the oracle: the default tcgen05 3xTF32 kernel (codec_tc.cu) in this
    process, the fp32 FMA kernels (LADIFF_CODEC_SIMT=1, read once per process) in a subprocess.  Both must give the reference's RVQ
    codes; the encoder / decoder outputs of the two must agree within the codec tolerance."""
    import os, subprocess, sys, tempfile
    if case["name"] != "B_3kbps":
        pytest.skip("one layout is enough for the A/B of the two implementations")
    c, m, wav = case["c"], case["m"], case["wav"]
    img, _ = _unet_inputs(case)
    z_tc = c.encoder(wav.cuda()).cpu()
    _, codes_tc = c.get_cond(wav.cuda(), return_codes=True)
    d_tc = m.decoder(img.cuda()).cpu()
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "fma.pt")
        code = (
            "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import parity_common as pc\n"
            "fx, args, sdm, sdc, wav, noise = pc.case_setup('B_3kbps')\n"
            "m, c = pc.cuda_models(args, sdm, sdc)\n"
            "img = pc.normalized_img(fx['cond'], sdm, args)\n"
            "z = c.encoder(wav.cuda()).cpu()\n"
            "_, codes = c.get_cond(wav.cuda(), return_codes=True)\n"
            "d = m.decoder(img.cuda()).cpu()\n"
            "torch.save(dict(z=z, codes=codes.cpu(), d=d), %r)\n"
        ) % (pc.ROOT, os.path.join(pc.ROOT, "tests"), out)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, LADIFF_CODEC_SIMT="1"))
        assert r.returncode == 0, r.stderr[-2000:]
        fma = torch.load(out)
    pc.record("codec_tensor_core_vs_fma", encoder_max_abs=(z_tc - fma["z"]).abs().max().item(), decoder_max_abs=(d_tc - fma["d"]).abs().max().item())
    assert torch.equal(codes_tc.cpu(), fma["codes"])
    assert torch.equal(codes_tc.cpu().to(torch.int16), case["fx"]["codes"])
    assert (z_tc - fma["z"]).abs().max().item() <= pc.TOL["codec_abs"]
    assert (d_tc - fma["d"]).abs().max().item() <= pc.TOL["codec_abs"]


@pytest.mark.gpu
def test_linear_attention_tcgen05_against_fma_kernel():
    """Second half of the linear attention: the default tcgen05 kernel with split operands (linattn_out_tc_kernel) against the fp32 FMA
    kernel (LADIFF_LINATTN_OUT_SIMT=1, read at launch time once per process -> subprocess) through a whole UNet evaluation."""
    import os, subprocess, sys, tempfile
    fx, args, sdm, sdc, wav, noise = pc.case_setup("B_3kbps")
    m, c = pc.cuda_models(args, sdm, sdc)
    B = fx["B"]
    img = pc.normalized_img(fx["cond"], sdm, args)
    tp = torch.full((B,), fx["t_probe"], dtype=torch.long)
    e_tc = m.diff_model(img.cuda(), tp.cuda(), fx["cond"].cuda().contiguous()).cpu()
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "simt.pt")
        code = (
            "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import parity_common as pc\n"
            "fx, args, sdm, sdc, wav, noise = pc.case_setup('B_3kbps')\n"
            "m, c = pc.cuda_models(args, sdm, sdc)\n"
            "img = pc.normalized_img(fx['cond'], sdm, args)\n"
            "tp = torch.full((fx['B'],), fx['t_probe'], dtype=torch.long)\n"
            "e = m.diff_model(img.cuda(), tp.cuda(), fx['cond'].cuda().contiguous()).cpu()\n"
            "torch.save(e, %r)\n"
        ) % (pc.ROOT, os.path.join(pc.ROOT, "tests"), out)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, LADIFF_LINATTN_OUT_SIMT="1"))
        assert r.returncode == 0, r.stderr[-2000:]
        e_simt = torch.load(out)
    rel = pc.rel_l2(e_tc, e_simt)
    pc.record("linattn_tcgen05_vs_fma", unet_eval_rel_l2=rel)
    # both kernels are fp32-accurate before the 16-bit store, but ANY two evaluation orders of this UNet differ by ~1.4e-3 (16-bit
    # rounding flips propagate through 80 convs: the tcgen05-vs-check-kernel and per-tap-vs-shared figures are the same size)
    assert rel <= pc.TOL["unet_rel_l2"], rel
