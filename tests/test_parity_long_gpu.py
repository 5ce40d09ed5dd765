"""GPU parity at the step counts the benchmarks run (N = 50, 200, 1000) and for the other samplers of GaussianDiffusion1D
(p_sample_loop from noise, ddim_sample, infilling) and the forward-only training loss — the CUDA path through the C-ABI against
(a) the oracle on the same pre-drawn noise and (b) golden vectors produced by the REAL reference (tests/golden/r2_*.pt).

Every measured figure is also appended to gpurun_out/parity_report.json (when that directory exists) — the per-step error-growth
curves committed under profiles/ come from there.  Tolerances: tests/parity_common.py:TOL (stated in DESIGN.md §2)."""
import pytest
import torch

import parity_common as pc
from oracle import ladiff_oracle as O

pytestmark = pytest.mark.gpu
REPORT = pc.MEASURED


def _models(args, sdm, sdc):
    return pc.cuda_models(args, sdm, sdc)


def _wav_from_latent(m, lat):
    x = m.decoder(lat)
    B = x.shape[0]
    m._lib.ladiff_normalize_clips(pc.ptr(x), B, x[0].numel(), 1, pc.stream())
    return x


def _oracle_wav(lat, sdm, args):
    with torch.no_grad():
        y = O.seanet_decoder(lat, sdm, list(args.enc_ratios))
    B = y.shape[0]
    y = y / (y.reshape(B, -1).std(1).reshape(B, 1, 1) + 1e-8)
    return y / (y.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)


@pytest.mark.parametrize("name", ["B_N50", "A_N200"])
def test_long_halfway_trajectory_error_growth(name):
    """halfway_sampling at the benched step counts: rel-L2 of x after EVERY step against the fp32 oracle fed the same noise
    (the error-growth curve), final latent and waveform against the oracle and against the real reference's vectors."""
    fx, args, sdm, sdc, wav, L, d = pc.r2_setup(name)
    m, c = _models(args, sdm, sdc)
    N = fx["case"]["n_steps"]
    cond = c.get_cond(wav.cuda())
    assert torch.equal(cond.cpu(), fx["cond"])
    img = pc.normalized_img(fx["cond"], sdm, args)
    # oracle trajectory (CPU fp32), per step
    otrace = []
    with torch.no_grad():
        x = img.clone()
        k = 0
        for i in reversed(range(N)):
            z = None
            if i > 0:
                z = d["noise"][k]; k += 1
            x, _ = O.p_sample(x, i, fx["cond"], sdm, z, pc.unet_kwargs(args))
            otrace.append(x)
    lat_o = otrace[-1]
    # CUDA trajectory, one step per call so that every intermediate state is visible
    xg = img.cuda()
    noise = d["noise"].cuda()
    gtrace = []
    k = 0
    for i in reversed(range(N)):
        nz = noise[k:k + 1] if i > 0 else torch.zeros(0, *xg.shape, device="cuda")
        k += 1 if i > 0 else 0
        xg, _ = m.diffusion.p_sample(xg, i, cond, noise=nz)
        gtrace.append(xg.cpu())
    # … and the one-call form is the same computation
    lat = m.diffusion.halfway_sampling(img=img.cuda(), t=N, condition=cond, noise=noise).cpu()
    assert torch.equal(lat, gtrace[-1])
    curve = [pc.rel_l2(g, o) for g, o in zip(gtrace, otrace)]
    w = _wav_from_latent(m, lat.cuda()).cpu()
    rep = dict(n_steps=N, latent_rel_l2_vs_oracle=curve[-1], latent_rel_l2_vs_reference=pc.rel_l2(lat, fx["latent"]),
               max_rel_l2_over_steps=max(curve), wav_snr_db_vs_oracle=pc.snr_db(w, _oracle_wav(lat_o, sdm, args)),
               wav_snr_db_vs_reference=pc.snr_db(w, fx["wav_hat"]), rel_l2_per_step=curve)
    REPORT["halfway_" + name] = rep
    print(name, {k: v for k, v in rep.items() if k != "rel_l2_per_step"})
    assert rep["latent_rel_l2_vs_oracle"] <= pc.TOL["latent_rel_l2_long"]
    assert rep["latent_rel_l2_vs_reference"] <= pc.TOL["latent_rel_l2_long"]
    assert rep["wav_snr_db_vs_oracle"] >= pc.TOL["wav_snr_db_long"]
    assert rep["wav_snr_db_vs_reference"] >= pc.TOL["wav_snr_db_long"]
    assert lat.abs().max().item() <= 1.0 + 1e-5
    del m, c
    torch.cuda.empty_cache()


def test_p_sample_loop_first_steps_from_noise():
    """t = 999 … 980 from N(0, I): sqrt_recipm1_alphas_cumprod is 2e4 there, x0 is almost everywhere clamped."""
    fx, args, sdm, sdc, wav, L, d = pc.r2_setup("A_loop20")
    m, c = _models(args, sdm, sdc)
    cond = fx["cond"].cuda()
    lat = m.diffusion.p_sample_loop(d["init"].shape, cond, noise=d["noise"], init=d["init"], n_steps=20).cpu()
    with torch.no_grad():
        lat_o = O.p_sample_loop(d["init"].clone(), fx["cond"], sdm, d["noise"], pc.unet_kwargs(args), n_steps=20)
    rep = dict(latent_rel_l2_vs_oracle=pc.rel_l2(lat, lat_o), latent_rel_l2_vs_reference=pc.rel_l2(lat, fx["latent"]))
    REPORT["loop20"] = rep
    print(rep)
    assert rep["latent_rel_l2_vs_oracle"] <= pc.TOL["latent_rel_l2_long"]
    assert rep["latent_rel_l2_vs_reference"] <= pc.TOL["latent_rel_l2_long"]
    del m, c
    torch.cuda.empty_cache()


def test_sample_full_1000_steps():
    """diffusion.sample() = p_sample_loop, all 1000 steps, against the real reference's final latent, intermediate states and
    waveform (the oracle is pinned to the same chain segment-wise on CPU; 1000 CPU steps do not fit a test)."""
    fx, args, sdm, sdc, wav, L, d = pc.r2_setup("A_full1000")
    m, c = _models(args, sdm, sdc)
    cond = fx["cond"].cuda().contiguous()
    m.diffusion.seq_length = L
    lat = m.diffusion.sample(batch_size=fx["B"], condition=cond, noise=d["noise"], init=d["init"]).cpu()
    # intermediate states: run the chain in the same three pieces the fixture stores
    x20 = m.diffusion.p_sample_loop(d["init"].shape, cond, noise=d["noise"][:20], init=d["init"], n_steps=20).cpu()
    w = _wav_from_latent(m, lat.cuda()).cpu()
    rep = dict(latent_rel_l2_vs_reference=pc.rel_l2(lat, fx["latent"]), state20_rel_l2=pc.rel_l2(x20, fx["states"][20]),
               wav_snr_db_vs_reference=pc.snr_db(w, fx["wav_hat"]))
    # the stored 20-step subsample of the chain: where along the 1000 steps does the CUDA chain leave the reference's?
    x = d["init"].cuda()
    cur, curve = 1000, []
    for j in range(fx["trace_sub"].shape[0]):
        n = fx["trace_every"]
        nz = d["noise"][1000 - cur:1000 - cur + n].cuda()
        x = m.diffusion._steps(x.clone(), cond, cur, n, nz, 0)
        cur -= n
        curve.append(pc.rel_l2(x.cpu()[:, ::8, ::8], fx["trace_sub"][j]))
    lat2 = m.diffusion.sample(batch_size=fx["B"], condition=cond, noise=d["noise"], init=d["init"]).cpu()
    rep["chunked_equals_one_call"] = [bool(torch.equal(x.cpu(), lat)), bool(torch.equal(x.cpu(), lat2)), bool(torch.equal(lat, lat2))]
    rep["rel_l2_every_20_steps"] = curve
    REPORT["sample_full1000"] = rep
    print({k: v for k, v in rep.items() if k != "rel_l2_every_20_steps"}, "max over chain", max(curve))
    assert all(rep["chunked_equals_one_call"]), rep["chunked_equals_one_call"]     # chunked == one call, run to run
    assert rep["latent_rel_l2_vs_reference"] <= pc.TOL["latent_rel_l2_1000"]
    assert rep["wav_snr_db_vs_reference"] >= pc.TOL["wav_snr_db_1000"]
    del m, c
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", ["A_ddim20", "A_ddim10_eta"])
def test_ddim_sample(name):
    fx, args, sdm, sdc, wav, L, d = pc.r2_setup(name)
    m, c = _models(args, sdm, sdc)
    cond = fx["cond"].cuda()
    m.diffusion.sampling_timesteps = fx["case"]["sampling_timesteps"]
    m.diffusion.ddim_sampling_eta = fx["case"]["eta"]
    lat = m.diffusion.ddim_sample(d["init"].shape, cond, noise=d["noise"], init=d["init"]).cpu()
    with torch.no_grad():
        lat_o = O.ddim_sample(d["init"].clone(), fx["cond"], sdm, d["noise"], pc.unet_kwargs(args), fx["case"]["sampling_timesteps"],
                              eta=fx["case"]["eta"])
    w = _wav_from_latent(m, lat.cuda()).cpu()
    rep = dict(latent_rel_l2_vs_oracle=pc.rel_l2(lat, lat_o), latent_rel_l2_vs_reference=pc.rel_l2(lat, fx["latent"]),
               wav_snr_db_vs_reference=pc.snr_db(w, fx["wav_hat"]))
    REPORT["ddim_" + name] = rep
    print(name, rep)
    assert rep["latent_rel_l2_vs_oracle"] <= pc.TOL["latent_rel_l2_ddim"]
    assert rep["latent_rel_l2_vs_reference"] <= pc.TOL["latent_rel_l2_ddim"]
    assert rep["wav_snr_db_vs_reference"] >= pc.TOL["wav_snr_db_ddim"]
    # is_ddim_sampling routes sample() to ddim_sample like the reference (ddpm_loss.py:307-308)
    m.diffusion.is_ddim_sampling, m.diffusion.seq_length = True, L
    lat2 = m.diffusion.sample(batch_size=fx["B"], condition=cond, noise=d["noise"], init=d["init"]).cpu()
    assert torch.equal(lat2, lat)
    # the one-call entry (get_cond -> ddim_sample -> decoder -> normalise) with the same draws gives the same latent
    from ladiffcodec_b200.sample import synthesize
    if fx["case"]["eta"] == 0.0:
        out, lat3 = synthesize(m, c, wav.cuda(), n_steps=fx["case"]["sampling_timesteps"], sampler="ddim", init_noise=d["init"],
                               noise=d["noise"], return_latent=True)
        # (another workspace = another plan; the handle-level tuning cache makes both plans cut the convs identically)
        assert torch.equal(lat3.cpu(), lat) and pc.snr_db(out, w) > 100.0
    del m, c
    torch.cuda.empty_cache()


def test_infilling():
    fx, args, sdm, sdc, wav, L, d = pc.r2_setup("A_infill")
    m, c = _models(args, sdm, sdc)
    cond = fx["cond"].cuda()
    img = pc.normalized_img(fx["cond"], sdm, args)
    m.diffusion.seq_length = L
    lat = m.diffusion.infilling(img.cuda(), cond, midway_t=fx["case"]["midway_t"], lam=fx["case"]["lam"], step_noise=d["noise"],
                                init=d["init"]).cpu()
    with torch.no_grad():
        lat_o = O.infilling(d["init"].clone(), img.clone(), fx["cond"], sdm, d["noise"], pc.unet_kwargs(args), fx["case"]["midway_t"],
                            fx["case"]["lam"])
    rep = dict(latent_rel_l2_vs_oracle=pc.rel_l2(lat, lat_o), latent_rel_l2_vs_reference=pc.rel_l2(lat, fx["latent"]))
    REPORT["infilling"] = rep
    print(rep)
    assert rep["latent_rel_l2_vs_oracle"] <= pc.TOL["latent_rel_l2_long"]
    assert rep["latent_rel_l2_vs_reference"] <= pc.TOL["latent_rel_l2_long"]
    # throughput mode draws: deterministic in the seed
    a = m.diffusion.infilling(img.cuda(), cond, midway_t=3, step_noise=None, seed=5)
    b = m.diffusion.infilling(img.cuda(), cond, midway_t=3, step_noise=None, seed=5)
    assert torch.equal(a, b) and bool(torch.isfinite(a).all())
    del m, c
    torch.cuda.empty_cache()


def test_training_loss_forward():
    """GaussianDiffusion1D.forward → p_losses (ddpm_loss.py:404-450) and DiffAudioRep.forward (model.py:146-221), forward only."""
    fx, args, sdm, sdc, wav, L, d = pc.r2_setup("A_loss")
    m, c = _models(args, sdm, sdc)
    cond = fx["cond"].cuda()
    img = pc.normalized_img(fx["cond"], sdm, args)
    t = torch.tensor(fx["case"]["t"], dtype=torch.long)
    m.diffusion.seq_length = L
    loss, pred, xt, t_out = m.diffusion(img.cuda(), cond, t=t.cuda(), noise=d["noise"])
    with torch.no_grad():
        xt_o = O.q_sample(img, t, d["noise"], sdm)
    assert torch.equal(xt.cpu(), xt_o)                                              # q_sample: bit-exact against the same-box oracle
    assert torch.allclose(xt.cpu(), fx["x_t"], atol=2e-6, rtol=0)                   # … and the reference's, up to the CPU conv's ulps in img
    assert torch.equal(t_out.cpu(), t)
    rep = dict(loss=float(loss), loss_ref=float(fx["loss"]), pred_x_start_rel_l2=pc.rel_l2(pred, fx["pred_x_start"]))
    REPORT["p_losses"] = rep
    print(rep)
    assert abs(rep["loss"] - rep["loss_ref"]) <= pc.TOL["loss_abs"]
    assert rep["pred_x_start_rel_l2"] <= pc.TOL["pred_x0_rel_l2"]
    assert torch.equal(m.diffusion.q_sample(img.cuda(), t.cuda(), d["noise"]).cpu(), xt_o)
    # DiffAudioRep.forward against the oracle's composition of the same stages
    with torch.no_grad():
        x_rep = O.seanet_encoder(wav, sdm, list(args.enc_ratios)) / 18.0
        lo, pred_o, xt_o = O.p_losses(x_rep, t, fx["cond"], d["noise"], sdm, pc.unet_kwargs(args))
        xh_o = O.seanet_decoder(pred_o * 18.0, sdm, list(args.enc_ratios))
        neg_o = O.sd_sdr_neg(wav, xh_o).clamp(min=-30.0).mean()
    losses, x_hat, x_rep_g, pred_g, xt_g, t_g, qtz, scale = m(wav.cuda(), t=t.cuda(), cond=cond, noise=d["noise"])
    assert scale == 18.0 and qtz is None
    assert (x_rep_g.cpu() - x_rep).abs().max().item() <= pc.TOL["codec_abs"]
    assert abs(float(losses["diff_loss"]) - float(lo)) <= pc.TOL["loss_abs"]
    REPORT["p_losses"].update(diff_loss=float(losses["diff_loss"]), diff_loss_oracle=float(lo), neg_loss=float(losses["neg_loss"]),
                              neg_loss_oracle=float(neg_o), x_hat_rel_l2=pc.rel_l2(x_hat, xh_o))
    print(REPORT["p_losses"])
    assert pc.rel_l2(x_hat, xh_o) <= 4 * pc.TOL["pred_x0_rel_l2"]
    assert abs(float(losses["neg_loss"]) - float(neg_o)) <= 0.3                     # dB
    # the codec's own forward (quantising model): eval-mode losses
    lc, xh_c = c(wav.cuda())
    with torch.no_grad():
        q = O.get_cond(wav, sdc, args.cond_bandwidth)
        xh_co = O.seanet_decoder(q, sdc, [8, 5, 4, 2])
    assert (xh_c.cpu() - xh_co).abs().max().item() <= 10 * pc.TOL["codec_abs"]
    assert abs(float(lc["neg_sdr"]) - float(O.sd_sdr_neg(wav, xh_co).clamp(min=-30.0).mean())) <= 1e-3
    del m, c
    torch.cuda.empty_cache()


def test_in_kernel_noise_is_keyed_by_timestep_and_global_clip():
    """Throughput mode: (1) a trajectory split into several calls equals the single call (the counter is the absolute timestep, not
    the call-local step index); (2) a rank that decodes clips [4, 8) of a job with clip_offset = 4 reproduces the job's clips."""
    from ladiffcodec_b200.sample import synthesize
    fx, args, sdm, sdc, wav, noise = pc.case_setup("B_3kbps")
    m, c = _models(args, sdm, sdc)
    wav8 = pc.make_clips(8, fx["T"], seed=77)
    cond = c.get_cond(wav8.cuda())
    img = cond
    for layer in m.diff_model.upsampling_layers:
        img = layer(img)
    one = m.diffusion.halfway_sampling(img=img, t=6, condition=cond, noise=None, seed=3)
    x = img.clone()
    x = m.diffusion._steps(x, cond, 6, 2, None, 3)
    x = m.diffusion._steps(x, cond, 4, 4, None, 3)
    assert torch.equal(x, one)
    steps = [m.diffusion.p_sample(img, t, cond, noise=None, seed=3)[0] for t in (5, 4)]
    assert not torch.equal(steps[0] - img, steps[1] - img)                          # different timesteps draw different noise
    full = synthesize(m, c, wav8.cuda(), n_steps=4, noise=None, seed=9)
    m._lib.ladiff_set_clip_offset(m._h, 4)
    part = synthesize(m, c, wav8[4:].cuda(), n_steps=4, noise=None, seed=9)
    m._lib.ladiff_set_clip_offset(m._h, 0)
    part0 = synthesize(m, c, wav8[4:].cuda(), n_steps=4, noise=None, seed=9)
    assert pc.snr_db(part, full[4:]) > 55.0
    assert pc.snr_db(part0, full[4:]) < 40.0                                        # without the offset the draws differ
    del m, c
    torch.cuda.empty_cache()


def test_full_size_config2_against_oracle():
    """BASELINE config 2 at its full size (B = 32 clips of 2.4 s): the conditioning codec's RVQ codes against the oracle bit for bit,
    and one UNet evaluation on two of the 32 clips (evaluated inside the full batch) against the fp32 oracle."""
    import bench
    cfg = bench.CONFIGS[2]
    args, sdm, sdc = bench.build_state(cfg)
    m, c = _models(args, sdm, sdc)
    B, T = cfg["batch"], bench.T_SAMPLES
    wav = pc.make_clips(B, T, seed=4242)
    cond, codes = c.get_cond(wav.cuda(), return_codes=True)
    with torch.no_grad():
        cond_o, codes_o, z_o = O.get_cond(wav, sdc, args.cond_bandwidth, return_codes=True, fast_lstm=True)
    mism = int((codes.cpu() != codes_o).sum())
    REPORT["config2_full_size"] = dict(code_mismatches=mism, codes=int(codes_o.numel()))
    assert mism == 0, f"{mism} of {codes_o.numel()} RVQ codes differ from the oracle at B=32, T=38400"
    assert torch.equal(cond.cpu(), cond_o)
    img = cond
    for layer in m.diff_model.upsampling_layers:
        img = layer(img)
    m._lib.ladiff_normalize_clips(pc.ptr(img), B, img[0].numel(), 0, pc.stream())
    tt = torch.full((B,), 49, dtype=torch.long, device="cuda")
    eps = m.diff_model(img, tt, cond).cpu()
    sel = [3, 29]
    with torch.no_grad():
        eps_o = O.unet_forward(img.cpu()[sel], tt.cpu()[sel], cond_o[sel], sdm, **pc.unet_kwargs(args))
    r = pc.rel_l2(eps[sel], eps_o)
    REPORT["config2_full_size"]["unet_rel_l2_2_of_32_clips"] = r
    print(REPORT["config2_full_size"])
    assert r <= pc.TOL["unet_rel_l2"]
    del m, c
    torch.cuda.empty_cache()
