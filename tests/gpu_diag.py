"""Stage-by-stage GPU diagnostic (development aid; the graded checks are tests/test_parity_gpu.py).

    python tests/gpu_diag.py            # runs every stage in its own subprocess (a CUDA fault cannot poison the rest)
    python tests/gpu_diag.py --stage conv_op

Writes gpurun_out/diag_<stage>.json.
"""
import argparse
import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")

STAGES = ["conv_op", "codec", "unet_simt", "unet_tc", "ddpm", "synth", "synth_B"]   # + "conv_prof" on request


def stage_conv_op(res):
    import ctypes
    import torch
    import torch.nn.functional as F
    from ladiffcodec_b200 import _lib
    lib = _lib.get_lib()
    shapes = [  # (B, L, Cin, Cout, k)
        (2, 75, 128, 128, 1), (2, 80, 256, 256, 3), (1, 300, 1024, 1024, 3), (3, 1200, 256, 384, 1),
        (2, 640, 256, 256, 7), (5, 37, 2048, 1024, 3), (1, 4800, 512, 256, 3), (2, 16, 64, 128, 3),
        (7, 75, 1024, 1024, 3), (32, 75, 256, 128, 3), (3, 150, 512, 512, 3), (2, 1200, 256, 256, 3),
    ]
    P = ctypes.c_void_p
    for (B, L, Cin, Cout, k) in shapes:
        g = torch.Generator().manual_seed(L * 7 + Cin)
        x = torch.randn(B, L, Cin, generator=g).to(torch.bfloat16)
        w = torch.randn(Cout, Cin, k, generator=g) * (Cin * k) ** -0.5
        bias = torch.randn(Cout, generator=g) * 0.1
        ref = F.conv1d(x.float().permute(0, 2, 1), w.to(torch.bfloat16).float(), bias, padding=(k - 1) // 2).permute(0, 2, 1)
        xd, wd, bd = x.cuda(), w.cuda(), bias.cuda()
        out = {}
        for impl in (1, 2, 0):          # SIMT check, tcgen05 per-tap tiles, tcgen05 tap-shared tiles
            for f32 in (1, 0):
                y = torch.full((B, L, Cout), float("nan"), device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
                st = torch.zeros(B, Cout // 32, 2, device="cuda")
                rc = lib.ladiff_op_conv1d_cl(P(xd.data_ptr()), P(wd.data_ptr()), P(bd.data_ptr()), B, L, Cin, Cout, k, P(y.data_ptr()),
                                             f32, impl, P(st.data_ptr()))
                key = f"impl{impl}_{'f32' if f32 else 'bf16'}"
                if rc != 0:
                    out[key] = "rc=%d %s" % (rc, lib.ladiff_last_error().decode())
                    torch.cuda.synchronize()
                    continue
                err = (y.float().cpu() - ref).abs().max().item()
                s_ref = ref.reshape(B, L, Cout // 32, 32).sum(dim=(1, 3))
                out[key] = dict(max_abs_err=round(err, 6), nan=int(torch.isnan(y).sum().item()),
                                stats_err=round((st[:, :, 0].cpu() - s_ref).abs().max().item(), 5))
        res[f"B{B}_L{L}_Cin{Cin}_Cout{Cout}_k{k}"] = dict(out=out, ref_absmax=ref.abs().max().item())
        print(f"conv B{B} L{L} Cin{Cin} Cout{Cout} k{k}: " + " ".join(f"{k2}={v}" for k2, v in out.items()), flush=True)


def stage_conv_prof(res):
    """Config-2-sized conv shapes through the operator entry point (run with LADIFF_TC_PROF=1 for per-role wait cycles)."""
    import ctypes
    import torch
    from ladiffcodec_b200 import _lib
    lib = _lib.get_lib()
    P = ctypes.c_void_p
    shapes = [(32, 1200, 256, 256, 3), (32, 600, 768, 512, 3), (32, 75, 1024, 1024, 3), (32, 1200, 256, 384, 1)]
    if os.environ.get("LADIFF_PROF_ALL"):
        shapes += [(32, 1200, 512, 256, 3), (32, 300, 512, 512, 3), (32, 150, 1024, 1024, 3), (32, 75, 2048, 1024, 3), (32, 75, 1024, 384, 1),
                   (32, 1200, 256, 256, 7)]
    for (B, L, Cin, Cout, k) in shapes:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, L, Cin, generator=g).to(torch.bfloat16).cuda()
        w = (torch.randn(Cout, Cin, k, generator=g) * (Cin * k) ** -0.5).cuda()
        bias = torch.zeros(Cout).cuda()
        y = torch.empty(B, L, Cout, device="cuda", dtype=torch.bfloat16)
        for impl in (0,):
            rc = lib.ladiff_op_conv1d_cl(P(x.data_ptr()), P(w.data_ptr()), P(bias.data_ptr()), B, L, Cin, Cout, k, P(y.data_ptr()), 0, impl, None)
            torch.cuda.synchronize()
            print(f"conv_prof B{B} L{L} Cin{Cin} Cout{Cout} k{k} impl{impl} rc={rc} gflop={2e-9 * B * L * Cin * Cout * k:.2f}", flush=True)
            res[f"L{L}_Cin{Cin}_Cout{Cout}_k{k}_impl{impl}"] = rc


def _setup(name="A_3kbps"):
    import parity_common as pc
    fx, args, sdm, sdc, wav, noise = pc.case_setup(name)
    m, c = pc.cuda_models(args, sdm, sdc)
    return pc, fx, args, sdm, sdc, wav, noise, m, c


def stage_codec(res):
    import torch
    from oracle import ladiff_oracle as O
    for name in ("A_3kbps", "B_3kbps"):
        pc, fx, args, sdm, sdc, wav, noise, m, c = _setup(name)
        with torch.no_grad():
            cond_o, codes_o, z_o = O.get_cond(wav, sdc, args.cond_bandwidth, return_codes=True)
        z = c.encoder(wav.cuda())
        res[name + "_enc_abs"] = (z.cpu() - z_o).abs().max().item()
        res[name + "_enc_scale"] = z_o.abs().max().item()
        print(name, "encoder max abs err", res[name + "_enc_abs"], "scale", res[name + "_enc_scale"], flush=True)
        cond, codes = c.get_cond(wav.cuda(), return_codes=True)
        mism = (codes.cpu() != codes_o).sum().item()
        res[name + "_code_mismatch"] = [mism, codes_o.numel()]
        res[name + "_codes_vs_golden"] = (codes.cpu().to(torch.int16) != fx["codes"]).sum().item()
        res[name + "_cond_abs"] = (cond.cpu() - cond_o).abs().max().item()
        print(name, "codes mismatch", mism, "/", codes_o.numel(), "cond err", res[name + "_cond_abs"], flush=True)
        # RVQ on the oracle's encoder output: isolates the search + lookup
        q2, c2 = c.quantizer._run(z_o.cuda(), codes_o.shape[0], True, True)
        res[name + "_rvq_only_mismatch"] = (c2.cpu() != codes_o).sum().item()
        res[name + "_rvq_only_q_equal"] = bool(torch.equal(q2.cpu(), cond_o))
        res[name + "_rvq_decode_equal"] = bool(torch.equal(c.quantizer.decode(codes_o.cuda()).cpu(), cond_o))
        print(name, "rvq-only mismatch", res[name + "_rvq_only_mismatch"], "q equal", res[name + "_rvq_only_q_equal"],
              "decode equal", res[name + "_rvq_decode_equal"], flush=True)
        with torch.no_grad():
            img_o = O.cond_upsample(cond_o, sdm, args.upsampling_ratios)
        img = cond_o.cuda()
        for layer in m.diff_model.upsampling_layers:
            img = layer(img)
        res[name + "_upsample_abs"] = (img.cpu() - img_o).abs().max().item()
        print(name, "upsample err", res[name + "_upsample_abs"], "scale", img_o.abs().max().item(), flush=True)
        B = wav.shape[0]
        zin = img_o / (img_o.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
        with torch.no_grad():
            d_o = O.seanet_decoder(zin, sdm, list(args.enc_ratios))
        d = m.decoder(zin.cuda())
        res[name + "_dec_abs"] = (d.cpu() - d_o).abs().max().item()
        res[name + "_dec_scale"] = d_o.abs().max().item()
        print(name, "decoder err", res[name + "_dec_abs"], "scale", res[name + "_dec_scale"], flush=True)
        del m, c
        torch.cuda.empty_cache()


def _unet(res, impl, names=("A_3kbps", "B_3kbps")):
    import torch
    from oracle import ladiff_oracle as O
    for name in names:
        pc, fx, args, sdm, sdc, wav, noise, m, c = _setup(name)
        m.set_conv_impl(impl)
        B = wav.shape[0]
        cond_o = fx["cond"]
        with torch.no_grad():
            img = O.cond_upsample(cond_o, sdm, args.upsampling_ratios)
            img = img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
            tp = torch.full((B,), fx["t_probe"], dtype=torch.long)
            eps_o = O.unet_forward(img, tp, cond_o, sdm, **pc.unet_kwargs(args))
        eps = m.diff_model(img.cuda(), tp.cuda(), cond_o.cuda())
        torch.cuda.synchronize()
        r = pc.rel_l2(eps, eps_o)
        res[f"{name}_impl{impl}_rel_l2"] = r
        res[f"{name}_impl{impl}_golden_sub_rel_l2"] = pc.rel_l2(eps.cpu()[:, ::4, ::4], fx["eps_sub"])
        res[f"{name}_impl{impl}_nan"] = int(torch.isnan(eps).sum().item())
        print(name, "impl", impl, "unet rel_l2 vs oracle", r, "nan", res[f"{name}_impl{impl}_nan"], flush=True)
        del m, c
        torch.cuda.empty_cache()


def stage_unet_simt(res):
    _unet(res, 1)


def stage_unet_tc(res):
    _unet(res, 0)


def stage_ddpm(res):
    import torch
    from oracle import ladiff_oracle as O
    for name in ("A_3kbps", "B_3kbps"):
        pc, fx, args, sdm, sdc, wav, noise, m, c = _setup(name)
        B = wav.shape[0]
        cond_o = fx["cond"]
        with torch.no_grad():
            img = O.cond_upsample(cond_o, sdm, args.upsampling_ratios)
            img = img / (img.abs().reshape(B, -1).max(1).values.reshape(B, 1, 1) + 1e-8)
            lat_o = O.halfway_sampling(img.clone(), fx["n_steps"], cond_o, sdm, noise, pc.unet_kwargs(args))
        lat = m.diffusion.halfway_sampling(img=img.cuda(), t=fx["n_steps"], condition=cond_o.cuda(), noise=noise.cuda())
        torch.cuda.synchronize()
        res[name + "_latent_rel_l2"] = pc.rel_l2(lat, lat_o)
        res[name + "_latent_golden_sub_rel_l2"] = pc.rel_l2(lat.cpu()[:, ::4, ::4], fx["latent_sub"])
        print(name, "latent rel_l2 after", fx["n_steps"], "steps:", res[name + "_latent_rel_l2"], flush=True)
        del m, c
        torch.cuda.empty_cache()


def _synth(res, names):
    import torch
    from ladiffcodec_b200.sample import synthesize
    for name in names:
        pc, fx, args, sdm, sdc, wav, noise, m, c = _setup(name)
        out, lat = synthesize(m, c, wav, n_steps=fx["n_steps"], noise=noise, return_latent=True)
        torch.cuda.synchronize()
        res[name + "_wav_snr_db"] = pc.snr_db(out, fx["wav_hat"])
        res[name + "_latent_sub_rel_l2"] = pc.rel_l2(lat.cpu()[:, ::4, ::4], fx["latent_sub"])
        res[name + "_wav_absmax"] = out.abs().max().item()
        print(name, "synthesize: wav SNR vs reference golden", res[name + "_wav_snr_db"], "dB; latent rel_l2",
              res[name + "_latent_sub_rel_l2"], flush=True)
        # Philox mode runs and is deterministic in the seed
        a = synthesize(m, c, wav, n_steps=fx["n_steps"], noise=None, seed=5)
        b = synthesize(m, c, wav, n_steps=fx["n_steps"], noise=None, seed=5)
        res[name + "_philox_deterministic"] = bool(torch.equal(a, b))
        res[name + "_launches"] = m.take_launch_count() + c.take_launch_count()
        del m, c
        torch.cuda.empty_cache()


def stage_synth(res):
    _synth(res, ("A_3kbps", "A_1p5kbps"))


def stage_synth_B(res):
    _synth(res, ("B_3kbps",))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default=None)
    ap.add_argument("--timeout", type=int, default=240)
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if a.stage is None:
        summary = {}
        for s in STAGES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", s], timeout=a.timeout,
                                   capture_output=True, text=True)
                summary[s] = dict(rc=r.returncode, sec=round(time.time() - t0, 1))
                print(f"===== {s}: rc={r.returncode} ({summary[s]['sec']} s)\n{r.stdout[-3000:]}\n{r.stderr[-1500:]}", flush=True)
            except subprocess.TimeoutExpired as e:
                summary[s] = dict(rc="timeout", sec=a.timeout)
                print(f"===== {s}: TIMEOUT\n{(e.stdout or b'')[-2000:]}", flush=True)
        json.dump(summary, open(os.path.join(OUT, "diag_summary.json"), "w"), indent=1)
        return
    res = {}
    try:
        globals()["stage_" + a.stage](res)
        res["ok"] = True
    except Exception:
        res["ok"] = False
        res["traceback"] = traceback.format_exc()
        print(res["traceback"], flush=True)
    json.dump(res, open(os.path.join(OUT, f"diag_{a.stage}.json"), "w"), indent=1, default=str)
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
