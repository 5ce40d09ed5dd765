#!/usr/bin/env python
"""bench.py — audio-seconds decoded per wall-second on the LaDiffCodec sampling path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5] [--batch B]
                    [--ddpm_steps N] [--sampler ddpm|ddim] [--depth D]

One "step" = one pass of the hot path (sample.py:94-134: get_cond → upsample → N-step DDPM → decoder → normalise) over one
batch of synthetic 2.4 s / 16 kHz clips.  Default workload = BASELINE config 2: B=32 per GPU, 3 kbps, enc_ratios [8,4]
(+ upsampling [5,2], latent L=1200), diff_dims 256, 50 DDPM steps; weak scaling (every rank decodes its own B clips, no
data-path collective).  --config 4 is BASELINE's strong-scaling case (1024 clips in all, scattered from / gathered to rank 0
over NCCL inside the timed region); --config 5 sweeps the step count at B=128 (one JSON line per N); --config 1/3 are the
other BASELINE configurations.  N > 1: one process per GPU; `python bench.py --gpus N` re-executes itself under torchrun when it
was not launched by it.  Prints ONE JSON line per measurement on rank 0.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_SECONDS = 2.4
T_SAMPLES = 38400
LAYOUT_A = dict(run_diff=True, scaling_global=True, unet_scale_cond=True, diff_dims=256, model_for_cond="c", model_path="m")
LAYOUT_B = dict(run_diff=True, enc_ratios=[8, 4], upsampling_ratios=[5, 2], diff_dims=256, model_for_cond="c", model_path="m")

CONFIGS = {
    1: dict(name="config1: single 2.4 s clip, 3 kbps, README layout (enc_ratios [8], L=4800), 50-step DDPM",
            flags=dict(LAYOUT_A, cond_bandwidth=3.0), batch=1, n_steps=50, scaling="weak"),
    2: dict(name="config2: B=32/GPU, 3 kbps, enc_ratios [8,4], upsampling [5,2] (L=1200), diff_dims 256, 50-step DDPM",
            flags=dict(LAYOUT_B, cond_bandwidth=3.0), batch=32, n_steps=50, scaling="weak"),
    3: dict(name="config3: B=256, 1.5 kbps, enc_ratios [8] (L=4800), scaling_global + unet_scale_cond, 200-step DDPM",
            flags=dict(LAYOUT_A, cond_bandwidth=1.5), batch=256, n_steps=200, scaling="weak"),
    4: dict(name="config4: 1024 clips in all, 3 kbps, README layout (L=4800), 50-step DDPM, sharded over the GPUs of the box; "
                 "NCCL scatter from / gather to rank 0 inside the timed region",
            flags=dict(LAYOUT_A, cond_bandwidth=3.0), batch=1024, n_steps=50, scaling="strong", sub_batch=128),
    5: dict(name="config5: B=128, 3 kbps, README layout (L=4800), DDPM step sweep",
            flags=dict(LAYOUT_A, cond_bandwidth=3.0), batch=128, n_steps=50, scaling="weak", sweep=[10, 50, 200, 1000]),
}
PARITY_REPORT = os.path.join(ROOT, "profiles", "r2d", "parity_report_f16.json")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.path = gpu_index, None, f"/tmp/ladiff_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def build_state(cfg, seed_m=101, seed_c=102):
    from ladiffcodec_b200.config import sample_args
    from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs
    from ladiffcodec_b200.synthetic import make_state_dict
    args = sample_args(**cfg["flags"])
    sdm = make_state_dict(seed=seed_m, **ladiff_model_kwargs(args))
    sdc = make_state_dict(seed=seed_c, **cond_model_kwargs(args))
    return args, sdm, sdc


def time_reference_sample(args, sdm, sdc, n_steps, clips, ddpm_steps_timed, threads):
    """Times the oracle port of the reference's CPU path (the same ATen kernels the reference dispatches to) on `clips` clips
    (min(B, 8) clips, k >= 3 DDPM steps): the complete codec stages + `ddpm_steps_timed` DDPM steps (after one
    untimed step), extrapolated linearly to n_steps (every step is identical work).  Returns (audio_s_per_s, detail)."""
    import torch
    from oracle import ladiff_oracle as O
    from ladiffcodec_b200.synthetic import make_clips
    torch.set_num_threads(threads)
    wav = make_clips(clips, T_SAMPLES, seed=4321)
    uk = dict(dim=args.diff_dims, upsampling_ratios=tuple(args.upsampling_ratios), unet_scale_cond=args.unet_scale_cond)
    with torch.no_grad():
        t0 = time.perf_counter()
        cond = O.get_cond(wav, sdc, args.cond_bandwidth, fast_lstm=True)
        img = O.cond_upsample(cond, sdm, args.upsampling_ratios)
        img = img / (img.abs().reshape(clips, -1).max(1).values.reshape(clips, 1, 1) + 1e-8)
        t_front = time.perf_counter() - t0
        x = img
        x, _ = O.p_sample(x, n_steps - 1, cond, sdm, torch.randn_like(x), uk)          # warm-up step (untimed)
        t0 = time.perf_counter()
        for i in range(ddpm_steps_timed):
            x, _ = O.p_sample(x, n_steps - 2 - i, cond, sdm, torch.randn_like(x), uk)
        t_steps = time.perf_counter() - t0
        t_step = t_steps / ddpm_steps_timed
        t0 = time.perf_counter()
        y = O.seanet_decoder(x, sdm, list(args.enc_ratios), fast_lstm=True)
        y = y / (y.reshape(clips, -1).std(1).reshape(clips, 1, 1) + 1e-8)
        y = y / (y.abs().reshape(clips, -1).max(1).values.reshape(clips, 1, 1) + 1e-8)
        t_back = time.perf_counter() - t0
    total = t_front + n_steps * t_step + t_back
    detail = dict(clips=clips, t_codec_front_s=round(t_front, 3), t_ddpm_step_s=round(t_step, 3), t_decoder_s=round(t_back, 3),
                  ddpm_steps_timed=ddpm_steps_timed, measured_s=round(t_front + t_steps + t_back, 2), extrapolated_total_s=round(total, 2))
    return clips * CLIP_SECONDS / total, detail


def cpu_sample_desc(clips, timed, n_steps):
    return (f"{clips} clips: full get_cond + upsample + decoder, {timed} timed DDPM steps (1 untimed), extrapolated linearly to {n_steps} "
            "steps; oracle port of the reference (the ATen conv/LSTM kernels the reference dispatches to, minus its per-step weight "
            "standardisation and process_cond recomputation), fp32")


def run_reference(a, cfg):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; /root/reference cannot travel to the GPU
    box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args, sdm, sdc = build_state(cfg)
    threads = os.cpu_count() or 1
    clips, timed = min(cfg["batch"], 8), 3
    vals, last = [], None
    for i in range(a.warmup + a.steps):
        v, d = time_reference_sample(args, sdm, sdc, cfg["n_steps"], clips, timed, threads)
        if i >= a.warmup:
            vals.append(v); last = d
    value = statistics.mean(vals)
    sample = cpu_sample_desc(clips, timed, cfg["n_steps"])
    line = dict(metric="audio-sec/s decoded", value=value, unit="audio-s/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=1000.0 * clips * CLIP_SECONDS / value, higher_is_better=True, scaling=cfg["scaling"], vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=cfg["name"], n_ddpm_steps=cfg["n_steps"], clip_seconds=CLIP_SECONDS),
                cpu_baseline=dict(value=value, unit="audio-s/s", cores=threads, kind="port", sample=sample, detail=last),
                e2e=dict(value=value, unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def parity_block(dtype):
    """What the GPU parity tests measured for this arithmetic (committed report; the bench itself never runs the oracle on the
    measured path)."""
    blk = dict(dtype=dtype, source=None)
    if dtype == "f16" and os.path.exists(PARITY_REPORT):
        r = json.load(open(PARITY_REPORT))
        g = lambda k, f: r.get(k, {}).get(f)
        blk.update(source="profiles/r2d/parity_report_f16.json (tests/test_parity_gpu.py + test_parity_long_gpu.py on one B200, final build, CUDA path vs the fp32 oracle and "
                          "vs vectors of the real reference, same pre-drawn noise)",
                   unet_eval_rel_l2_config2_full_size=g("config2_full_size", "unet_rel_l2_2_of_32_clips"),
                   rvq_code_mismatches_config2_full_size=g("config2_full_size", "code_mismatches"),
                   codec_max_abs=dict(cond_encoder=g("cond_encoder_B_3kbps", "max_abs_vs_reference"), cond_upsample=g("cond_upsample_B_3kbps", "max_abs_vs_oracle"),
                                      decoder=g("decoder_B_3kbps", "max_abs_vs_oracle")),
                   n50_layout_b=dict(latent_rel_l2=g("halfway_B_N50", "latent_rel_l2_vs_reference"), wav_snr_db=g("halfway_B_N50", "wav_snr_db_vs_reference")),
                   n200_layout_a=dict(latent_rel_l2=g("halfway_A_N200", "latent_rel_l2_vs_reference"), wav_snr_db=g("halfway_A_N200", "wav_snr_db_vs_reference")),
                   n1000_from_noise=dict(latent_rel_l2=g("sample_full1000", "latent_rel_l2_vs_reference"), wav_snr_db=g("sample_full1000", "wav_snr_db_vs_reference")),
                   ddim20=dict(latent_rel_l2=g("ddim_A_ddim20", "latent_rel_l2_vs_reference"), wav_snr_db=g("ddim_A_ddim20", "wav_snr_db_vs_reference")),
                   stated_tolerance="tests/parity_common.py:_TOL_F16 (measured x ~2): UNet evaluation rel-L2 <= 4e-3; N=50/200 latent rel-L2 <= 2e-3, "
                                    "waveform SNR >= 54 dB; RVQ codes and lookup bit-exact; fp32 codec stages <= 5e-5 abs")
    return blk


def run_ours(a, cfg):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    if a.gpus != world:
        raise SystemExit(f"bench.py: --gpus {a.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {a.gpus} "
                         "(plain `python bench.py --gpus N` does that by itself)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ladiffcodec_b200 import _lib, profiling
    from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs
    from ladiffcodec_b200.model import DiffAudioRep
    from ladiffcodec_b200.sample import synthesize, SynthesisPipeline
    from ladiffcodec_b200.shard import scatter_clips, gather_clips, clip_range
    from ladiffcodec_b200.synthetic import make_clips
    from ladiffcodec_b200.utils import load_model

    dtype = _lib.get_lib().ladiff_act_dtype().decode()
    args, sdm, sdc = build_state(cfg)
    model = DiffAudioRep(**ladiff_model_kwargs(args)).to("cuda")
    load_model(model, sdm, strict=True)
    cmodel = DiffAudioRep(**cond_model_kwargs(args)).to("cuda")
    load_model(cmodel, sdc)
    strong = cfg["scaling"] == "strong"
    N = cfg["n_steps"]
    if strong:
        total_clips = a.batch or cfg["batch"]
        lo, hi = clip_range(total_clips, world, rank)
        B = hi - lo                                          # this rank's share
        sub = min(cfg.get("sub_batch", B), B)
        model._lib.ladiff_set_clip_offset(model._h, lo)      # in-kernel noise is keyed by the GLOBAL clip index
    else:
        B = a.batch or cfg["batch"]
        total_clips = world * B
        sub = B
        model._lib.ladiff_set_clip_offset(model._h, rank * B)
    n_sets = min(a.steps, 2 if total_clips > 512 else 4)
    # distinct clips per rank and per step so nothing can be reused between timed iterations
    if strong:
        job_dev = [make_clips(total_clips, T_SAMPLES, seed=9000 + 37 * s).cuda() for s in range(n_sets)] if rank == 0 else [None] * n_sets
        host_sets = dev_sets = None
    else:
        host_sets = [make_clips(B, T_SAMPLES, seed=9000 + 1000 * rank + 37 * s).pin_memory() for s in range(n_sets)]
        dev_sets = [h.cuda() for h in host_sets]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, sets, drain=lambda: None, steps=None, warmup=None, prof_range=False):
        steps, warmup = steps or a.steps, a.warmup if warmup is None else warmup
        for i in range(warmup):
            fn(sets[i % len(sets)], i)
        drain()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if prof_range:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(steps):
            fn(sets[i % len(sets)], warmup + i)
        drain()
        e1.record()
        sync_all()
        if prof_range:
            torch.cuda.profiler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    keep = []
    pipe = SynthesisPipeline(model, cmodel, depth=a.depth)
    pipe1 = SynthesisPipeline(model, cmodel, depth=1) if a.depth != 1 else pipe
    tickets = []

    def make_step(p, depth):
        def step_dev(w, i):
            tickets.append((p, p.submit(w, n_steps=N, seed=i, sampler=a.sampler)))
            if len(tickets) > depth:
                q, tk = tickets.pop(0)
                keep[0:1] = [q.result(tk)]
        return step_dev

    def drain():
        while tickets:
            q, tk = tickets.pop(0)
            keep[0:1] = [q.result(tk)]

    # ---- strong scaling (config 4): the whole job starts and ends on rank 0's HBM; scatter and gather are inside `value`
    def strong_step(w, i):
        loc = scatter_clips(w, total_clips, T_SAMPLES, src=0, device=torch.device("cuda", local)) if world > 1 else w
        outs, tk = [], []
        for s0 in range(0, B, sub):                          # sub-batches bound the workspace; two of them in flight
            tk.append(pipe.submit(loc[s0:s0 + sub], n_steps=N, seed=i, sampler=a.sampler))
            if len(tk) > a.depth:
                outs.append(pipe.result(tk.pop(0)))
        while tk:
            outs.append(pipe.result(tk.pop(0)))
        out = torch.cat(outs) if len(outs) > 1 else outs[0]
        keep[0:1] = [gather_clips(out, total_clips, dst=0) if world > 1 else out]

    model.take_launch_count(); cmodel.take_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    if strong:
        ms = timed(strong_step, job_dev)
    else:
        ms = timed(make_step(pipe, a.depth), dev_sets, drain, prof_range=a.profiler_range)
    clocks = sampler.stop()
    launches = model.take_launch_count() + cmodel.take_launch_count()
    launches = launches * a.steps // (a.steps + a.warmup)
    # ---- the same with strictly one batch in flight (BASELINE quotes config 2 at ONE batch of 32)
    ms_d1 = None
    if not strong and a.depth != 1 and not a.quick:
        ms_d1 = timed(make_step(pipe1, 1), dev_sets, drain, steps=max(2, a.steps // 2), warmup=1)
    # ---- end-to-end through the public API with HOST buffers (H2D + D2H inside the timed region).  N > 1: the clips of the
    # whole job start and end in pinned host memory on rank 0 — H2D, one NCCL scatter over NVLink, decode on every rank, one
    # NCCL gather, D2H (ladiffcodec_b200/shard.py: the only collectives of the path, outside the step loop)
    if a.quick:
        ms_e2e, e2e_steps = None, 1
    elif strong:
        job_host = [j.cpu().pin_memory() for j in job_dev] if rank == 0 else [None] * n_sets
        host_out = torch.empty(total_clips, 1, T_SAMPLES, pin_memory=True) if rank == 0 else None

        def e2e_strong(w, i):
            strong_step(w.cuda(non_blocking=True) if rank == 0 else None, i)
            if rank == 0:
                host_out.copy_(keep[0], non_blocking=True)
                torch.cuda.current_stream().synchronize()
        ms_e2e = timed(e2e_strong, job_host, steps=max(1, a.steps // 2), warmup=1)
        e2e_steps = max(1, a.steps // 2)
    elif world == 1:
        ms_e2e = timed(make_step(pipe, a.depth), host_sets, drain)      # host tensors: H2D and D2H ride on the pipeline's streams
        e2e_steps = a.steps
    else:
        job_sets = [torch.cat([make_clips(B, T_SAMPLES, seed=9000 + 1000 * r + 37 * s) for r in range(world)]).pin_memory()
                    for s in range(n_sets)] if rank == 0 else [None] * n_sets
        host_outs = [torch.empty(world * B, 1, T_SAMPLES, pin_memory=True) for _ in range(a.depth + 1)] if rank == 0 else None
        jobs = []

        def finish(job):                      # the current stream waits for that job's decode, then ONE gather and the D2H copy
            out = gather_clips(pipe.result(job["ticket"]), world * B, dst=0)
            if rank == 0:
                host_outs[job["i"] % len(host_outs)].copy_(out, non_blocking=True)
                torch.cuda.current_stream().synchronize()
            keep[0:1] = [out]

        def e2e_step(w, i):                   # H2D on rank 0, ONE scatter, decode enqueued on the pipeline (two jobs in flight)
            dev_all = w.cuda(non_blocking=True) if rank == 0 else None
            loc = scatter_clips(dev_all, world * B, T_SAMPLES, src=0, device=torch.device("cuda", local))
            jobs.append(dict(ticket=pipe.submit(loc, n_steps=N, seed=i, sampler=a.sampler), i=i))
            if len(jobs) > a.depth:
                finish(jobs.pop(0))

        def e2e_drain():
            while jobs:
                finish(jobs.pop(0))
        ms_e2e = timed(e2e_step, job_sets, e2e_drain)
        e2e_steps = a.steps
    # ---- roofline of the dominant kernel: the timed region replays each UNet evaluation as ONE CUDA graph, so the per-kernel
    # CUDA events are taken in one more pass of the same workload right after it, launched kernel by kernel on the same stream
    probe = dev_sets[0] if not strong else make_clips(sub, T_SAMPLES, seed=77).cuda()
    profiling.enable(model, 1)
    synthesize(model, cmodel, probe, n_steps=min(N, 50), noise=None, seed=12345)
    torch.cuda.synchronize()
    prof = profiling.report(model)
    if a.dump_profile and rank == 0:     # plus events around every op of the evaluation
        profiling.enable(model, 2)
        synthesize(model, cmodel, probe, n_steps=2, noise=None, seed=1)
        rows = profiling.dump(model)
        os.makedirs(os.path.dirname(os.path.abspath(a.dump_profile)), exist_ok=True)
        with open(a.dump_profile, "w") as f:
            f.write("# every launch group of one UNet evaluation (CUDA events, in-stream): ms, algorithmic GFLOP, TFLOP/s, label\n")
            tot = sum(r[0] for r in rows)
            for ms_i, gf, label in rows:
                f.write(f"{ms_i:.4f} {gf:9.3f} {gf / ms_i if ms_i > 0 else 0:8.1f}  {label}\n")
            f.write(f"# sum {tot:.3f} ms; conv {sum(r[0] for r in rows if r[1] > 0):.3f} ms; other {sum(r[0] for r in rows if r[1] == 0):.3f} ms\n")
    profiling.enable(model, False)
    # ---- the same attribution INSIDE the CUDA-graph replay the timed region runs (PDL overlap between neighbouring kernels included):
    # per-step time of the DDPM loop with all kernels, without the conv launches, and without the k >= 3 conv launches
    # (ladiff_set_skip_ops); a class's time is the difference.  Per-step = (t(25 steps) - t(5 steps)) / 20, best of 3.
    graph_attr = None
    if not a.quick or True:
        cond_p = cmodel.get_cond(probe)
        x_p = torch.randn(probe.shape[0], 128, T_SAMPLES // int(math.prod(args.enc_ratios)), device="cuda")

        def loop_ms(k):
            best = 1e30
            for _ in range(3):
                x = x_p.clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                model.diffusion._steps(x, cond_p, 60, k, None, 7)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best

        def per_step(mask):
            _lib.check(model._lib.ladiff_set_skip_ops(model._h, mask), "set_skip_ops")
            loop_ms(3)                                   # re-capture + warm
            return (loop_ms(25) - loop_ms(5)) / 20.0

        t_all, t_noconv, t_nobig = per_step(0), per_step(24), per_step(16)
        _lib.check(model._lib.ladiff_set_skip_ops(model._h, 0), "set_skip_ops")
        graph_attr = dict(ms_per_ddpm_step=t_all, conv_ms=t_all - t_noconv, conv_k3plus_ms=t_all - t_nobig, other_ms=t_noconv)

    audio_s = total_clips * CLIP_SECONDS
    value = audio_s * a.steps / (ms / 1e3)
    e2e = audio_s * e2e_steps / (ms_e2e / 1e3) if ms_e2e else None
    pk = peaks()
    roof = None
    traffic, traffic_note = None, "no ncu --set full capture of this build and workload committed"
    tp = os.path.join(ROOT, "profiles", a.traffic_dir, "conv_dram_traffic.json") if a.traffic_dir else None
    if tp and a.config == 2 and not a.batch and os.path.exists(tp):      # ncu capture of the same build and workload (per launch, like `achieved`)
        t = json.load(open(tp))
        traffic = (t["dram_read_bytes"] + t["dram_write_bytes"]) / t["conv_launches"]
        traffic_note = (f"profiles/{a.traffic_dir}/conv_dram_traffic.json: DRAM bytes per conv launch, mean over the {t['conv_launches']} launches of one "
                        f"UNet evaluation (ncu flushes L2 before each launch; L2->SM traffic {t['l2_bytes'] / t['conv_launches'] / 1e6:.0f} MB per launch)")
    if prof and prof["conv_ms"] > 0:
        ach_eager = prof["conv_flops"] / (prof["conv_ms"] * 1e-3) / 1e12
        ach = prof["conv_flops"] / (graph_attr["conv_ms"] * 1e-3) / 1e12 if graph_attr and graph_attr["conv_ms"] > 0 else ach_eager
        roof = dict(bound="tensor", kernel="tc_conv_kernel (tcgen05 implicit-GEMM Conv1d)", achieved=ach, peak=pk["bf16_sustained"],
                    unit="TFLOP/s", frac=ach / pk["bf16_sustained"], traffic=traffic, traffic_source=traffic_note,
                    achieved_eager_events=ach_eager, frac_eager_events=ach_eager / pk["bf16_sustained"], in_graph=graph_attr,
                    peak_source=pk["src"] + " 16-bit dense sustained (cuBLAS bf16; tcgen05 kind::f16 runs fp16 and bf16 at one rate)",
                    launches_per_unet_eval=prof["conv_launches"], conv_ms_per_unet_eval=prof["conv_ms"],
                    unet_eval_ms=prof["eval_ms"], algorithmic_gflop_per_clip_eval=prof["conv_flops"] / probe.shape[0] / 1e9,
                    how="achieved = algorithmic FLOPs of the conv launches of one UNet evaluation / their time inside the CUDA-graph replay "
                        "the timed region runs: CUDA events around 20 DDPM steps with and without the conv launches (ladiff_set_skip_ops), "
                        "the difference is the convs' time with the programmatic-dependent-launch overlap of the real step; "
                        "achieved_eager_events = the same FLOPs / CUDA events around every conv launch of one evaluation launched kernel "
                        "by kernel (no graph: each launch pays its own launch gap and event pair)",
                    conv_share_of_unet_eval=(graph_attr["conv_ms"] / graph_attr["ms_per_ddpm_step"]) if graph_attr else None,
                    conv_share_eager=prof["conv_ms"] / prof["eval_ms"] if prof["eval_ms"] else None,
                    whole_pass_frac=None)
        flop_pass = probe.shape[0] * (N * prof["conv_flops"] / probe.shape[0] + 7.3e9)     # + codec, SURVEY §8d
        per_batch_ms = ms / a.steps * (probe.shape[0] / max(B, 1)) if not strong else None
        if per_batch_ms:
            roof["whole_pass_frac"] = flop_pass / (per_batch_ms * 1e-3) / 1e12 / pk["bf16_sustained"]
    L = T_SAMPLES // int(math.prod(args.enc_ratios))
    line = dict(metric="audio-sec/s decoded", value=value, unit="audio-s/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms / a.steps, higher_is_better=True, scaling=cfg["scaling"], vs_baseline=None, dtype=dtype, data="synthetic",
                config=dict(workload=cfg["name"], batch_per_gpu=B, clips_total=total_clips, n_ddpm_steps=N, sampler=a.sampler, clip_seconds=CLIP_SECONDS,
                            latent_len=L, noise="in-kernel Philox keyed by (seed, timestep, global clip)", batches_in_flight=a.depth,
                            sub_batch=sub if strong else None,
                            l2="no explicit flush: the per-step working set (weights 271 MB + activations >0.4 GB per batch) exceeds the 126 MB L2 "
                               "and every timed step decodes different clips"),
                clocks=clocks,
                value_depth1=(audio_s * max(2, a.steps // 2) / (ms_d1 / 1e3)) if ms_d1 else (value if a.depth == 1 else None),
                e2e=dict(value=e2e, unit="audio-s/s", h2d_bytes_per_step=total_clips * T_SAMPLES * 4, d2h_bytes_per_step=total_clips * T_SAMPLES * 4,
                         ms_per_step=ms_e2e / e2e_steps if ms_e2e else None,
                         path=("pinned host -> H2D -> synthesize -> D2H" if world == 1 else
                               "rank 0 pinned host -> H2D -> NCCL scatter -> pipelined synthesize on every rank -> NCCL gather -> D2H on rank 0")),
                gpu_launches=int(launches), roofline=roof, parity=parity_block(dtype))
    if a.quick:
        line["quick"] = "1 warm-up pass, no end-to-end / depth-1 legs (step-count sweeps; not a driver measurement)"
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            clips = min(cfg["batch"], 8)
            v, d = time_reference_sample(args, sdm, sdc, N, clips, 3, threads)
            line["cpu_baseline"] = dict(value=v, unit="audio-s/s", cores=threads, kind="port", sample=cpu_sample_desc(clips, 3, N), detail=d)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def relaunch_under_torchrun(a):
    """`python bench.py --gpus N` outside torchrun: one process per GPU over NCCL, as the driver launches it."""
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    raise SystemExit(subprocess.call(cmd))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (strong scaling: clips in all); default: the config's")
    ap.add_argument("--ddpm_steps", type=int, default=0)
    ap.add_argument("--sampler", default="ddpm", choices=["ddpm", "ddim"],
                    help="ddpm: the script's halfway_sampling(t = N) from the condition; ddim: the reference's ddim_sample (N steps from noise)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--profiler_range", action="store_true",
                    help="bracket the timed `value` region with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    ap.add_argument("--depth", type=int, default=2, help="batches in flight (SynthesisPipeline); 1 = strictly one pass at a time")
    ap.add_argument("--dump_profile", default="", help="write the per-conv-launch table of one UNet evaluation here")
    ap.add_argument("--quick", action="store_true", help="1 warm-up pass, skip the e2e and depth-1 legs (long sweeps)")
    ap.add_argument("--traffic_dir", default="r2c", help="profiles/<dir>/conv_dram_traffic.json holds the ncu DRAM-traffic capture of this build")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.ddpm_steps:
        cfg["n_steps"] = a.ddpm_steps
    if a.impl == "ours" and a.gpus > 1 and "WORLD_SIZE" not in os.environ:
        relaunch_under_torchrun(a)
    if a.warmup < 3 and a.impl == "ours" and cfg["scaling"] != "strong" and not a.quick:
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a, cfg)
    elif cfg.get("sweep") and not a.ddpm_steps:
        for n in cfg["sweep"]:                      # config 5: one JSON line per step count
            c = dict(cfg, n_steps=n)
            run_ours(a, c)
    else:
        run_ours(a, cfg)


if __name__ == "__main__":
    main()
