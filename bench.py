#!/usr/bin/env python
"""bench.py — audio-seconds decoded per wall-second on the LaDiffCodec sampling path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--config 2|3]

One "step" = one pass of the hot path (sample.py:94-134: get_cond → upsample → N-step DDPM → decoder →
normalise) over one batch of B synthetic 2.4 s / 16 kHz clips per GPU.  Default workload = BASELINE config 2:
B=32, 3 kbps, enc_ratios [8,4] (+ upsampling [5,2], latent L=1200), diff_dims 256, 50 DDPM steps.
N > 1: one process per GPU (torchrun), clips sharded across ranks, no data-path collective (weak scaling).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP_SECONDS = 2.4
T_SAMPLES = 38400

CONFIGS = {
    2: dict(name="config2: B=32/GPU, 3 kbps, enc_ratios [8,4], upsampling [5,2] (L=1200), diff_dims 256, 50-step DDPM",
            flags=dict(run_diff=True, cond_bandwidth=3.0, enc_ratios=[8, 4], upsampling_ratios=[5, 2], diff_dims=256,
                       model_for_cond="c", model_path="m"), batch=32, n_steps=50),
    3: dict(name="config3: B=256, 1.5 kbps, enc_ratios [8] (L=4800), scaling_global + unet_scale_cond, 200-step DDPM",
            flags=dict(run_diff=True, cond_bandwidth=1.5, scaling_global=True, unet_scale_cond=True, diff_dims=256,
                       model_for_cond="c", model_path="m"), batch=256, n_steps=200),
}


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
    """Times the oracle port of the reference's CPU path (same ATen kernels the reference dispatches to) on `clips`
    clips: the complete codec stages + `ddpm_steps_timed` DDPM steps (after one untimed step), extrapolated linearly
    to n_steps (every step is identical work).  Returns (audio_s_per_s, detail)."""
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
        t_step = (time.perf_counter() - t0) / ddpm_steps_timed
        t0 = time.perf_counter()
        y = O.seanet_decoder(x, sdm, list(args.enc_ratios), fast_lstm=True)
        y = y / (y.reshape(clips, -1).std(1).reshape(clips, 1, 1) + 1e-8)
        y = y / (y.abs().reshape(clips, -1).max(1).values.reshape(clips, 1, 1) + 1e-8)
        t_back = time.perf_counter() - t0
    total = t_front + n_steps * t_step + t_back
    detail = dict(clips=clips, t_codec_front_s=round(t_front, 3), t_ddpm_step_s=round(t_step, 3), t_decoder_s=round(t_back, 3),
                  ddpm_steps_timed=ddpm_steps_timed, extrapolated_total_s=round(total, 2))
    return clips * CLIP_SECONDS / total, detail


def run_reference(a, cfg):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; /root/reference cannot
    travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    args, sdm, sdc = build_state(cfg)
    threads = os.cpu_count() or 1
    clips, timed = 4, 2
    vals, last = [], None
    for i in range(a.warmup + a.steps):
        v, d = time_reference_sample(args, sdm, sdc, cfg["n_steps"], clips, timed, threads)
        if i >= a.warmup:
            vals.append(v); last = d
    value = statistics.mean(vals)
    sample = (f"{clips} clips: full get_cond + upsample + decoder, {timed} timed DDPM steps (1 untimed), extrapolated linearly to "
              f"{cfg['n_steps']} steps; oracle port of the reference (same ATen conv/LSTM kernels), fp32")
    line = dict(metric="audio-sec/s decoded", value=value, unit="audio-s/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
                ms_per_step=1000.0 * clips * CLIP_SECONDS / value, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=cfg["name"], n_ddpm_steps=cfg["n_steps"], clip_seconds=CLIP_SECONDS),
                cpu_baseline=dict(value=value, unit="audio-s/s", cores=threads, kind="port", sample=sample, detail=last),
                e2e=dict(value=value, unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def run_ours(a, cfg):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):    # keeps stdout to the ONE JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs
    from ladiffcodec_b200.model import DiffAudioRep
    from ladiffcodec_b200.sample import synthesize
    from ladiffcodec_b200.synthetic import make_clips
    from ladiffcodec_b200.utils import load_model
    from ladiffcodec_b200 import profiling

    args, sdm, sdc = build_state(cfg)
    model = DiffAudioRep(**ladiff_model_kwargs(args)).to("cuda")
    load_model(model, sdm, strict=True)
    cmodel = DiffAudioRep(**cond_model_kwargs(args)).to("cuda")
    load_model(cmodel, sdc)
    B, N = a.batch or cfg["batch"], cfg["n_steps"]
    # distinct clips per rank and per step so nothing can be reused between timed iterations
    n_sets = min(a.steps, 4)
    host_sets = [make_clips(B, T_SAMPLES, seed=9000 + 1000 * rank + 37 * s).pin_memory() for s in range(n_sets)]
    dev_sets = [h.cuda() for h in host_sets]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, sets, drain=lambda: None):
        for i in range(a.warmup):
            fn(sets[i % n_sets], i)
        drain()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if a.profiler_range and sets is dev_sets:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(a.steps):
            fn(sets[i % n_sets], a.warmup + i)
        drain()
        e1.record()
        sync_all()
        if a.profiler_range and sets is dev_sets:
            torch.cuda.profiler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    keep = []
    # ---- device-resident throughput (`value`)
    model.take_launch_count(); cmodel.take_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    # two batches in flight (SynthesisPipeline, depth 2): step i+1 is enqueued on the other stream while step i runs; every
    # ticket is collected (the current stream waits for it) before the closing synchronize of the timed region
    from ladiffcodec_b200.sample import SynthesisPipeline
    pipe = SynthesisPipeline(model, cmodel, depth=a.depth)
    tickets = []

    def step_dev(w, i):
        tickets.append(pipe.submit(w, n_steps=N, seed=i))
        if len(tickets) > a.depth:
            keep[0:1] = [pipe.result(tickets.pop(0))]

    def drain():
        while tickets:
            keep[0:1] = [pipe.result(tickets.pop(0))]

    ms = timed(step_dev, dev_sets, drain)
    clocks = sampler.stop()
    launches = model.take_launch_count() + cmodel.take_launch_count()
    launches = launches * a.steps // (a.steps + a.warmup)
    # ---- end-to-end through the public API with HOST buffers (H2D + D2H inside the timed region).  N > 1: the clips of the
    # whole job start and end in pinned host memory on rank 0 — H2D, one NCCL scatter over NVLink, decode on every rank, one
    # NCCL gather, D2H (ladiffcodec_b200/shard.py: the only collectives of the path, outside the step loop)
    if world == 1:
        ms_e2e = timed(step_dev, host_sets, drain)      # host tensors: H2D and D2H ride on the pipeline's streams
    else:
        from ladiffcodec_b200.shard import scatter_clips, gather_clips
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
            jobs.append(dict(ticket=pipe.submit(loc, n_steps=N, seed=i), i=i))
            if len(jobs) > a.depth:
                finish(jobs.pop(0))

        def e2e_drain():
            while jobs:
                finish(jobs.pop(0))
        ms_e2e = timed(e2e_step, job_sets, e2e_drain)
    # ---- roofline of the dominant kernel: the timed region replays each UNet evaluation as ONE CUDA graph, so the per-kernel
    # CUDA events are taken in one more pass of the same workload right after it, launched kernel by kernel on the same stream
    profiling.enable(model, 1)
    synthesize(model, cmodel, dev_sets[0], n_steps=N, noise=None, seed=12345)
    torch.cuda.synchronize()
    prof = profiling.report(model)
    if a.dump_profile and rank == 0:     # plus events around every op of the evaluation
        profiling.enable(model, 2)
        synthesize(model, cmodel, dev_sets[0], n_steps=2, noise=None, seed=1)
        rows = profiling.dump(model)
        os.makedirs(os.path.dirname(os.path.abspath(a.dump_profile)), exist_ok=True)
        with open(a.dump_profile, "w") as f:
            f.write("# every launch group of one UNet evaluation (CUDA events, in-stream): ms, algorithmic GFLOP, TFLOP/s, label\n")
            tot = sum(r[0] for r in rows)
            for ms_i, gf, label in rows:
                f.write(f"{ms_i:.4f} {gf:9.3f} {gf / ms_i if ms_i > 0 else 0:8.1f}  {label}\n")
            f.write(f"# sum {tot:.3f} ms; conv {sum(r[0] for r in rows if r[1] > 0):.3f} ms; other {sum(r[0] for r in rows if r[1] == 0):.3f} ms\n")
    profiling.enable(model, False)

    audio_s = world * B * CLIP_SECONDS * a.steps
    value, e2e = audio_s / (ms / 1e3), audio_s / (ms_e2e / 1e3)
    pk = peaks()
    roof = None
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "r1d", "conv_dram_traffic.json")
    if a.config == 2 and not a.batch and os.path.exists(tp):      # committed ncu capture of the same workload (per launch, like `achieved`)
        t = json.load(open(tp))
        traffic = (t["dram_read_bytes"] + t["dram_write_bytes"]) / t["conv_launches"]
        traffic_note = (f"profiles/r1d/conv_dram_traffic.json: DRAM bytes per conv launch, mean over the {t['conv_launches']} launches of one UNet "
                        f"evaluation (ncu flushes L2 before each launch; L2->SM traffic {t['l2_bytes'] / t['conv_launches'] / 1e6:.0f} MB per launch)")
    if prof and prof["conv_ms"] > 0:
        ach = prof["conv_flops"] / (prof["conv_ms"] * 1e-3) / 1e12
        roof = dict(bound="tensor", kernel="tc_conv_kernel (tcgen05 implicit-GEMM Conv1d)", achieved=ach, peak=pk["bf16_sustained"],
                    unit="TFLOP/s", frac=ach / pk["bf16_sustained"], traffic=traffic, traffic_source=traffic_note, peak_source=pk["src"] + " bf16 sustained",
                    launches_per_unet_eval=prof["conv_launches"], conv_ms_per_unet_eval=prof["conv_ms"],
                    unet_eval_ms=prof["eval_ms"], algorithmic_gflop_per_clip_eval=prof["conv_flops"] / B / 1e9,
                    how="CUDA events around each of the conv launches of one UNet evaluation, eager pass right after the timed region "
                        "(the timed region replays the evaluation as a CUDA graph)",
                    conv_share_of_unet_eval=prof["conv_ms"] / prof["eval_ms"] if prof["eval_ms"] else None)
    line = dict(metric="audio-sec/s decoded", value=value, unit="audio-s/s", n_gpus=world, steps=a.steps, warmup=a.warmup,
                ms_per_step=ms / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                config=dict(workload=cfg["name"], batch_per_gpu=B, n_ddpm_steps=N, clip_seconds=CLIP_SECONDS, latent_len=T_SAMPLES // int(
                    __import__("math").prod(args.enc_ratios)), noise="in-kernel Philox", batches_in_flight=a.depth,
                    l2="no explicit flush: the per-step working set (271 MB bf16 weights + >0.4 GB activations) exceeds the 126 MB L2 "
                       "and every timed step decodes different clips"),
                clocks=clocks,
                e2e=dict(value=e2e, unit="audio-s/s", h2d_bytes_per_step=world * B * T_SAMPLES * 4, d2h_bytes_per_step=world * B * T_SAMPLES * 4,
                         ms_per_step=ms_e2e / a.steps,
                         path=("pinned host -> H2D -> synthesize -> D2H" if world == 1 else
                               "rank 0 pinned host -> H2D -> NCCL scatter -> pipelined synthesize on every rank -> NCCL gather -> D2H on rank 0")),
                gpu_launches=int(launches), roofline=roof)
    if rank == 0:
        if world == 1 and not a.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, d = time_reference_sample(args, sdm, sdc, N, 4, 2, threads)
            line["cpu_baseline"] = dict(value=v, unit="audio-s/s", cores=threads, kind="port",
                                        sample="4 clips: full codec stages + 2 timed DDPM steps extrapolated linearly to "
                                               f"{N} steps (oracle port of the reference, fp32 ATen CPU kernels)", detail=d)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="clips per GPU (default: the config's)")
    ap.add_argument("--ddpm_steps", type=int, default=0)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--profiler_range", action="store_true",
                    help="bracket the timed `value` region with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    ap.add_argument("--depth", type=int, default=2, help="batches in flight (SynthesisPipeline); 1 = strictly one pass at a time")
    ap.add_argument("--dump_profile", default="", help="write the per-conv-launch table of one UNet evaluation here")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.ddpm_steps:
        cfg["n_steps"] = a.ddpm_steps
    if a.warmup < 3 and a.impl == "ours":
        a.warmup = 3
    if a.impl == "reference":
        run_reference(a, cfg)
    else:
        run_ours(a, cfg)


if __name__ == "__main__":
    main()
