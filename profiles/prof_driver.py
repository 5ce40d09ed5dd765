"""Profiling driver: one warmed-up pass of the hot path (ladiff_synthesize) bracketed by cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/prof_driver.py --config 2 --ddpm_steps 2
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_conv -c 8 \
        -o gpurun_out/tc_conv python profiles/prof_driver.py --config 2 --ddpm_steps 1

Numbers printed under a profiler are never bench values (bench.py is the measurement).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs
    from ladiffcodec_b200.model import DiffAudioRep
    from ladiffcodec_b200.sample import synthesize
    from ladiffcodec_b200.synthetic import make_clips
    from ladiffcodec_b200.utils import load_model

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--ddpm_steps", type=int, default=2)
    ap.add_argument("--warm", type=int, default=1)
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    args, sdm, sdc = bench.build_state(cfg)
    model = DiffAudioRep(**ladiff_model_kwargs(args)).to("cuda")
    load_model(model, sdm, strict=True)
    cmodel = DiffAudioRep(**cond_model_kwargs(args)).to("cuda")
    load_model(cmodel, sdc)
    B = a.batch or cfg["batch"]
    wav = make_clips(B, bench.T_SAMPLES, seed=77).cuda()
    for i in range(a.warm):
        synthesize(model, cmodel, wav, n_steps=a.ddpm_steps, noise=None, seed=i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = synthesize(model, cmodel, wav, n_steps=a.ddpm_steps, noise=None, seed=99)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("prof_driver: done", tuple(out.shape), float(out.abs().max()))


if __name__ == "__main__":
    main()
