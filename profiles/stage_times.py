"""Splits the hot path's time into its fixed part (cond codec + upsample + decoder) and the per-DDPM-step part.

    python profiles/stage_times.py [--config 2] [--batch B] [--steps_a 10] [--steps_b 50]

Times ladiff_synthesize (CUDA events, best of 3 after a warm-up) at two step counts; the difference is the per-step
cost as it runs in production (one CUDA-graph replay per UNet evaluation + the posterior kernel).  Environment knobs
read by the library (LADIFF_NO_GRAPH, LADIFF_NO_PDL) can be set by the caller to compare launch modes.
"""
import argparse
import json
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
    ap.add_argument("--steps_a", type=int, default=10)
    ap.add_argument("--steps_b", type=int, default=50)
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    args, sdm, sdc = bench.build_state(cfg)
    model = DiffAudioRep(**ladiff_model_kwargs(args)).to("cuda")
    load_model(model, sdm, strict=True)
    cmodel = DiffAudioRep(**cond_model_kwargs(args)).to("cuda")
    load_model(cmodel, sdc)
    B = a.batch or cfg["batch"]
    wav = make_clips(B, bench.T_SAMPLES, seed=77).cuda()

    def t(n):
        best = 1e30
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            synthesize(model, cmodel, wav, n_steps=n, noise=None, seed=i)
            e1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1))
        return best

    model.take_launch_count(); cmodel.take_launch_count()
    ta = t(a.steps_a)
    launches_a = (model.take_launch_count() + cmodel.take_launch_count()) // 4
    tb = t(a.steps_b)
    per_step = (tb - ta) / (a.steps_b - a.steps_a)
    out = dict(config=a.config, batch=B, ms_a=ta, ms_b=tb, steps_a=a.steps_a, steps_b=a.steps_b, ms_per_ddpm_step=per_step,
               ms_fixed=ta - a.steps_a * per_step, launches_per_pass_a=launches_a,
               env={k: os.environ[k] for k in os.environ if k.startswith("LADIFF_")})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
