"""Experiment: decode one batch as K independent sub-batches on K CUDA streams (same handles, separate workspaces), so
that the latency-bound tails of one sub-batch's kernels overlap the other's.  Prints ms per pass for K = 1, 2, 4.

    python profiles/two_stream.py [--config 2] [--batch 32] [--n_steps 50]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from ladiffcodec_b200 import _lib
    from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs
    from ladiffcodec_b200.model import DiffAudioRep, _ptr
    from ladiffcodec_b200.synthetic import make_clips
    from ladiffcodec_b200.utils import load_model

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--n_steps", type=int, default=50)
    ap.add_argument("--splits", default="1,2,4")
    ap.add_argument("--full", action="store_true", help="every stream decodes a FULL batch (K batches in flight) instead of B/K clips")
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    args, sdm, sdc = bench.build_state(cfg)
    model = DiffAudioRep(**ladiff_model_kwargs(args)).to("cuda")
    load_model(model, sdm, strict=True)
    cmodel = DiffAudioRep(**cond_model_kwargs(args)).to("cuda")
    load_model(cmodel, sdc)
    B = a.batch or cfg["batch"]
    T = bench.T_SAMPLES
    wav = make_clips(B, T, seed=77).cuda()
    lib = model._lib
    res = {}
    for K in [int(s) for s in a.splits.split(",")]:
        Bs = B if a.full else B // K
        streams = [torch.cuda.Stream() for _ in range(K)]
        need = lib.ladiff_synthesize_workspace_bytes(model._h, cmodel._h, Bs, T)
        wss = [torch.empty(int(need) + 1024, dtype=torch.uint8, device="cuda") for _ in range(K)]
        outs = [torch.empty(Bs, 1, T, device="cuda") for _ in range(K)]
        ins = [wav.clone() for k in range(K)] if a.full else [wav[k * Bs:(k + 1) * Bs].contiguous() for k in range(K)]

        def one_pass(seed):
            cur = torch.cuda.current_stream()
            for k in range(K):
                streams[k].wait_stream(cur)
                _lib.check(lib.ladiff_synthesize(model._h, cmodel._h, _ptr(ins[k]), Bs, T, a.n_steps, None, 0, ctypes.c_uint64(seed + k),
                                                 _ptr(outs[k]), None, _ptr(wss[k]), wss[k].numel(),
                                                 ctypes.c_void_p(streams[k].cuda_stream)), "synthesize")
            for k in range(K):
                cur.wait_stream(streams[k])

        best = 1e30
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            one_pass(10 * i)
            e1.record()
            torch.cuda.synchronize()
            if i:
                best = min(best, e0.elapsed_time(e1))
        res[K] = dict(ms=best, audio_s_per_s=(B * K if a.full else B) * 2.4 / (best * 1e-3), absmax=float(torch.stack(outs).abs().max()))
        del wss, outs
    print(json.dumps(res))


if __name__ == "__main__":
    main()
