"""Operator-level sweep of the tcgen05 conv kernel over the conv shapes of one UNet evaluation (config 2: B=32, L=1200).
    python profiles/conv_sweep.py [quick]        (on the GPU box; with LADIFF_TC_PROF=1 the kernel prints its cycle counters)
Prints, per shape and tile-shape variant: mean µs per launch (CUDA events around 20 back-to-back launches) and TFLOP/s."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ladiffcodec_b200 import _lib

lib = _lib.get_lib()
P = ctypes.c_void_p
B = int(os.environ.get("SWEEP_B", "32"))
SHAPES = [  # (L, Cin, Cout, k) of config 2 (Layout-B); Cout doubled where res_conv rides along
    (1200, 256, 256, 3), (1200, 512, 512, 3), (1200, 256, 384, 1), (1200, 128, 256, 1), (1200, 256, 256, 7),
    (600, 256, 256, 3), (600, 768, 1024, 3), (600, 512, 512, 3), (300, 512, 512, 3), (300, 1024, 1024, 3),
    (150, 512, 512, 3), (150, 1024, 1024, 3), (150, 1536, 2048, 3), (75, 1024, 1024, 3), (75, 2048, 2048, 3), (75, 1024, 384, 1),
]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    SHAPES = SHAPES[:2] + [SHAPES[8], SHAPES[13]]
VARIANTS = [("model", dict()), ("NT256", dict(nt=256)), ("NT208", dict(nt=208)), ("NT160", dict(nt=160)), ("NT128", dict(nt=128)),
            ("pair", dict(two=2)), ("pair NT208", dict(nt=208, two=2)), ("pair NT160", dict(nt=160, two=2)), ("pair NT256", dict(nt=256, two=2)),
            ("two/SM NT128", dict(nt=128, two=1)), ("two/SM NT96", dict(nt=96, two=1)), ("posM", dict(t=1)), ("posM pair", dict(t=2))]
for (L, Cin, Cout, k) in SHAPES:
    g = torch.Generator().manual_seed(L + Cin)
    x = torch.randn(B, L, Cin, generator=g).to(_lib.act_dtype()).cuda()
    w = (torch.randn(Cout, Cin, k, generator=g) * (Cin * k) ** -0.5).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    flops = 2.0 * Cout * Cin * k * L * B
    print(f"--- L={L} Cin={Cin} Cout={Cout} k={k}: {flops / 1e9:.1f} GFLOP, ideal {flops / 1398.1e12 * 1e6:.1f} us at 1398 TFLOP/s")
    for name, v in VARIANTS:
        ms = ctypes.c_float()
        label = ctypes.create_string_buffer(200)
        rc = lib.ladiff_op_conv1d_bench(P(x.data_ptr()), P(w.data_ptr()), P(bias.data_ptr()), B, L, Cin, Cout, k, v.get("nt", 0), v.get("nclip", 0),
                                        v.get("two", 0), v.get("t", 0), 1, 5, 20, ctypes.byref(ms), label, 200)
        if rc != 0:
            continue
        print(f"  {name:14s} {ms.value * 1e3:7.2f} us  {flops / (ms.value * 1e-3) / 1e12:7.1f} TFLOP/s   {label.value.decode()}")
