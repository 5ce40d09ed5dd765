"""Per-launch times of the codec stages at config 2 (LADIFF_CODEC_PROF=1: events around every codec launch, serialised)."""
import os, sys
os.environ["LADIFF_CODEC_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs
from ladiffcodec_b200.model import DiffAudioRep
from ladiffcodec_b200.synthetic import make_clips
from ladiffcodec_b200.utils import load_model
cfg = bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
args, sdm, sdc = bench.build_state(cfg)
m = DiffAudioRep(**ladiff_model_kwargs(args)).to("cuda"); load_model(m, sdm, strict=True)
c = DiffAudioRep(**cond_model_kwargs(args)).to("cuda"); load_model(c, sdc)
B = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["batch"]
wav = make_clips(B, bench.T_SAMPLES, seed=77).cuda()
for rep in range(2):
    print(f"==== pass {rep}: get_cond", file=sys.stderr)
    cond = c.get_cond(wav)
    img = cond
    print("==== cond upsample", file=sys.stderr)
    for layer in m.diff_model.upsampling_layers:
        img = layer(img)
    print("==== decoder", file=sys.stderr)
    y = m.decoder(img)
    torch.cuda.synchronize()
