"""Flag surface of the sampling path.

Mirrors the argparse block of the reference's ``srcs/sample.py:141-201`` (same names, same
defaults) and the constructor defaults of ``DiffAudioRep`` (``srcs/model.py:34``), so that
``DiffAudioRep(other_cond=..., **vars(args))`` means the same thing here as there.
"""
import argparse
import math

# sample.py:141-201 — name -> (type, default)
SAMPLE_DEFAULTS = dict(
    data_folder_path="/data/hy17/librispeech/librispeech", n_spks=500, seq_len_in_sec=1.8, sample_rate=16000,
    model_path="", qtzer_path="", note="",
    rep_dims=128, emb_dims=128, quantization=False, bandwidth=3.0, n_filters=32, lstm=2, n_residual_layers=1,
    enc_ratios=[8], final_activation=None,
    run_diff=False, run_vae=False, train_time_diff=False,
    diff_dims=256, qtz_condition=False, self_condition=False, seq_length=16000, model_type="unet",
    scaling_frame=False, scaling_feature=False, scaling_global=False, scaling_dim=False,
    sampling_timesteps=1000, use_film=False,
    model_for_cond="", upsampling_ratios=[5, 4, 2], cond_enc_ratios=[8, 5, 4, 2],
    cond_bandwidth=3.0, cond_global=3.0, unet_scale_cond=False, unet_scale_x=False,
    input_dir="", output_dir="outputs/",
)

# model.py:34 — DiffAudioRep.__init__ keyword defaults
MODEL_DEFAULTS = dict(
    rep_dims=128, emb_dims=128, diff_dims=128, norm="weight_norm", causal=True, dilation_base=2,
    n_residual_layers=1, n_filters=32, lstm=0, quantization=False, bandwidth=3, sample_rate=16000,
    qtz_condition=False, self_condition=False, other_cond=False, seq_length=320, enc_ratios=[8, 5, 4, 2],
    run_diff=False, run_vae=False, model_type="", scaling_frame=False, scaling_feature=False,
    scaling_global=False, scaling_dim=False, freeze_ed=False, final_activation=None,
    sampling_timesteps=None, use_film=False, cond_global=1, cond_channels=128,
    upsampling_ratios=[5, 4, 2], unet_scale_x=False, unet_scale_cond=True,
)

UNET_DIM_MULTS = (1, 2, 2, 4, 4)       # model.py:74
NUM_TIMESTEPS = 1000                   # ddpm_loss.py:84
CODEBOOK_BINS = 1024                   # vq.py:46
ATTN_HEADS, ATTN_DIM_HEAD = 4, 32      # unet.py:195,225


def sample_args(**overrides):
    """An argparse.Namespace carrying sample.py's defaults (+ overrides)."""
    d = {k: (list(v) if isinstance(v, list) else v) for k, v in SAMPLE_DEFAULTS.items()}
    for k, v in overrides.items():
        if k not in d:
            raise KeyError(f"unknown sample.py flag: {k}")
        d[k] = v
    return argparse.Namespace(**d)


def readme_args(kbps=3.0, **overrides):
    """The README.md:35-39 command line (pretrained layout, "Layout-A")."""
    kw = dict(run_diff=True, scaling_global=True, cond_bandwidth=kbps, unet_scale_cond=True,
              model_for_cond="cond.amlt", model_path="ladiff.amlt")
    kw.update(overrides)
    return sample_args(**kw)


def num_quantizers(bandwidth, hop_length, sample_rate=16000):
    """model.py:64-65 — n_q as constructed."""
    frame_rate = sample_rate / hop_length
    return int(1000 * bandwidth // (math.ceil(frame_rate) * 10))


def num_quantizers_at_call(bandwidth, frame_rate, n_q_built, bins=CODEBOOK_BINS):
    """vq.py:86-98 — n_q actually used by ResidualVectorQuantizer.forward."""
    bw_per_q = math.log2(bins) * frame_rate / 1000
    n_q = n_q_built
    if bandwidth and bandwidth > 0.0:
        n_q = int(max(1, math.floor(bandwidth / bw_per_q)))
    return n_q


def build_parser():
    """The reference's CLI, flag for flag (sample.py:141-199)."""
    p = argparse.ArgumentParser(description="Encodec_baseline")
    for name, default in SAMPLE_DEFAULTS.items():
        if isinstance(default, bool):
            p.add_argument(f"--{name}", dest=name, action="store_true")
        elif isinstance(default, list):
            p.add_argument(f"--{name}", nargs="+", type=int, default=list(default))
        elif default is None:
            p.add_argument(f"--{name}", type=str, default=None)
        else:
            p.add_argument(f"--{name}", type=type(default), default=default)
    return p
