"""BASELINE config 1 ("plumbing"): runs the REAL reference's own script entry `srcs.sample.synthesis(inp_args)` (sample.py:50-136)
on CPU over synthetic wav FILES — README flags, seeded synthetic checkpoints written as .amlt files, the script's hard-coded
midway_t = 100 — and commits what it saves.  tests/test_cli_gpu.py runs `ladiffcodec_b200.sample.synthesis` on the same files.

Run in the build container only:   python tests/golden/make_golden_cli.py
Shims: the sys.modules stubs of oracle/ref_import.py, and torchaudio.load / torchaudio.save (TorchCodec is not
in this image) replaced by scipy.io.wavfile equivalents: load → (float32 [1,T] in [-1,1), sr) for 16-bit PCM; save captured in
memory.  The load shim also seeds the global generator (torch.manual_seed(SEED) per file, right where the loop starts on a
file), so a file's noise depends neither on glob order nor on the random initialisation of the modules the script builds first.
"""
import argparse
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

from oracle.ref_import import import_reference, REFERENCE_ROOT      # noqa: E402
from ladiffcodec_b200.config import SAMPLE_DEFAULTS, sample_args      # noqa: E402
from ladiffcodec_b200.layout import ladiff_model_kwargs, cond_model_kwargs  # noqa: E402
from ladiffcodec_b200.synthetic import make_state_dict, make_clips  # noqa: E402

SEED_MODEL, SEED_COND, SEED_NOISE = 31, 32, 33
FILES = {"spk1/utt_a.wav": dict(sr=16000, n=12000, seed=41), "spk2/deep/utt_b.wav": dict(sr=8000, n=4100, seed=42)}
FLAGS = dict(run_diff=True, scaling_global=True, cond_bandwidth=3.0, unet_scale_cond=True)      # README.md:35


def write_inputs(root):
    from scipy.io import wavfile
    for rel, f in FILES.items():
        x = make_clips(1, f["n"], seed=f["seed"])[0, 0].numpy()
        pcm = np.clip(np.round(x * 32767.0 * 0.9), -32768, 32767).astype(np.int16)
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        wavfile.write(p, f["sr"], pcm)


def main():
    torch.set_num_threads(os.cpu_count())
    cli = os.path.join(HERE, "cli")
    shutil.rmtree(cli, ignore_errors=True)
    write_inputs(os.path.join(cli, "in"))
    import_reference()
    import torchaudio
    from scipy.io import wavfile
    saved = {}

    def load(path):
        sr, data = wavfile.read(path)
        assert data.dtype == np.int16
        # synthesis() builds its models (random init = generator draws) before the file loop: seed HERE, at the first thing
        # the loop does per file, so that the only draws after the seed are halfway_sampling's randn_like per step
        torch.manual_seed(SEED_NOISE)
        return torch.from_numpy(data.astype(np.float32) / 32768.0)[None], sr

    def save(path, wav, sr):
        saved[path] = (wav.clone(), sr)

    torchaudio.load, torchaudio.save = load, save
    sys.path.insert(0, REFERENCE_ROOT)
    import srcs.sample as ref_sample                                   # the reference's script module
    tmp = tempfile.mkdtemp()
    args = sample_args(**FLAGS, model_path=os.path.join(tmp, "ladiff.amlt"), model_for_cond=os.path.join(tmp, "cond.amlt"))
    torch.save(make_state_dict(seed=SEED_MODEL, **ladiff_model_kwargs(args)), args.model_path)
    torch.save(make_state_dict(seed=SEED_COND, **cond_model_kwargs(args)), args.model_for_cond)
    out = {}
    for rel in FILES:
        one = tempfile.mkdtemp()
        os.makedirs(os.path.dirname(os.path.join(one, rel)), exist_ok=True)
        shutil.copy(os.path.join(cli, "in", rel), os.path.join(one, rel))
        ns = argparse.Namespace(**{**vars(args), "input_dir": one, "output_dir": os.path.join(tmp, "out")})
        saved.clear()
        ref_sample.synthesis(ns)
        assert len(saved) == 1, list(saved)
        (path, (wav, sr)), = saved.items()
        assert sr == 16000 and path.endswith(rel), (path, sr)
        out[rel] = wav
        print(rel, tuple(wav.shape), float(wav.abs().max()), flush=True)
    torch.save(dict(files=FILES, flags=FLAGS, seeds=dict(model=SEED_MODEL, cond=SEED_COND, noise=SEED_NOISE), midway_t=100, wav_hat=out),
               os.path.join(cli, "reference_outputs.pt"))
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
