"""BASELINE config 1 (plumbing) on the GPU: `ladiffcodec_b200.sample.synthesis` — the mirror of the reference's script entry
srcs/sample.py:50-136 — over the same wav FILES, checkpoint FILES and flags the real reference's own `srcs.sample.synthesis` was
run on (tests/golden/make_golden_cli.py, CPU, midway_t = 100), compared with the waveforms the reference saved."""
import argparse
import os
import shutil
import subprocess
import sys

import pytest
import torch

import parity_common as pc

pytestmark = pytest.mark.gpu
CLI = os.path.join(pc.GOLDEN_DIR, "cli")


def _setup(tmp_path):
    from ladiffcodec_b200.config import sample_args
    fx = torch.load(os.path.join(CLI, "reference_outputs.pt"), weights_only=False)
    args = sample_args(**fx["flags"], model_path=str(tmp_path / "ladiff.amlt"), model_for_cond=str(tmp_path / "cond.amlt"))
    torch.save(pc.make_state_dict(seed=fx["seeds"]["model"], **pc.ladiff_model_kwargs(args)), args.model_path)
    torch.save(pc.make_state_dict(seed=fx["seeds"]["cond"], **pc.cond_model_kwargs(args)), args.model_for_cond)
    return fx, args


def test_synthesis_on_wav_files_matches_the_reference_script(tmp_path):
    from scipy.io import wavfile
    from ladiffcodec_b200.sample import synthesis
    fx, args = _setup(tmp_path)
    for rel, ref in fx["wav_hat"].items():
        one = tmp_path / ("in_" + rel.replace("/", "_"))
        os.makedirs(os.path.dirname(one / rel), exist_ok=True)
        shutil.copy(os.path.join(CLI, "in", rel), one / rel)
        out_dir = str(tmp_path / "out")
        ns = argparse.Namespace(**{**vars(args), "input_dir": str(one), "output_dir": out_dir})
        synthesis(ns, noise_device="cpu", noise_seed=fx["seeds"]["noise"])    # the reference ran on CPU: same generator, same draw order
        path = out_dir + "/" + rel                           # sample.py:75-76: output_dir + path below input_dir
        assert os.path.exists(path), path
        sr, data = wavfile.read(path)
        got = torch.from_numpy(data.copy()).reshape(1, -1)
        assert sr == 16000 and got.dtype == torch.float32 and got.shape == ref.shape
        snr = pc.snr_db(got, ref)
        pc.record("cli_file_" + rel, n_steps=fx["midway_t"], wav_snr_db_vs_reference_script=snr, samples=int(got.shape[-1]))
        print(rel, tuple(got.shape), f"{snr:.1f} dB")
        assert snr >= pc.TOL["wav_snr_db_long"]
        assert abs(got.abs().max().item() - 1.0) < 1e-5       # sample.py:134


def test_module_entry_point_runs(tmp_path):
    """`python -m ladiffcodec_b200.sample <the reference's flags>` end to end (device noise: no sample-level comparison)."""
    fx, args = _setup(tmp_path)
    out_dir = str(tmp_path / "out2")
    cmd = [sys.executable, "-m", "ladiffcodec_b200.sample", "--model_for_cond", args.model_for_cond, "--model_path", args.model_path,
           "--run_diff", "--scaling_global", "--cond_bandwidth", "3", "--unet_scale_cond", "--input_dir", os.path.join(CLI, "in"),
           "--output_dir", out_dir]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=pc.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    from scipy.io import wavfile
    for rel, ref in fx["wav_hat"].items():
        sr, data = wavfile.read(out_dir + "/" + rel)
        assert sr == 16000 and data.shape[-1] == ref.shape[-1] and bool(torch.isfinite(torch.from_numpy(data.copy())).all())
