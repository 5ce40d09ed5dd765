"""CPU: the C-ABI library loads, exports every symbol include/ladiff_b200.h declares, and its host-side
logic (handle creation, strict key layout, workspace sizing, error reporting) works without a GPU."""
import ctypes
import os
import re

import pytest
import torch

from ladiffcodec_b200 import _lib
from ladiffcodec_b200.config import readme_args, sample_args
from ladiffcodec_b200.layout import state_dict_spec, ladiff_model_kwargs, cond_model_kwargs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from ladiffcodec_b200 import build
        build.build()
    return _lib.get_lib()


def test_every_declared_symbol_is_exported_and_bound(lib):
    hdr = open(os.path.join(ROOT, "include", "ladiff_b200.h")).read()
    declared = set(re.findall(r"\b(ladiff_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.ladiff_abi_version() == _lib.ABI_VERSION


def test_library_has_no_torch_or_cuda_link_dependency():
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libcuda.so" not in out and "libcudart" not in out, out


@pytest.mark.parametrize("flags", [
    dict(run_diff=True, scaling_global=True, cond_bandwidth=3.0, unet_scale_cond=True, model_for_cond="c", model_path="m"),
    dict(run_diff=True, cond_bandwidth=1.5, enc_ratios=[8, 4], upsampling_ratios=[5, 2], model_for_cond="c", model_path="m"),
])
def test_c_side_key_layout_matches_python_spec(lib, flags):
    from ladiffcodec_b200.model import DiffAudioRep
    args = sample_args(**flags)
    for kw in (ladiff_model_kwargs(args), cond_model_kwargs(args)):
        m = DiffAudioRep(**kw)
        spec = state_dict_spec(**kw)
        assert m.expected_keys() == [(k, tuple(v)) for k, v in spec.items()]
        assert lib.ladiff_workspace_bytes(m._h, 2, 5120) > 0


def test_unsupported_flags_fail_loudly(lib):
    from ladiffcodec_b200.model import DiffAudioRep, DiffAudioTime
    base = ladiff_model_kwargs(readme_args())
    for bad in (dict(use_film=True), dict(self_condition=True), dict(qtz_condition=True), dict(unet_scale_x=True),
                dict(run_vae=True), dict(model_type="transformer"), dict(other_cond=False)):
        with pytest.raises(NotImplementedError):
            DiffAudioRep(**{**base, **bad})
    with pytest.raises(NotImplementedError):
        DiffAudioTime()
    m = DiffAudioRep(**base)
    with pytest.raises(_lib.LadiffError):            # forward-only training loss exists now, but not without a GPU / loaded weights
        m(torch.zeros(1, 1, 640))
    with pytest.raises(NotImplementedError):
        m.diffusion.interpolate(None, None)


def test_error_codes_and_messages(lib):
    h = ctypes.c_void_p()
    cfg = _lib.make_config(rep_dims=128, diff_dims=256, n_filters=32, lstm=2, enc_ratios=[8, 5, 4, 2], quantization=True, n_q=6,
                           n_q_used=6, run_diff=False, cond_channels=128, upsampling_ratios=None, unet_scale_cond=False,
                           sample_rate=16000)
    assert lib.ladiff_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    assert lib.ladiff_expected_keys(h) == 148
    shp = (ctypes.c_int64 * 4)(1, 0, 0, 0)
    rc = lib.ladiff_load_weight(h, b"not.a.key", None, shp, 1)
    assert rc == -3 and b"unexpected key" in lib.ladiff_last_error()
    rc = lib.ladiff_load_weight(h, b"encoder.model.0.conv.conv.bias", None, shp, 1)
    assert rc == -3 and b"size mismatch" in lib.ladiff_last_error()
    rc = lib.ladiff_finalize(h)
    assert rc == -3 and b"missing key" in lib.ladiff_last_error()
    rc = lib.ladiff_get_cond(h, None, 1, 640, None, None, None, None, 0, None)
    assert rc == -2                                                     # not finalized
    cfg.n_q_used = 9
    h2 = ctypes.c_void_p()
    assert lib.ladiff_create(ctypes.byref(cfg), ctypes.byref(h2)) == -1
    assert lib.ladiff_destroy(h) == 0


def test_no_cpu_fallback(lib):
    from ladiffcodec_b200.model import DiffAudioRep
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = DiffAudioRep(**cond_model_kwargs(readme_args()))
    with pytest.raises(_lib.LadiffError):
        m.load_state_dict({})
    with pytest.raises(_lib.LadiffError):
        m.to("cpu")


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "ladiffcodec_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f


def test_ddim_time_grid_equals_the_references(lib):
    """ladiff_ddim_times restates torch.linspace(-1, T-1, S+1).int() reversed (ddpm_loss.py:273-274) in C: every S in 1..1000."""
    for S in list(range(1, 200)) + [250, 333, 500, 999, 1000]:
        out = (ctypes.c_int32 * (S + 1))()
        assert lib.ladiff_ddim_times(1000, S, out) == 0
        ref = list(reversed(torch.linspace(-1, 999, steps=S + 1).int().tolist()))
        assert list(out) == ref, S


def test_sampler_surface_exists_and_refuses_what_the_reference_cannot_run():
    from ladiffcodec_b200.model import GaussianDiffusion1D
    for name in ("p_sample", "p_sample_loop", "sample", "ddim_sample", "infilling", "halfway_sampling", "q_sample", "p_losses", "forward"):
        assert callable(getattr(GaussianDiffusion1D, name)), name
    import inspect
    sig = inspect.signature(GaussianDiffusion1D.infilling)
    for arg in ("infill_img", "condition", "midway_t", "noise", "offset", "lam"):          # ddpm_loss.py:331
        assert arg in sig.parameters, arg


def test_bench_refuses_a_gpu_count_it_was_not_launched_with():
    import subprocess, sys
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT, env=env)
    assert r.returncode != 0
    assert ("--gpus 2 but WORLD_SIZE=1" in (r.stderr + r.stdout)) or ("no CUDA device" in (r.stderr + r.stdout))
