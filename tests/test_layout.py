"""CPU: checkpoint layout (host logic).  The key/shape lists are pinned to hashes of the real
reference's ``state_dict()`` taken by tests/golden/make_golden.py; when /root/reference is
present (build container) the live reference is checked too."""
import hashlib
import json
import os

import pytest
import torch

from conftest import load_golden, GOLDEN_DIR
from ladiffcodec_b200.config import sample_args, readme_args, num_quantizers, num_quantizers_at_call, build_parser
from ladiffcodec_b200.layout import state_dict_spec, ladiff_model_kwargs, cond_model_kwargs
from ladiffcodec_b200.synthetic import make_state_dict


def _hash(spec):
    return hashlib.sha256("\n".join(f"{k}:{tuple(v)}" for k, v in spec.items()).encode()).hexdigest()


@pytest.mark.parametrize("name", ["A_3kbps", "B_3kbps", "A_1p5kbps"])
def test_spec_matches_reference_key_hash(name):
    fx = load_golden(name)
    args = sample_args(**fx["flags"])
    sm = state_dict_spec(**ladiff_model_kwargs(args))
    sc = state_dict_spec(**cond_model_kwargs(args))
    assert len(sm) == fx["n_keys_model"] and _hash(sm) == fx["keys_model_sha"]
    assert len(sc) == fx["n_keys_cond"] and _hash(sc) == fx["keys_cond_sha"]


def test_cond_codec_keys_in_clear():
    ref = json.load(open(os.path.join(GOLDEN_DIR, "cond_codec_keys.json")))
    spec = state_dict_spec(**cond_model_kwargs(readme_args()))
    assert list(ref.keys()) == list(spec.keys())
    assert all(tuple(ref[k]) == tuple(spec[k]) for k in ref)
    assert len(spec) == 148


def test_ladiff_checkpoint_has_745_keys_and_aliases():
    args = readme_args()
    sd = make_state_dict(seed=3, **ladiff_model_kwargs(args))
    assert len(sd) == 745
    assert sum(k.startswith("diffusion.model.") for k in sd) == 340
    for k in sd:
        if k.startswith("diffusion.model."):
            assert sd[k].data_ptr() == sd["diff_model." + k[16:]].data_ptr()
    assert all(v.dtype == torch.float32 for v in sd.values())
    # deterministic in the seed, independent of key order
    sd2 = make_state_dict(seed=3, **ladiff_model_kwargs(args))
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)


def test_num_quantizers():
    assert num_quantizers(3.0, 320) == 6 and num_quantizers(1.5, 320) == 3     # model.py:65
    assert num_quantizers_at_call(3.0, 50.0, 6) == 6 and num_quantizers_at_call(1.5, 50.0, 3) == 3  # vq.py:86-98


def test_cli_defaults_match_sample_py():
    a = build_parser().parse_args([])
    assert a.enc_ratios == [8] and a.upsampling_ratios == [5, 4, 2] and a.diff_dims == 256
    assert a.unet_scale_cond is False and a.model_type == "unet" and a.cond_bandwidth == 3.0
    b = build_parser().parse_args("--run_diff --scaling_global --cond_bandwidth 1.5 --unet_scale_cond "
                                  "--enc_ratios 8 4 --upsampling_ratios 5 2".split())
    assert b.enc_ratios == [8, 4] and b.upsampling_ratios == [5, 2] and b.unet_scale_cond and b.cond_bandwidth == 1.5


@pytest.mark.skipif(not os.path.isdir("/root/reference/srcs"), reason="reference tree only exists in the build container")
def test_spec_matches_live_reference():
    from oracle.ref_import import build_reference_models
    args = readme_args()
    m, c = build_reference_models(vars(args))
    sm = state_dict_spec(**ladiff_model_kwargs(args))
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(v)) for k, v in sm.items()]
    sc = state_dict_spec(**cond_model_kwargs(args))
    assert [(k, tuple(v.shape)) for k, v in c.state_dict().items()] == [(k, tuple(v)) for k, v in sc.items()]
