"""TEST INFRASTRUCTURE ONLY — imports the *real* reference (haiciyang/LaDiffCodec) from
/root/reference on CPU so that the oracle restatement (oracle/ladiff_oracle.py) and the
golden fixtures (tests/golden/) can be pinned against it.

The reference does not import as shipped: third-party modules are
missing from this image (pesq, matplotlib, librosa, asteroid, labml_*) and two of its own
files are absent (srcs/modules/transformer_discrete.py, srcs/losses/discrete_diff.py).
We register inert stand-ins in ``sys.modules`` *before* the import; nothing on the
sampling path touches them.

/root/reference exists only in the build container, never on the GPU box: callers must
gate on ``reference_available()``.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LADIFF_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "srcs"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch

    if "pesq" not in sys.modules:
        _mod("pesq", pesq=lambda *a, **k: 0.0)                       # sample.py:11
    if "matplotlib" not in sys.modules:
        plt = _mod("matplotlib.pyplot")
        _mod("matplotlib", pyplot=plt)                               # sample.py:15, utils.py:9
    if "librosa" not in sys.modules:
        _mod("librosa")                                              # dataset_libri.py:5

    class _SDR(torch.nn.Module):                                     # losses_fn.py:15,60
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            raise RuntimeError("asteroid stub: training-only loss, not on the sampling path")

    if "asteroid" not in sys.modules:
        sdr = _mod("asteroid.losses.sdr", MultiSrcNegSDR=_SDR)
        losses = _mod("asteroid.losses", sdr=sdr)
        _mod("asteroid", losses=losses)
    if "labml_helpers" not in sys.modules:
        m = _mod("labml_helpers.module", Module=torch.nn.Module)     # unet2d.py:30
        _mod("labml_helpers", module=m)
    if "labml_nn" not in sys.modules:
        u = _mod("labml_nn.diffusion.ddpm.utils", gather=lambda *a, **k: None)  # ddpm_loss_lab.py:169
        d = _mod("labml_nn.diffusion.ddpm", utils=u)
        df = _mod("labml_nn.diffusion", ddpm=d)
        _mod("labml_nn", diffusion=df)
    # files the reference imports but does not ship (modules/__init__.py:28, losses/__init__.py:14)
    if "srcs.modules.transformer_discrete" not in sys.modules:
        _mod("srcs.modules.transformer_discrete", Transformer=type("Transformer", (), {}))
    if "srcs.losses.discrete_diff" not in sys.modules:
        _mod("srcs.losses.discrete_diff", AbsorbingDiffusion=type("AbsorbingDiffusion", (), {}))


def import_reference():
    """Returns the reference's ``srcs.model`` module (DiffAudioRep lives there)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import srcs.model as ref_model  # noqa: E402
    if not getattr(ref_model, "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("`srcs` resolved to something other than the reference: " + str(ref_model.__file__))
    return ref_model


def build_reference_models(cfg):
    """Builds the two DiffAudioRep objects exactly as sample.py:52-65 does.

    cfg: dict with the sample.py argparse names (see ladiffcodec_b200.config.SampleConfig).
    Returns (ladiff_model, cond_model), both .eval(), on CPU.
    """
    ref = import_reference()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):   # model.py:62 prints the bandwidth
        model = ref.DiffAudioRep(other_cond=True, **cfg)
        cond = ref.DiffAudioRep(rep_dims=cfg["rep_dims"], emb_dims=cfg["emb_dims"],
                                n_residual_layers=cfg["n_residual_layers"], n_filters=cfg["n_filters"],
                                lstm=cfg["lstm"], quantization=True, bandwidth=cfg["cond_bandwidth"],
                                ratios=cfg["cond_enc_ratios"], final_activation=cfg["final_activation"])
    return model.eval(), cond.eval()
