"""ladiffcodec_b200 — B200-native (sm_100a) sampling path of LaDiffCodec behind the reference's own API.

    from ladiffcodec_b200 import DiffAudioRep, load_model, synthesize
"""
from .config import sample_args, readme_args, build_parser          # noqa: F401
from .utils import load_model                                        # noqa: F401


def __getattr__(name):   # model/sample import torch + the CUDA library lazily
    if name in ("DiffAudioRep", "DiffAudioTime", "GaussianDiffusion1D", "Unet1D"):
        from . import model
        return getattr(model, name)
    if name in ("synthesize", "synthesis", "build_models"):
        from . import sample
        return getattr(sample, name)
    raise AttributeError(name)
