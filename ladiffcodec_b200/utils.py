"""``load_model`` — the checkpoint boundary of the sampling path (reference ``srcs/utils.py:98-108``)."""
import re
from collections import OrderedDict

import torch


def strip_module_prefix(state_dict):
    """utils.py:100-107: checkpoints saved from DistributedDataParallel carry 'module.' in their keys."""
    if not any("module" in k for k in state_dict):
        return state_dict
    out = OrderedDict()
    for k, v in state_dict.items():
        out[re.sub("module.", "", k) if "module" in k else k] = v
    return out


def load_model(model, model_path, strict=True):
    """torch.load(path) → strip 'module.' → model.load_state_dict(strict).  `model_path` may also be an
    in-memory state dict (tests)."""
    state_dict = model_path if isinstance(model_path, dict) else torch.load(model_path, map_location="cpu")
    model.load_state_dict(strip_module_prefix(state_dict), strict=strict)


def save_checkpoints(state_dict, path):
    """utils.py:91 — plain ``torch.save(state_dict)`` to ``*.amlt``."""
    torch.save(OrderedDict(state_dict), path)
