"""Per-launch CUDA-event timing of the tcgen05 conv kernel inside a UNet evaluation (bench.py roofline)."""
import ctypes

from . import _lib


def enable(model, on=True):
    """on: False/0 off, True/1 events around the tcgen05 conv launches, 2 events around every op of a UNet evaluation."""
    _lib.check(model._lib.ladiff_set_profiling(model._h, int(on)), "set_profiling")


def report(model):
    """dict(conv_ms, conv_flops, conv_launches, eval_ms) for the most recent profiled UNet evaluation, or None."""
    out = (ctypes.c_double * 4)()
    rc = model._lib.ladiff_profile_report(model._h, out)
    if rc != 0:
        return None
    return dict(conv_ms=out[0], conv_flops=out[1], conv_launches=int(out[2]), eval_ms=out[3])


def dump(model):
    """Per-launch table [(ms, gflop, label)] of the most recent profiled UNet evaluation."""
    buf = ctypes.create_string_buffer(1 << 17)
    if model._lib.ladiff_profile_dump(model._h, buf, len(buf)) != 0:
        return []
    rows = []
    for line in buf.value.decode().splitlines():
        ms, gf, label = line.split(" ", 2)
        rows.append((float(ms), float(gf), label))
    return rows
