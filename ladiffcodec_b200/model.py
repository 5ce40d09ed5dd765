"""Host-side mirror of the reference's model surface for the sampling path.

``DiffAudioRep`` keeps the constructor signature, attribute layout and call conventions of
``srcs/model.py:32-238`` (``.encoder .decoder .quantizer .diff_model .diffusion``,
``.get_cond``), so ``srcs/sample.py:50-136`` runs unchanged on top of it — but every tensor op
is a call into libladiff_b200.so (hand-written sm_100a CUDA) through the C-ABI.  PyTorch is
used for device memory, streams and (optionally) the RNG only.

Out of scope and rejected loudly (``NotImplementedError``): training ``forward``, ``run_vae``,
``use_film``, ``self_condition``, ``qtz_condition``, ``unet_scale_x``, ``model_type != 'unet'``.
"""
import ctypes
import functools
import math
from collections import OrderedDict
from dataclasses import dataclass, field

import torch

from . import _lib
from .config import MODEL_DEFAULTS, NUM_TIMESTEPS, num_quantizers, num_quantizers_at_call
from .layout import state_dict_spec, model_cfg, SCHEDULE_BUFFERS

_NOISE_CHUNK_BYTES = 1 << 30


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _on_device(fn):
    """Runs a method with the model's device current (a model on cuda:1 must not launch on cuda:0's context); the library
    call then uses the current stream OF THAT DEVICE."""
    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        owner = getattr(self, "_m", self)
        if owner.device.type != "cuda":
            raise _lib.LadiffError("no CUDA device: ladiffcodec_b200 has no CPU fallback")
        with torch.cuda.device(owner.device):
            return fn(self, *a, **k)
    return wrapper


def _f32(t, device):
    if not torch.is_tensor(t):
        raise TypeError(f"expected a tensor, got {type(t)}")
    return t.to(device=device, dtype=torch.float32).contiguous()


@dataclass
class QuantizedResult:               # srcs/quantization/vq.py:19-25
    quantized: torch.Tensor
    codes: torch.Tensor
    bandwidth: torch.Tensor
    penalty: torch.Tensor = None
    metrics: dict = field(default_factory=dict)


class _Sub:
    def __init__(self, owner):
        self._m = owner

    def eval(self):
        return self

    def to(self, *a, **k):
        return self


class SEANetEncoder(_Sub):
    """model.encoder(x) — srcs/modules/seanet.py:66-154."""

    def __init__(self, owner):
        super().__init__(owner)
        self.ratios = list(reversed(owner.cfg["enc_ratios"]))
        self.hop_length = int(math.prod(self.ratios))
        self.dimension = owner.cfg["rep_dims"]

    @_on_device
    def __call__(self, x):
        m = self._m
        x = _f32(x, m.device)
        B, C, T = x.shape
        if C != 1:
            raise ValueError("SEANetEncoder expects [B,1,T]")
        z = torch.empty(B, self.dimension, T // self.hop_length, device=m.device)
        ws = m._workspace(B, T)
        _lib.check(m._lib.ladiff_encode(m._h, _ptr(x), B, T, _ptr(z), _ptr(ws), ws.numel(), _stream()), "encoder")
        return z

    forward = __call__


class SEANetDecoder(_Sub):
    """model.decoder(z) — srcs/modules/seanet.py:157-248."""

    def __init__(self, owner):
        super().__init__(owner)
        self.ratios = list(owner.cfg["enc_ratios"])
        self.hop_length = int(math.prod(self.ratios))
        self.dimension = owner.cfg["rep_dims"]

    @_on_device
    def __call__(self, z):
        m = self._m
        z = _f32(z, m.device)
        B, C, L = z.shape
        if C != self.dimension:
            raise ValueError(f"SEANetDecoder expects [B,{self.dimension},L]")
        T = L * self.hop_length
        wav = torch.empty(B, 1, T, device=m.device)
        ws = m._workspace(B, T)
        _lib.check(m._lib.ladiff_decode(m._h, _ptr(z), B, L, _ptr(wav), _ptr(ws), ws.numel(), _stream()), "decoder")
        return wav

    forward = __call__


class ResidualVectorQuantizer(_Sub):
    """model.quantizer — srcs/quantization/vq.py:28-113 (eval mode)."""

    def __init__(self, owner, n_q, bins=1024):
        super().__init__(owner)
        self.n_q, self.bins, self.dimension = n_q, bins, owner.cfg["rep_dims"]

    def get_bandwidth_per_quantizer(self, sample_rate):
        return math.log2(self.bins) * sample_rate / 1000

    def get_num_quantizers_for_bandwidth(self, sample_rate, bandwidth=None):
        return num_quantizers_at_call(bandwidth, sample_rate, self.n_q, self.bins)

    @_on_device
    def _run(self, x, n_q, want_q, want_codes):
        m = self._m
        x = _f32(x, m.device)
        B, D, F = x.shape
        q = torch.empty_like(x) if want_q else None
        codes = torch.empty(n_q, B, F, dtype=torch.int64, device=m.device) if want_codes else None
        _lib.check(m._lib.ladiff_rvq_encode(m._h, _ptr(x), n_q, B, F, _ptr(codes), _ptr(q), _stream()), "rvq_encode")
        return q, codes

    def __call__(self, x, sample_rate, bandwidth=None, n_q=None):
        bw_per_q = self.get_bandwidth_per_quantizer(sample_rate)
        n_q = self.get_num_quantizers_for_bandwidth(sample_rate, bandwidth) if n_q is None else n_q
        q, codes = self._run(x, n_q, True, True)
        bw = torch.tensor(n_q * bw_per_q).to(q)
        return QuantizedResult(q, codes, bw, penalty=torch.zeros((), device=q.device))

    forward = __call__

    def encode(self, x, sample_rate, bandwidth=None):
        n_q = self.get_num_quantizers_for_bandwidth(sample_rate, bandwidth)
        return self._run(x, n_q, False, True)[1]

    @_on_device
    def decode(self, codes):
        m = self._m
        codes = codes.to(device=m.device, dtype=torch.int64).contiguous()
        n_q, B, F = codes.shape
        out = torch.empty(B, self.dimension, F, device=m.device)
        _lib.check(m._lib.ladiff_rvq_decode(m._h, _ptr(codes), n_q, B, F, _ptr(out), _stream()), "rvq_decode")
        return out


class _UpsamplingLayer:
    """One entry of Unet1D.upsampling_layers — SConvTranspose1d(causal=False), unet.py:372-377."""

    def __init__(self, owner, index, ratio):
        self._m, self.index, self.ratio = owner, index, ratio

    @_on_device
    def __call__(self, x):
        m = self._m
        x = _f32(x, m.device)
        B, C, Lin = x.shape
        y = torch.empty(B, C, Lin * self.ratio, device=m.device)
        _lib.check(m._lib.ladiff_upsample_layer(m._h, self.index, _ptr(x), B, Lin, _ptr(y), _stream()), "upsampling_layer")
        return y


class Unet1D(_Sub):
    """model.diff_model — srcs/modules/unet.py:250-469 (other_cond=True, use_film=False)."""

    def __init__(self, owner):
        super().__init__(owner)
        c = owner.cfg
        self.channels = c["rep_dims"]
        self.self_condition = False
        self.use_film = False
        self.unet_scale_cond = c["unet_scale_cond"]
        self.unet_scale_x = False
        self.upsampling_ratios = c["upsampling_ratios"]
        self.upsampling_layers = [_UpsamplingLayer(owner, i, r) for i, r in enumerate(c["upsampling_ratios"] or [])]

    @_on_device
    def process_cond(self, x_cond):
        """unet.py:407-420."""
        for layer in self.upsampling_layers:
            x_cond = layer(x_cond)
        if self.unet_scale_cond:
            x_cond = x_cond.clone() if not self.upsampling_layers else x_cond
            B = x_cond.shape[0]
            _lib.check(self._m._lib.ladiff_normalize_clips(_ptr(x_cond), B, x_cond[0].numel(), 2, _stream()), "scaling")
        return x_cond

    @_on_device
    def __call__(self, x, time, x_cond=None):
        m = self._m
        if x_cond is None:
            raise NotImplementedError("Unet1D without a condition is not on the sampling path (other_cond=True)")
        x, x_cond = _f32(x, m.device), _f32(x_cond, m.device)
        time = time.to(device=m.device, dtype=torch.int64).contiguous()
        B, C, L = x.shape
        F = x_cond.shape[-1]
        eps = torch.empty_like(x)
        ws = m._workspace(B, L * m.decoder.hop_length)
        _lib.check(m._lib.ladiff_unet_forward(m._h, _ptr(x), _ptr(time), _ptr(x_cond), B, L, F, _ptr(eps), _ptr(ws),
                                              ws.numel(), _stream()), "unet_forward")
        return eps

    forward = __call__


class GaussianDiffusion1D(_Sub):
    """model.diffusion — srcs/losses/ddpm_loss.py:78-450: every sampler (p_sample, halfway_sampling, p_sample_loop, sample,
    ddim_sample, infilling), q_sample and the forward-only training loss (forward → p_losses).

    Randomness.  The reference draws from torch's global generator inside these methods.  Every method here takes the draws
    explicitly instead: `noise=` a pre-drawn tensor consumed in the reference's draw order (parity mode), "torch" → drawn here
    with torch.randn per step on the model's device (what the reference does when it runs on that device), or None → the
    in-kernel counter-based generator keyed by (seed, timestep, global clip, element) (throughput mode)."""

    def __init__(self, owner, seq_length, sampling_timesteps=None):
        super().__init__(owner)
        self.model = owner.diff_model
        self.channels = owner.diff_model.channels
        self.self_condition = False
        self.seq_length = seq_length
        self.num_timesteps = NUM_TIMESTEPS
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else NUM_TIMESTEPS
        self.is_ddim_sampling = False          # ddpm_loss.py:132
        self.ddim_sampling_eta = 0.0           # ddpm_loss.py:134
        self.objective = "pred_noise"
        self.loss_type = "l1"

    # ---- plumbing
    def _check_noise(self, noise, B, C, L):
        m = self._m
        noise = _f32(noise, m.device)
        if noise.dim() != 4 or tuple(noise.shape[1:]) != (B, C, L):
            raise ValueError(f"noise must be [n,{B},{C},{L}], got {tuple(noise.shape)}")
        return noise

    def _steps(self, x, condition, t_start, n_steps, noise, seed):
        """DDPM steps t_start-1 … t_start-n_steps on x, in place.  noise: None → in-kernel generator(seed); 'torch' →
        torch.randn_like per step (the draws the reference makes, ddpm_loss.py:249); tensor [n,B,C,L] → consumed in loop order."""
        m = self._m
        B, C, L = x.shape
        condition = _f32(condition, m.device)            # raw pointers below: fp32, contiguous [B,C,F]
        if not (x.is_contiguous() and x.dtype == torch.float32 and x.device == m.device):
            raise ValueError("x must be a contiguous fp32 tensor on the model's device (it is updated in place)")
        F = condition.shape[-1]
        ws = m._workspace(B, L * m.decoder.hop_length)

        def call(xx, nz, n_nz, t0, n):
            _lib.check(m._lib.ladiff_ddpm_steps(m._h, _ptr(xx), _ptr(condition), _ptr(nz), n_nz, ctypes.c_uint64(seed),
                                                t0, n, B, L, F, _ptr(ws), ws.numel(), _stream(m.device)), "ddpm_steps")

        if isinstance(noise, str):
            if noise != "torch":
                raise ValueError("noise must be None, 'torch' or a tensor")
            per = max(1, min(n_steps, _NOISE_CHUNK_BYTES // (x.numel() * 4)))
            buf = torch.empty(per, B, C, L, device=x.device)
            t = t_start
            while t > t_start - n_steps:
                n = min(per, t - (t_start - n_steps))
                draws = 0
                for i in range(n):
                    if t - 1 - i > 0:
                        buf[draws].copy_(torch.randn_like(x)); draws += 1
                call(x, buf, draws, t, n)
                t -= n
        elif noise is None:
            call(x, None, 0, t_start, n_steps)
        else:
            noise = self._check_noise(noise, B, C, L)
            call(x, noise, noise.shape[0], t_start, n_steps)
        return x

    def _initial(self, shape, init, seed, uniform=False):
        """The sampler's first draw (torch.randn(shape), ddpm_loss.py:256,277; torch.rand for infilling, :336)."""
        m = self._m
        if init is not None and not isinstance(init, str):
            x = _f32(init, m.device).clone()
            if tuple(x.shape) != tuple(shape):
                raise ValueError(f"init must have shape {tuple(shape)}")
            return x
        if init == "torch":
            return (torch.rand if uniform else torch.randn)(tuple(shape), device=m.device)
        x = torch.empty(tuple(shape), device=m.device)
        _lib.check(m._lib.ladiff_randn(m._h, _ptr(x), shape[0], x[0].numel(), ctypes.c_uint64(seed), int(uniform), _stream(m.device)), "randn")
        return x

    # ---- samplers
    @torch.no_grad()
    @_on_device
    def p_sample(self, x, t: int, condition=None, clip_denoised=True, noise="torch", seed=0):
        """ddpm_loss.py:244-251.  Returns (pred_img, None): x_start is only consumed by self-conditioning,
        which is off on this path."""
        if not clip_denoised:
            raise NotImplementedError("clip_denoised=False is not on the sampling path")
        m = self._m
        x = _f32(x, m.device).clone()
        condition = _f32(condition, m.device)
        return self._steps(x, condition, t + 1, 1, noise, seed), None

    @torch.no_grad()
    @_on_device
    def halfway_sampling(self, img=None, t=None, condition=None, noise="torch", seed=0):
        """ddpm_loss.py:370-385: steps i = t-1 … 0 of the 1000-step schedule starting from `img`."""
        m = self._m
        img, condition = _f32(img, m.device), _f32(condition, m.device)
        if img.shape == condition.shape:
            for layer in self.model.upsampling_layers:
                img = layer(img)
        else:
            img = img.clone()
        return self._steps(img, condition, int(t), int(t), noise, seed)

    @torch.no_grad()
    @_on_device
    def p_sample_loop(self, shape, condition=None, noise="torch", seed=0, init="torch", n_steps=None):
        """ddpm_loss.py:253-266: from N(0, I), all 1000 steps (`n_steps` < 1000 stops early: the first n_steps of the loop)."""
        m = self._m
        img = self._initial(shape, init if noise is not None or init != "torch" else None, seed)
        n = self.num_timesteps if n_steps is None else int(n_steps)
        return self._steps(img, _f32(condition, m.device), self.num_timesteps, n, noise, seed)

    @torch.no_grad()
    @_on_device
    def ddim_sample(self, shape, condition=None, clip_denoised=True, noise="torch", seed=0, init="torch"):
        """ddpm_loss.py:268-303: `sampling_timesteps` DDIM steps from N(0, I) with eta = ddim_sampling_eta."""
        if not clip_denoised:
            raise NotImplementedError("clip_denoised=False is not on the sampling path")
        m = self._m
        B, C, L = tuple(shape)
        condition = _f32(condition, m.device)
        # the reference's own host-side time grid (ddpm_loss.py:273-275); plain Python integers, no tensor work
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        n_pairs = len(times) - 1
        img = self._initial(shape, init if noise is not None or init != "torch" else None, seed)
        if isinstance(noise, str):
            if noise != "torch":
                raise ValueError("noise must be None, 'torch' or a tensor")
            n_draw = sum(1 for tn in times[1:] if tn >= 0)
            noise = torch.stack([torch.randn_like(img) for _ in range(n_draw)]) if n_draw else torch.empty(0, B, C, L, device=m.device)
        n_noise = 0
        if noise is not None:
            noise = self._check_noise(noise, B, C, L) if noise.numel() else noise
            n_noise = noise.shape[0]
        F = condition.shape[-1]
        ws = m._workspace(B, L * m.decoder.hop_length)
        arr = (ctypes.c_int32 * len(times))(*times)
        _lib.check(m._lib.ladiff_ddim_steps(m._h, _ptr(img), _ptr(condition), arr, n_pairs, float(self.ddim_sampling_eta),
                                            _ptr(noise) if n_noise else None, n_noise, ctypes.c_uint64(seed), B, L, F, _ptr(ws), ws.numel(),
                                            _stream(m.device)), "ddim_steps")
        return img

    @torch.no_grad()
    def sample(self, batch_size=16, condition=None, noise="torch", seed=0, init="torch"):
        """ddpm_loss.py:305-309."""
        sample_fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return sample_fn((batch_size, self.channels, self.seq_length), condition, noise=noise, seed=seed, init=init)

    @torch.no_grad()
    @_on_device
    def infilling(self, infill_img, condition, midway_t=None, noise=None, offset=0, lam=0.8, step_noise="torch", seed=0, init="torch"):
        """ddpm_loss.py:331-367.  `noise` and `offset` are accepted and unused, as in the reference (its `noise` only suppresses
        an unused draw).  Per step t = midway_t-1 … 0: img ← p_sample(img); img ← (1-lam) img + lam infill; infill ← p_sample(infill);
        img ← (1-lam) img + lam infill.  step_noise: "torch" | None | tensor [2(midway_t-1), B, C, L] in the reference's draw order."""
        m = self._m
        condition = _f32(condition, m.device)
        infill = _f32(infill_img, m.device).clone()
        B, C, L = condition.shape[0], self.channels, self.seq_length
        if tuple(infill.shape) != (B, C, L):
            raise ValueError(f"infill_img must be [{B},{C},{L}] (batch, channels, seq_length), got {tuple(infill.shape)}")
        img = self._initial((B, C, L), init if step_noise is not None or init != "torch" else None, seed, uniform=True)
        pre = None
        if step_noise is not None and not isinstance(step_noise, str):
            pre = self._check_noise(step_noise, B, C, L)
        k = 0
        n = int(B) * C * L

        def mix():
            _lib.check(m._lib.ladiff_axpby(_ptr(img), 1 - lam, _ptr(infill), lam, n, _stream(m.device)), "axpby")

        for t in reversed(range(0, int(midway_t))):
            for which, xx in ((0, img), (1, infill)):
                if pre is not None:
                    nz = pre[k:k + 1] if t > 0 else None
                    k += 1 if t > 0 else 0
                    self._steps(xx, condition, t + 1, 1, nz if nz is not None else torch.zeros(0, B, C, L, device=m.device), seed)
                else:
                    # two draws per timestep: the second p_sample of a step uses a different key
                    self._steps(xx, condition, t + 1, 1, step_noise, seed if which == 0 else seed ^ 0x9E3779B97F4A7C15)
                mix()
        return img

    @torch.no_grad()
    def interpolate(self, *a, **k):
        raise NotImplementedError("interpolate (ddpm_loss.py:311-328) calls p_sample without a condition, which the reference's own "
                                  "conditional UNet (other_cond=True) cannot evaluate (unet.py:428): not usable with the released models")

    @torch.no_grad()
    @_on_device
    def q_sample(self, x_start, t, noise=None):
        """ddpm_loss.py:387-393."""
        m = self._m
        x_start = _f32(x_start, m.device)
        noise = torch.randn_like(x_start) if noise is None else _f32(noise, m.device)
        t = t.to(device=m.device, dtype=torch.int64).contiguous()
        out = torch.empty_like(x_start)
        ws = m._workspace(x_start.shape[0], m.decoder.hop_length * 16)
        _lib.check(m._lib.ladiff_q_sample(m._h, _ptr(x_start), _ptr(t), _ptr(noise), _ptr(out), x_start.shape[0], x_start[0].numel(),
                                          _ptr(ws), ws.numel(), _stream(m.device)), "q_sample")
        return out

    @torch.no_grad()
    @_on_device
    def p_losses(self, x_start, t, cond=None, noise=None, return_model_out=False):
        """ddpm_loss.py:404-437, forward only (no autograd graph): → (loss, predicted_x_start, x_t)."""
        m = self._m
        x_start, cond = _f32(x_start, m.device), _f32(cond, m.device)
        noise = torch.randn_like(x_start) if noise is None else _f32(noise, m.device)
        t = t.to(device=m.device, dtype=torch.int64).contiguous()
        B, C, L = x_start.shape
        F = cond.shape[-1]
        loss = torch.empty(1, device=m.device)
        pred, x_t = torch.empty_like(x_start), torch.empty_like(x_start)
        mo = torch.empty_like(x_start) if return_model_out else None
        ws = m._workspace(B, L * m.decoder.hop_length)
        _lib.check(m._lib.ladiff_p_losses(m._h, _ptr(x_start), _ptr(t), _ptr(cond), _ptr(noise), B, L, F, _ptr(loss), _ptr(pred), _ptr(x_t),
                                          _ptr(mo), _ptr(ws), ws.numel(), _stream(m.device)), "p_losses")
        out = (loss[0], pred, x_t)
        return out + (mo,) if return_model_out else out

    def forward(self, x, cond=None, t=None, *args, **kwargs):
        """ddpm_loss.py:439-450 → (loss, predicted_x_start, x_t, t)."""
        b, c, n = x.shape
        assert n == self.seq_length, f"seq length must be {self.seq_length}, now is {n}"
        if t is None:
            t = torch.randint(0, self.num_timesteps, (b,), device=self._m.device).long()
        return (*self.p_losses(x, t, cond, *args, **kwargs), t)

    __call__ = forward


class DiffAudioRep:
    """Drop-in for ``srcs.model.DiffAudioRep`` on the sampling path (same keyword arguments, model.py:34)."""

    def __init__(self, **kwargs):
        cfg = model_cfg(**kwargs)
        self.cfg = cfg
        for flag in ("use_film", "self_condition", "qtz_condition", "unet_scale_x", "run_vae"):
            if cfg[flag]:
                raise NotImplementedError(f"{flag}=True is outside the sampling path (SURVEY.md §2.1)")
        if cfg["run_diff"] and cfg["model_type"] != "unet":
            raise NotImplementedError("model_type must be 'unet' (model.py:73); transformer/unet2d are out of scope")
        if cfg["run_diff"] and not cfg["other_cond"]:
            raise NotImplementedError("run_diff without other_cond (unconditional / self-conditioned UNet) is out of scope")
        if cfg["norm"] != "weight_norm" or not cfg["causal"] or cfg["n_residual_layers"] != 1:
            raise NotImplementedError("only the reference defaults norm='weight_norm', causal=True, n_residual_layers=1")
        if cfg["final_activation"] is not None:
            raise NotImplementedError("final_activation is not used by any released configuration")
        self.quantization = cfg["quantization"]
        self.bandwidth = cfg["bandwidth"]
        self.sample_rate = cfg["sample_rate"]
        self.run_diff = cfg["run_diff"]
        self.model_type = cfg["model_type"]
        self.scaling_global = cfg["scaling_global"]
        self.training = False
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        hop = int(math.prod(cfg["enc_ratios"]))
        n_q = n_q_used = 0
        if self.quantization:
            self.frame_rate = self.sample_rate / hop                              # model.py:64
            n_q = num_quantizers(self.bandwidth, hop, self.sample_rate)           # model.py:65
            n_q_used = num_quantizers_at_call(self.bandwidth, self.frame_rate, n_q)
        self._lib = _lib.get_lib()
        ccfg = _lib.make_config(rep_dims=cfg["rep_dims"], diff_dims=cfg["diff_dims"], n_filters=cfg["n_filters"],
                                lstm=cfg["lstm"], enc_ratios=cfg["enc_ratios"], quantization=self.quantization, n_q=n_q,
                                n_q_used=n_q_used, run_diff=self.run_diff, cond_channels=cfg["cond_channels"],
                                upsampling_ratios=cfg["upsampling_ratios"] if cfg["other_cond"] else None,
                                unet_scale_cond=cfg["unet_scale_cond"], sample_rate=self.sample_rate)
        h = ctypes.c_void_p()
        _lib.check(self._lib.ladiff_create(ctypes.byref(ccfg), ctypes.byref(h)), "ladiff_create")
        self._h = h
        self._ws = None
        self._loaded = False
        self._buffers = {}
        self.encoder = SEANetEncoder(self)
        self.decoder = SEANetDecoder(self)
        if self.quantization:
            self.quantizer = ResidualVectorQuantizer(self, n_q)
        if self.run_diff:
            self.diff_model = Unet1D(self)
            self.diffusion = GaussianDiffusion1D(self, cfg["seq_length"], cfg["sampling_timesteps"])

    # ---- nn.Module-like surface used by sample.py
    def to(self, device=None, *a, **k):
        if device is not None:
            device = torch.device(device)
            if device.type != "cuda":
                raise _lib.LadiffError("ladiffcodec_b200 runs on CUDA (sm_100a) only; there is no CPU path")
            if self._loaded and device != self.device:
                raise _lib.LadiffError("move the model before loading weights")
            self.device = device if device.index is not None else torch.device("cuda", torch.cuda.current_device())
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def eval(self):
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("training is outside the sampling path")
        return self

    def expected_keys(self):
        """(name, shape) pairs the strict loader expects, as enumerated by the C library."""
        n = self._lib.ladiff_expected_keys(self._h)
        out = []
        for i in range(n):
            name, shp, nd = ctypes.c_char_p(), (ctypes.c_int64 * 4)(), ctypes.c_int32()
            _lib.check(self._lib.ladiff_expected_key_at(self._h, i, ctypes.byref(name), shp, ctypes.byref(nd)))
            out.append((name.value.decode(), tuple(shp[d] for d in range(nd.value))))
        return out

    def load_state_dict(self, state_dict, strict=True):
        """nn.Module.load_state_dict(strict) as used by utils.load_model (utils.py:98-108)."""
        if self._loaded:
            raise _lib.LadiffError("weights already loaded into this model")
        if not torch.cuda.is_available():
            raise _lib.LadiffError("no CUDA device: ladiffcodec_b200 has no CPU fallback")
        spec = OrderedDict(self.expected_keys())
        missing = [k for k in spec if k not in state_dict]
        unexpected = [k for k in state_dict if k not in spec]
        if missing or (strict and unexpected):
            msgs = []
            if unexpected and strict:
                msgs.append("Unexpected key(s) in state_dict: " + ", ".join(f'"{k}"' for k in unexpected[:8]) + (" …" if len(unexpected) > 8 else ""))
            if missing:
                msgs.append("Missing key(s) in state_dict: " + ", ".join(f'"{k}"' for k in missing[:8]) + (" …" if len(missing) > 8 else ""))
            raise RuntimeError("Error(s) in loading state_dict for DiffAudioRep:\n\t" + "\n\t".join(msgs))
        with torch.cuda.device(self.device):
            for name, shape in spec.items():
                t = state_dict[name]
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError(f"Error(s) in loading state_dict for DiffAudioRep:\n\tsize mismatch for {name}: "
                                       f"copying a param with shape {tuple(t.shape)}, the shape in current model is {tuple(shape)}.")
                shp = (ctypes.c_int64 * 4)(*shape) if shape else (ctypes.c_int64 * 4)()
                if name.startswith("diffusion.model."):
                    twin = state_dict["diff_model." + name[len("diffusion.model."):]]
                    if t.data_ptr() != twin.data_ptr() and not torch.equal(t, twin):
                        raise RuntimeError(f"{name} differs from its alias diff_model.* (the reference stores one storage twice)")
                    _lib.check(self._lib.ladiff_load_weight(self._h, name.encode(), None, shp, len(shape)), name)
                    continue
                tt = t.detach().to(torch.float32).contiguous()
                _lib.check(self._lib.ladiff_load_weight(self._h, name.encode(), _ptr(tt), shp, len(shape)), name)
                if name.startswith("diffusion.") and name.split(".")[1] in SCHEDULE_BUFFERS:
                    self._buffers[name.split(".")[1]] = tt.to(self.device)
            _lib.check(self._lib.ladiff_finalize(self._h), "ladiff_finalize")
        if self.run_diff:
            for k, v in self._buffers.items():
                setattr(self.diffusion, k, v)
        self._loaded = True
        return torch.nn.modules.module._IncompatibleKeys([], unexpected)

    def state_dict_spec(self):
        return state_dict_spec(**self.cfg)

    def _workspace(self, B, T, other=None):
        """Caller-owned scratch (a torch uint8 tensor) sized by the library; grown on demand."""
        if other is None:
            need = self._lib.ladiff_workspace_bytes(self._h, B, T)
        else:
            need = self._lib.ladiff_synthesize_workspace_bytes(self._h, other._h, B, T)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(int(need) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    @_on_device
    def get_cond(self, x, return_codes=False):
        """model.py:223-231: encoder → (if quantization) RVQ quantized output."""
        if not self._loaded:
            raise _lib.LadiffError("load weights first (utils.load_model)")
        x = _f32(x, self.device)
        B, C, T = x.shape
        F = T // self.encoder.hop_length
        cond = torch.empty(B, self.cfg["rep_dims"], F, device=self.device)
        codes = None
        if return_codes and self.quantization:
            codes = torch.empty(self._n_q_used(), B, F, dtype=torch.int64, device=self.device)
        ws = self._workspace(B, T)
        _lib.check(self._lib.ladiff_get_cond(self._h, _ptr(x), B, T, _ptr(cond), _ptr(codes), None, _ptr(ws), ws.numel(),
                                             _stream()), "get_cond")
        return (cond, codes) if return_codes else cond

    def _n_q_used(self):
        return num_quantizers_at_call(self.bandwidth, self.frame_rate, self.quantizer.n_q)

    def get_scale(self, x):
        """model.py:233-238 with scaling_global: the constant."""
        if not self.scaling_global:
            raise NotImplementedError("only scaling_global (model.py:137-139) is used by the released configurations")
        return 18.0

    @torch.no_grad()
    @_on_device
    def forward(self, x, t=None, cond=None, noise=None):
        """model.py:146-221, forward only (validation scoring; no autograd graph) for the configurations of the sampling path:
        run_diff with an external condition (`cond` from the conditioning codec) and global scaling, or a plain codec
        (run_diff=False).  Returns what the reference returns:
          run_diff:  ({'diff_loss', 'neg_loss'}, x_hat, x_rep, predicted_x_start, x_t, t, x_rep_qtz, scale)
          codec:     ({'tot_loss', 'qtz_loss', 'neg_sdr'}, x_hat)  /  ({'neg_sdr'}, x_hat) without a quantizer."""
        if not self._loaded:
            raise _lib.LadiffError("load weights first (utils.load_model)")
        x = _f32(x, self.device)
        B = x.shape[0]
        x_rep = self.encoder(x)
        x_rep_qtz, qtz_loss = None, None
        if self.quantization:
            res = self.quantizer(x_rep, sample_rate=self.frame_rate, bandwidth=self.bandwidth)
            x_rep_qtz, qtz_loss = res.quantized, res.penalty          # eval mode: commitment penalty is 0 (core_vq.py:299-303)

        def neg_sdr(a, b):                                            # sdr_loss(x, x_hat).mean(), model.py:198
            out = torch.empty(B, device=self.device)
            _lib.check(self._lib.ladiff_sdsdr(_ptr(a), _ptr(b), _ptr(out), B, a[0].numel(), -30.0, _stream(self.device)), "sdsdr")
            return out.mean()

        if self.run_diff:
            if cond is None:
                raise NotImplementedError("DiffAudioRep.forward without `cond` (unconditional / qtz_condition training) is out of scope")
            if not self.scaling_global or self.cfg.get("scaling_frame") or self.cfg.get("scaling_feature"):
                raise NotImplementedError("only scaling_global (model.py:137-139) is used by the released configurations")
            scale = 18.0
            x_rep = x_rep.clone()
            _lib.check(self._lib.ladiff_axpby(_ptr(x_rep), 1.0 / scale, None, 0.0, x_rep.numel(), _stream(self.device)), "scaling")
            self.diffusion.seq_length = x_rep.shape[-1] if self.diffusion.seq_length is None else self.diffusion.seq_length
            diff_loss, predicted_x_start, x_t, t = self.diffusion(x_rep, cond, t=t, noise=noise)
            in_dec = predicted_x_start.clone()
            _lib.check(self._lib.ladiff_axpby(_ptr(in_dec), scale, None, 0.0, in_dec.numel(), _stream(self.device)), "unscale")
            x_hat = self.decoder(in_dec)
            return ({"diff_loss": diff_loss, "neg_loss": neg_sdr(x, x_hat)}, x_hat, x_rep, predicted_x_start, x_t, t, x_rep_qtz, scale)
        x_hat = self.decoder(x_rep_qtz if self.quantization else x_rep)
        nl = neg_sdr(x, x_hat)
        if not self.quantization:
            return {"neg_sdr": nl}, x_hat
        return {"tot_loss": qtz_loss + nl, "qtz_loss": qtz_loss, "neg_sdr": nl}, x_hat

    __call__ = forward

    def take_launch_count(self):
        return int(self._lib.ladiff_take_launch_count(self._h))

    def set_conv_impl(self, impl):
        """0 = tcgen05 (default), 1 = SIMT check kernel (tests only)."""
        _lib.check(self._lib.ladiff_set_conv_impl(self._h, int(impl)))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ladiff_destroy(self._h)
                self._h = None
        except Exception:
            pass


class DiffAudioTime:
    def __init__(self, *a, **k):
        raise NotImplementedError("train_time_diff (DiffAudioTime, model.py:241) is outside the sampling path")
