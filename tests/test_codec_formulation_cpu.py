"""CPU checks of the algebra the codec conv kernels rely on (csrc/codec_ops.cu conv1d_f32_v2_kernel, csrc/codec_tc.cu):

  * a conv of stride S with K = KT*S taps == a stride-1 conv with KT taps over S "phase channels" per input channel, with the
    K-major weight layout conv_w_transpose_kernel produces;
  * SConvTranspose1d (conv.py:252-274) == a 2-tap stride-1 conv over s*Cout virtual channels with a phase-interleaved store;
  * the 3xTF32 split: v = hi + lo with round-to-nearest TF32 parts, x*w ~= lo*hi + hi*lo + hi*hi reproduces the fp32 product to
    ~2^-21, while a single TF32 product does not.

These are statements about index arithmetic and rounding, not about the CUDA code; the GPU parity tests check the kernels."""
import numpy as np
import torch
import torch.nn.functional as F


def _pad_reflect_causal(x, K, S):
    # SConv1d, causal: left padding (K - 1) - (S - 1), reflect (conv.py:217-232); the kernels use padL = (K - 1) - (S - 1)
    padL = (K - 1) - (S - 1)
    return F.pad(x, (padL, 0), mode="reflect"), padL


def test_strided_conv_is_a_stride1_conv_over_phase_channels():
    torch.manual_seed(0)
    for (Cin, Cout, K, S, L) in [(4, 6, 4, 2, 64), (3, 5, 8, 4, 96), (2, 4, 10, 5, 75), (4, 3, 16, 8, 128), (5, 7, 7, 1, 33)]:
        x = torch.randn(2, Cin, L, dtype=torch.float64)
        w = torch.randn(Cout, Cin, K, dtype=torch.float64)
        xp, padL = _pad_reflect_causal(x, K, S)
        y = F.conv1d(xp, w, stride=S)
        KT = K // S
        # wt[(ci*S + p)*KT + kt][co] = w[co][ci][kt*S + p]      (conv_w_transpose_kernel)
        wt = torch.zeros(Cin * S * KT, Cout, dtype=torch.float64)
        for ci in range(Cin):
            for p in range(S):
                for kt in range(KT):
                    wt[(ci * S + p) * KT + kt] = w[:, ci, kt * S + p]
        Lout = y.shape[-1]
        # xv[ci*S + p][u] = xpad[ci][u*S + p]   (u*S + p - padL in unpadded coordinates)
        U = Lout + KT - 1
        xv = torch.zeros(2, Cin * S, U, dtype=torch.float64)
        for ci in range(Cin):
            for p in range(S):
                idx = torch.arange(U) * S + p
                ok = idx < xp.shape[-1]
                xv[:, ci * S + p, ok] = xp[:, ci, idx[ok]]
        y2 = torch.zeros_like(y)
        for cv in range(Cin * S):
            for kt in range(KT):
                y2 += wt[cv * KT + kt][None, :, None] * xv[:, cv, None, kt:kt + Lout]
        assert torch.allclose(y, y2, atol=1e-12), (Cin, Cout, K, S)


def test_transposed_conv_is_a_two_tap_conv_over_phase_outputs():
    torch.manual_seed(1)
    for (Cin, Cout, s, L, causal) in [(4, 3, 2, 20, False), (3, 5, 4, 17, True), (2, 2, 8, 9, True), (3, 4, 5, 12, False)]:
        x = torch.randn(2, Cin, L, dtype=torch.float64)
        w = torch.randn(Cin, Cout, 2 * s, dtype=torch.float64)
        y_full = F.conv_transpose1d(x, w, stride=s)                       # length (L - 1) s + 2 s = (L + 1) s
        total = 2 * s - s                                                 # k - stride
        if causal:
            trim_l, trim_r = 0, total                                     # conv.py:267-270 (trim_right_ratio = 1)
        else:
            trim_r = total // 2
            trim_l = total - trim_r
        y = y_full[..., trim_l:y_full.shape[-1] - trim_r]
        assert y.shape[-1] == L * s
        # virtual channel (ph, co) at virtual position i (0..L): y_full[co, i*s + ph] = x[i-1] w[.., ph + s] + x[i] w[.., ph]
        xz = F.pad(x, (1, 1))                                             # zero extension: x[-1] = x[L] = 0
        y2 = torch.zeros(2, Cout, L * s, dtype=torch.float64)
        for ph in range(s):
            for i in range(L + 1):
                v = torch.einsum("bc,co->bo", xz[:, :, i], w[:, :, ph + s]) + torch.einsum("bc,co->bo", xz[:, :, i + 1], w[:, :, ph])
                pos = i * s + ph - trim_l                                 # il_trim in ConvF32Args
                if 0 <= pos < L * s:
                    y2[:, :, pos] = v
        assert torch.allclose(y, y2, atol=1e-12), (Cin, Cout, s, causal)


def _rna_tf32(a):
    """Round-to-nearest (ties away) to TF32's 10 explicit mantissa bits, as cvt.rna.tf32.f32 does."""
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def test_three_tf32_products_reproduce_the_fp32_product():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1 << 16).astype(np.float32)
    w = rng.standard_normal(1 << 16).astype(np.float32)
    exact = x.astype(np.float64) * w.astype(np.float64)
    xh, wh = _rna_tf32(x), _rna_tf32(w)
    xl, wl = _rna_tf32(x - xh), _rna_tf32(w - wh)
    # every partial product of two TF32 numbers is exact in fp32's 24-bit significand (11 + 11 bits)
    three = xl.astype(np.float64) * wh + xh.astype(np.float64) * wl + xh.astype(np.float64) * wh
    one = xh.astype(np.float64) * wh
    rel3 = np.abs(three - exact) / np.abs(exact)
    rel1 = np.abs(one - exact) / np.abs(exact)
    assert rel3.max() < 2.0 ** -20 and np.median(rel3) < 2.0 ** -23          # dropped lo*lo and the rounding of lo
    assert rel1.max() > 2.0 ** -12                                           # a single TF32 product: 10-bit operands
    # and a long dot product (K = 4096, the longest K loop of the codec) stays at fp32 level when the sum itself is exact
    K = 4096
    xs, ws = x[:K].astype(np.float64), w[:K].astype(np.float64)
    d_exact = float(np.dot(xs, ws))
    d3 = float(np.sum(three[:K]))
    assert abs(d3 - d_exact) < 1e-5 * np.sqrt(K)
