"""CPU: the arithmetic of the tensor-core leaf (csrc/leaf_tc.cu) restated with numpy and checked against the fp64 oracle.

The kernels evaluate the Gaussian leaf + product as a quadratic form on operands split into two tf32 words (hi = upper 19
bits of the fp32 value, lo = the remainder; products hi*hi + hi*lo + lo*hi accumulated in fp32).  This file pins that
FORMULATION -- operand order, the three coefficients, NaN handling, the gradient identities of the backward epilogue --
and its error bound without a GPU; tests/test_gpu_leaf_tc.py checks the kernels themselves."""
import numpy as np
import pytest
import torch

from oracle import spn_oracle as O

LOG_SQRT_2PI = 0.9189385332046727


def tf32_hi(v):
    return (v.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(v):
    v = v.astype(np.float32)
    hi = tf32_hi(v)
    return hi, tf32_hi((v - hi).astype(np.float32))   # the hardware reads the upper 19 bits of the remainder word


def leaf_tc_forward(x, mu, sigma):
    """x [B, card] (NaN = marginalised), mu / sigma [card, I] -> [B, I]; K entries per feature as in leaf_tc.cu:
    A (rows):   x2h x2h x2l xh xh xl v v        B (params): ah al ah bh bl bh ch cl"""
    f32 = np.float32
    ok = ~np.isnan(x)
    xz = np.where(ok, x, 0).astype(f32)
    v = ok.astype(f32)
    av = (f32(-1) / (f32(2) * sigma * sigma)).astype(f32)
    bv = (f32(-2) * av * mu).astype(f32)
    cv = (av * mu * mu - np.log(sigma).astype(f32) - f32(LOG_SQRT_2PI)).astype(f32)
    (ah, al), (bh, bl), (ch, cl) = split(av), split(bv), split(cv)
    (xh, xl), (qh, ql) = split(xz), split((xz * xz).astype(f32))
    acc = np.zeros((x.shape[0], mu.shape[1]), f32)
    for j in range(x.shape[1]):        # one K = 8 MMA step per feature, fp32 accumulation
        for a_col, b_row in ((qh, ah), (qh, al), (ql, ah), (xh, bh), (xh, bl), (xl, bh), (v, ch), (v, cl)):
            acc = (acc + (a_col[:, j, None] * b_row[None, j]).astype(f32)).astype(f32)
    return acc


def leaf_tc_backward(x, g, mu, sigma):
    """Parameter gradients from D = [x2h x2l xh xl v]^T [g_hi | g_lo] (fp32 accumulation over the rows) and the epilogue
    identities dmu = (sum g x - mu sum g) / sigma^2, dsigma = (sum g x^2 - 2 mu sum g x + mu^2 sum g) / sigma^3 - sum g / sigma."""
    f32 = np.float32
    ok = ~np.isnan(x)
    xz = np.where(ok, x, 0).astype(f32)
    (xh, xl), (qh, ql) = split(xz), split((xz * xz).astype(f32))
    v = ok.astype(f32)
    gh, gl = split(g)

    def dot(a_rows, b_cols):   # [B, card]^T [B, I] -> [card, I], row by row in fp32
        out = np.zeros((a_rows.shape[1], b_cols.shape[1]), f32)
        for b in range(a_rows.shape[0]):
            out = (out + (a_rows[b, :, None] * b_cols[b, None, :]).astype(f32)).astype(f32)
        return out

    sxx = dot(qh, gh) + dot(qh, gl) + dot(ql, gh)
    sx = dot(xh, gh) + dot(xh, gl) + dot(xl, gh)
    s0 = dot(v, gh) + dot(v, gl)
    inv = (f32(1) / sigma).astype(f32)
    s1 = sx - mu * s0
    s2 = sxx - f32(2) * mu * sx + mu * mu * s0
    return (s1 * inv * inv).astype(f32), (s2 * inv * inv * inv - s0 * inv).astype(f32)


@pytest.mark.parametrize("sig_min,x_scale,nan", [(0.05, 1.0, True), (0.0, 1.0, False), (0.001, 2.0, True)])
def test_split_tf32_quadratic_form_matches_the_oracle_leaf(sig_min, x_scale, nan):
    rng = np.random.default_rng(7)
    B, card, I = 96, 25, 16
    x = (rng.standard_normal((B, card)) * x_scale).astype(np.float32)
    if nan:
        x[3, ::3] = np.nan
        x[B - 1] = np.nan
    mu = rng.standard_normal((card, I)).astype(np.float32)
    sigma = (sig_min + (1 - sig_min) / (1 + np.exp(-rng.standard_normal((card, I))))).astype(np.float32)
    g = rng.standard_normal((B, I)).astype(np.float32)

    arch = O.Arch(F=card, D=0, R=1, S=2, I=I, C=1)       # one scope of `card` features, one repetition
    mu_t = torch.from_numpy(mu).double().reshape(1, card, I, 1).requires_grad_(True)
    sg_t = torch.from_numpy(sigma).double().reshape(1, card, I, 1).requires_grad_(True)
    xt = torch.from_numpy(x).double().reshape(B, 1, card, 1, 1)
    ref = O.leaf_product(O.leaf_log_prob(xt, mu_t, sg_t, arch), arch)          # [B, 1, 1, I, 1]
    (ref.reshape(B, I) * torch.from_numpy(g).double()).sum().backward()
    ref_np = ref.detach().reshape(B, I).numpy()

    got = leaf_tc_forward(x, mu, sigma)
    scale = np.abs(ref_np).max()
    err = np.abs(got - ref_np).max()
    assert err <= 4e-6 * scale + 1e-4, (err, scale)                             # the bound of the GPU test
    assert np.all(got[B - 1] == 0) if nan else True                             # fully marginalised row

    dmu, dsg = leaf_tc_backward(x, g, mu, sigma)
    for name, a, r in (("dmu", dmu, mu_t.grad.reshape(card, I).numpy()), ("dsigma", dsg, sg_t.grad.reshape(card, I).numpy())):
        e = np.abs(a - r).max()
        assert e <= 1e-4 * np.abs(r).max() + 1e-6, (name, e, np.abs(r).max())


def test_single_tf32_words_would_not_hold_the_cancellation():
    """Why the operands are split: with one tf32 word per operand the three terms no longer cancel (the reason the leaf stayed
    on the CUDA cores in round 1)."""
    rng = np.random.default_rng(1)
    B, card, I = 64, 25, 8
    x = rng.random((B, card)).astype(np.float32)
    mu = rng.standard_normal((card, I)).astype(np.float32)
    sigma = (0.05 + 0.95 * rng.random((card, I))).astype(np.float32)
    exact = (-(x[:, :, None].astype(np.float64) - mu[None]) ** 2 / (2 * sigma[None].astype(np.float64) ** 2)
             - np.log(sigma.astype(np.float64))[None] - LOG_SQRT_2PI).sum(1)
    av = -1 / (2 * sigma * sigma)
    single = (tf32_hi(x * x)[:, :, None] * tf32_hi(av)[None] + tf32_hi(x)[:, :, None] * tf32_hi(-2 * av * mu)[None]
              + tf32_hi(av * mu * mu - np.log(sigma) - LOG_SQRT_2PI)[None]).sum(1)
    err_single = np.abs(single - exact).max() / np.abs(exact).max()
    err_split = np.abs(leaf_tc_forward(x, mu, sigma) - exact).max() / np.abs(exact).max()
    assert err_single > 1e-4 > 4e-6 > err_split, (err_single, err_split)
