"""Autograd wrappers around the C-ABI kernels (include/spn_b200.h).

Each ``torch.autograd.Function`` here is a forward/backward kernel pair; PyTorch only provides device memory, the
current stream and the autograd tape.  All tensors must live on a CUDA device -- there is no other backend.
"""
import math

import torch

from . import _abi
from ._abi import call, ptr, require_cuda, stream_ptr


# Test hook (tests/test_gpu_ps.py): route per-conditional Sum layers through the general kernels instead of the streaming ones.
FORCE_GENERAL_SUM = False

# Bumped by everything that writes parameters behind autograd's version counters (parallel.FlatAdam.step,
# parallel.broadcast_parameters): part of the key of every derived-buffer cache (bf16 operand images, int32 permutations).
PARAM_EPOCH = 0


def bump_param_epoch():
    global PARAM_EPOCH
    PARAM_EPOCH += 1


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t.contiguous()


def _pow2_ceil(d: int) -> int:
    return 1 << max(0, (d - 1).bit_length())


# --------------------------------------------------------------------------------------------------------------
# Leaf: permutation gather + Gaussian log-density + tanh correction + NaN marginalisation + product over `card`
# --------------------------------------------------------------------------------------------------------------
# Leaf on the tensor cores (leaf_tc.cu: tf32 split operands) wherever the tensor-core path asks for the internal layout and
# the shape is supported; False keeps the exact fp32 kernels of leaf_tiled.cu (tests A/B the two).
LEAF_TENSOR_CORES = True


def leaf_tc_supported(I, R, card, d0):
    return bool(_abi.lib().spn_leaf_tc_supported(I, R, card, d0))


class LeafFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mu, sigma, perm32, card, tanh_corr, layout="ref"):
        require_cuda(x, mu, sigma, perm32)
        x, mu, sigma = _f32c(x), _f32c(mu), _f32c(sigma)
        Wsets, F, I, R = mu.shape
        *lead, w, Fx, oc, rr = x.shape
        assert Fx == F and oc in (1, I) and rr in (1, R)
        assert Wsets in (1, w), f"{Wsets} weight sets cannot be matched with {w} conditionals"
        B = x.numel() // (F * oc * rr)
        d0 = -(-F // card)
        sx = (F * oc * rr, oc * rr, rr if oc > 1 else 0, 1 if rr > 1 else 0)
        if layout == "ref":      # [*batch, w, d0, I, R]
            out = torch.empty(*lead, w, d0, I, R, device=x.device, dtype=torch.float32)
            so = (d0 * I * R, I * R, R, 1)
        else:                    # internal layout of the tensor-core path: [d0, R, B, I]
            out = torch.empty(d0, R, B, I, device=x.device, dtype=torch.float32)
            so = (I, R * B * I, 1, B * I)
        shared = Wsets == 1
        # shared parameters + plain inputs: register-tiled kernels on the feature-major copy of x
        tiled = shared and oc == 1 and rr == 1 and B >= 32 and bool(_abi.lib().spn_leaf_tiled_supported(I, R, d0))
        sp, i_fast = (I * R, R, 1), 0
        if shared and layout != "ref":
            # parameters in the order of the activation layout ([F][R][I]) so that a warp's loads are contiguous
            mu, sigma = mu.permute(0, 1, 3, 2).contiguous(), sigma.permute(0, 1, 3, 2).contiguous()
            sp, i_fast = (R * I, 1, I), 1
        st = stream_ptr(x.device)
        xT = None
        # tensor-core leaf: internal layout only, no tanh correction, 128-row tiles worth filling
        use_tc = (tiled and layout != "ref" and LEAF_TENSOR_CORES and not tanh_corr and B >= 128
                  and leaf_tc_supported(I, R, card, d0))
        if tiled:
            xT = torch.empty(_abi.lib().spn_leaf_workspace_floats(B, F), device=x.device, dtype=torch.float32)
            call("spn_leaf_transpose", ptr(x), sx[0], sx[1], B, F, ptr(xT), st)
        if use_tc:
            img = torch.empty(_abi.lib().spn_leaf_tc_image_bytes(I, R, card, d0), device=x.device, dtype=torch.uint8)
            call("spn_leaf_tc_fwd", ptr(xT), ptr(perm32), ptr(mu), ptr(sigma), *sp, B, F, I, R, card, d0, ptr(img),
                 ptr(out), st)
        elif tiled:
            call("spn_leaf_tiled_fwd", ptr(xT), ptr(perm32), ptr(mu), ptr(sigma), *sp, B, F, I, R, card, d0,
                 int(tanh_corr), ptr(out), *so, st)
        elif shared:
            call("spn_leaf_shared_fwd", ptr(x), *sx, ptr(perm32), ptr(mu), ptr(sigma), *sp, i_fast, B, F, I, R, card, d0,
                 int(tanh_corr), ptr(out), *so, st)
        else:
            call("spn_leaf_fwd", ptr(x), *sx, ptr(perm32), ptr(mu), ptr(sigma), Wsets, w, B, F, I, R, card, d0,
                 int(tanh_corr), ptr(out), *so, st)
        empty = torch.empty(0, device=x.device)
        ctx.save_for_backward(x, mu, sigma, perm32 if perm32 is not None else empty, xT if tiled else empty)
        ctx.meta = (sx, so, Wsets, w, B, F, I, R, card, d0, int(tanh_corr), perm32 is not None, sp, i_fast, tiled, use_tc)
        return out

    @staticmethod
    def backward(ctx, g):
        x, mu, sigma, perm32, xT = ctx.saved_tensors
        sx, so, Wsets, w, B, F, I, R, card, d0, tanh_corr, has_perm, sp, i_fast, tiled, use_tc = ctx.meta
        perm32 = perm32 if has_perm else None
        g = _f32c(g)
        dx = dmu = dsigma = None
        st = stream_ptr(x.device)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dmu = torch.empty_like(mu)
            dsigma = torch.empty_like(sigma)
            if use_tc:   # g is contiguous [d0][R][B][I] (_f32c above)
                call("spn_leaf_tc_bwd_params", ptr(xT), ptr(g), ptr(perm32), ptr(mu), ptr(sigma), *sp, B, F, I, R, card, d0,
                     ptr(dmu), ptr(dsigma), st)
            elif tiled:
                call("spn_leaf_tiled_bwd_params", ptr(xT), ptr(g), *so, ptr(perm32), ptr(mu), ptr(sigma), *sp, B, F, I, R,
                     card, d0, ptr(dmu), ptr(dsigma), st)
            elif Wsets == 1:
                call("spn_leaf_shared_bwd_params", ptr(g), *so, ptr(x), *sx, ptr(perm32), ptr(mu), ptr(sigma), *sp,
                     i_fast, B, F, I, R, card, d0, ptr(dmu), ptr(dsigma), st)
            else:
                call("spn_leaf_bwd_params", ptr(g), *so, ptr(x), *sx, ptr(perm32), ptr(mu), ptr(sigma), Wsets, w, B, F,
                     I, R, card, d0, ptr(dmu), ptr(dsigma), st)
            if i_fast:  # back to the parameter layout [1, F, I, R]
                dmu, dsigma = dmu.permute(0, 1, 3, 2), dsigma.permute(0, 1, 3, 2)
        if ctx.needs_input_grad[0]:
            if i_fast:
                mu, sigma = mu.permute(0, 1, 3, 2).contiguous(), sigma.permute(0, 1, 3, 2).contiguous()
            dx = torch.empty_like(x)
            call("spn_leaf_bwd_x", ptr(g), *so, ptr(x), *sx, ptr(perm32), ptr(mu), ptr(sigma), Wsets, w, B, F, I, R,
                 card, d0, tanh_corr, ptr(dx), dx.numel(), st)
        return dx, dmu, dsigma, None, None, None, None


# --------------------------------------------------------------------------------------------------------------
# Sum (optionally fused with the CrossProduct below it; optionally the root, which also reduces over repetitions)
# --------------------------------------------------------------------------------------------------------------
def _sum_geometry(x, lw, mode, layout="ref"):
    """Returns the integer / stride argument pack shared by spn_sum_fwd / _bwd_w / _bwd_x.

    layout 'ref': x [*batch, w, d, c, R] (reference layout); 'drbc': x [d, R, B, c] (internal, shared weights only)."""
    Wsets = lw.shape[0]
    if layout == "ref":
        *lead, w, dx_, cx, R = x.shape
        B = x.numel() // (dx_ * cx * R)
        sx = (dx_ * cx * R, cx * R, R, 1)
    else:
        dx_, R, B, cx = x.shape
        lead, w = (B,), 1
        sx = (cx, R * B * cx, 1, B * cx)
        assert Wsets == 1, "the internal layout is only used with shared weights"
    assert Wsets in (1, w), f"{Wsets} weight sets cannot be matched with {w} conditionals"
    if mode == "sum":
        _, d_out, ic, oc, Rw = lw.shape
        assert (d_out, ic, Rw) == (dx_, cx, R), f"input {tuple(x.shape)} does not fit weights {tuple(lw.shape)}"
        xprod, d_in, c, reduce_r = 0, d_out, 0, 0
        sw = (d_out * ic * oc * R, ic * oc * R, oc * R, R, 1)
        out_shape = (*lead, w, d_out, oc, R)
        so = (d_out * oc * R, oc * R, R, 1)
    elif mode == "xsum":
        _, d_out, ic, oc, Rw = lw.shape
        assert ic == cx * cx and Rw == R and _pow2_ceil(dx_) == 2 * d_out, \
            f"input {tuple(x.shape)} does not fit weights {tuple(lw.shape)}"
        xprod, d_in, c, reduce_r = 1, dx_, cx, 0
        sw = (d_out * ic * oc * R, ic * oc * R, oc * R, R, 1)
        out_shape = (*lead, w, d_out, oc, R)
        so = (d_out * oc * R, oc * R, R, 1)
        if layout != "ref":
            out_shape = (d_out, R, B, oc)
            so = (oc, R * B * oc, 1, B * oc)
    elif mode == "xroot":
        _, one, KR, oc, one2 = lw.shape
        ic = cx * cx
        assert one == 1 and one2 == 1 and KR == R * ic and dx_ <= 2, \
            f"input {tuple(x.shape)} does not fit root weights {tuple(lw.shape)}"
        xprod, d_in, c, reduce_r, d_out = 1, dx_, cx, 1, 1
        sw = (KR * oc, 0, oc, 1, ic * oc)
        out_shape = (*lead, w, 1, oc, 1)
        so = (oc, 0, 1, 0)
    else:
        raise ValueError(mode)
    assert layout == "ref" or mode in ("xsum", "xroot")
    ints = (Wsets, w, B, d_out, ic, oc, R, reduce_r)
    return xprod, sx, d_in, c, sw, ints, so, out_shape


def _sum_backward(x, lw, out, g, mode, layout, need_dx, need_dlw):
    """Exact fp32 backward kernels (general path)."""
    xprod, sx, d_in, c, sw, ints, so, _ = _sum_geometry(x, lw, mode, layout)
    g = _f32c(g)
    st = stream_ptr(x.device)
    dx = dlw = None
    if need_dlw:
        dlw = torch.empty_like(lw)
        call("spn_sum_bwd_w", xprod, ptr(x), *sx, d_in, c, ptr(lw), *sw, *ints, ptr(out), ptr(g), *so, ptr(dlw),
             dlw.numel(), st)
    if need_dx:
        dx = torch.empty_like(x)
        call("spn_sum_bwd_x", xprod, ptr(x), *sx, d_in, c, ptr(lw), *sw, *ints, ptr(out), ptr(g), *so, ptr(dx), *sx, st)
    return dx, dlw


class SumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lw, mode, layout="ref"):
        require_cuda(x, lw)
        x, lw = _f32c(x), _f32c(lw)
        xprod, sx, d_in, c, sw, ints, so, out_shape = _sum_geometry(x, lw, mode, layout)
        out = torch.empty(out_shape, device=x.device, dtype=torch.float32)
        Wsets, w, B, d_out, ic, oc, R, _ = ints
        if (mode == "xsum" and layout == "ref" and Wsets == w and w > 1 and d_out <= 65535 and ic * oc * R * 4 >= 16384
                and not FORCE_GENERAL_SUM  # A/B switch of the tests: force the general kernels
                and _abi.lib().spn_sum_ps_supported(c, oc, R)):
            # per-conditional weights, big enough slabs: TMA-staged HBM-streaming kernel
            call("spn_sum_ps_fwd", ptr(x), sx[0], sx[1], d_in, c, ptr(lw), w, B, d_out, oc, R, ptr(out), stream_ptr(x.device))
            # one row per weight set: the fused streaming backward applies as well
            ctx.ps_bwd = (B == w and bool(_abi.lib().spn_sum_ps_bwd_supported(c, oc, R)),
                          (sx[0], sx[1], d_in, c, w, B, d_out, oc, R))
        else:
            call("spn_sum_fwd", xprod, ptr(x), *sx, d_in, c, ptr(lw), *sw, *ints, ptr(out), *so, stream_ptr(x.device))
            ctx.ps_bwd = (False, None)
        ctx.save_for_backward(x, lw, out)
        ctx.mode, ctx.layout = mode, layout
        return out

    @staticmethod
    def backward(ctx, g):
        x, lw, out = ctx.saved_tensors
        if ctx.ps_bwd[0] and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            sxb, sxd, d_in, c, w, B, d_out, oc, R = ctx.ps_bwd[1]
            g = _f32c(g)
            dx, dlw = torch.empty_like(x), torch.empty_like(lw)
            call("spn_sum_ps_bwd", ptr(x), sxb, sxd, d_in, c, ptr(lw), w, B, d_out, oc, R, ptr(out), ptr(g), ptr(dlw),
                 ptr(dx), sxb, sxd, stream_ptr(x.device))
            return (dx if ctx.needs_input_grad[0] else None), (dlw if ctx.needs_input_grad[1] else None), None, None
        dx, dlw = _sum_backward(x, lw, out, g, ctx.mode, ctx.layout, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dx, dlw, None, None


def sum_raw_supported(x: torch.Tensor, raw: torch.Tensor) -> bool:
    """Can SumRawFn take this (activations [*, w, d_in, c, R], unnormalised per-conditional weights [w, d, c*c, oc, R])?
    One row per weight set, slabs of at least 16 KB, channel counts of the streaming kernels."""
    if x.dim() < 4 or raw.dim() != 5 or not x.is_cuda:
        return False
    w, d_out, ic, oc, R = raw.shape
    c = x.shape[-2]
    B = x.numel() // (x.shape[-3] * c * R) if x.shape[-1] == R else -1
    return (w > 1 and B == w and c * c == ic and d_out <= 65535 and ic * oc * R * 4 >= 16384
            and x.dtype == torch.float32 and raw.dtype == torch.float32
            and bool(_abi.lib().spn_sum_ps_supported(c, oc, R)) and bool(_abi.lib().spn_sum_ps_bwd_supported(c, oc, R)))


class SumRawFn(torch.autograd.Function):
    """Fused CrossProduct + Sum for per-conditional UNNORMALISED weights (CSPN gating-net head outputs): the
    F.log_softmax(dim=2) of cspn.py:274, :282 is folded into the streaming kernels (spn_sum_ps_fwd_raw / _bwd_raw), which
    saves two full passes over the weights in the forward and three in the backward."""

    @staticmethod
    def forward(ctx, x, raw):
        require_cuda(x, raw)
        x, raw = _f32c(x), _f32c(raw)
        w, d_out, ic, oc, R = raw.shape
        d_in, c = x.shape[-3], x.shape[-2]
        out = torch.empty(*x.shape[:-3], d_out, oc, R, device=x.device, dtype=torch.float32)
        logz = torch.empty(w, d_out, oc, R, device=x.device, dtype=torch.float32)
        sxb, sxd = d_in * c * R, c * R
        call("spn_sum_ps_fwd_raw", ptr(x), sxb, sxd, d_in, c, ptr(raw), w, w, d_out, oc, R, ptr(out), ptr(logz),
             stream_ptr(x.device))
        ctx.save_for_backward(x, raw, out, logz)
        return out

    @staticmethod
    def backward(ctx, g):
        x, raw, out, logz = ctx.saved_tensors
        w, d_out, ic, oc, R = raw.shape
        d_in, c = x.shape[-3], x.shape[-2]
        sxb, sxd = d_in * c * R, c * R
        dx, draw = torch.empty_like(x), torch.empty_like(raw)
        call("spn_sum_ps_bwd_raw", ptr(x), sxb, sxd, d_in, c, ptr(raw), w, w, d_out, oc, R, ptr(logz), ptr(out),
             ptr(_f32c(g)), ptr(draw), ptr(dx), sxb, sxd, stream_ptr(x.device))
        return (dx if ctx.needs_input_grad[0] else None), (draw if ctx.needs_input_grad[1] else None)


def sum_forward_raw(x, raw):
    return SumRawFn.apply(x, raw)


class SumTcFn(torch.autograd.Function):
    """Fused CrossProduct + Sum on the tensor cores for weights shared by many rows; activations in the internal
    [d, R, rows, c] layout with the rows of one weight set contiguous (``w.shape[0]`` sets of equal size).

    ``raw = False``: ``w`` holds log-weights.  ``raw = True``: ``w`` is the unnormalised parameter tensor and the
    ``F.log_softmax(dim=2)`` of layers.py:85 (forward and backward) is fused into the operand-image / dW-epilogue
    kernels.  ``cache`` (a dict owned by the layer) keeps the bf16 operand image while the parameter is unchanged."""

    @staticmethod
    def forward(ctx, x, w, raw=False, cache=None):
        wimg, logz = tc_prepare_cached(w, raw, cache)
        # when a backward pass will follow, the forward kernel hands its packed per-row vectors (exp(x - max), shifts) over
        fhand = tc_fhand_buffer(x, w) if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) else None
        out = tc_sum_forward(x, wimg, w, logz, fhand)
        empty = torch.empty(0, device=x.device)
        ctx.save_for_backward(x, w, out, wimg, logz if raw else empty, fhand if fhand is not None else empty)
        ctx.raw, ctx.has_fhand = raw, fhand is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, out, wimg, logz, fhand = ctx.saved_tensors
        logz = logz if ctx.raw else None
        g = _f32c(g)
        # forward -> dX -> dW hand-offs: the dX kernel takes exp(x - max) from the forward and passes G on to the dW
        # kernel, which recomputes nothing
        need_dw = ctx.needs_input_grad[1]
        if not ctx.has_fhand:
            raise RuntimeError("the backward kernels need the forward's hand-off buffer (forward ran without grad inputs)")
        hand = tc_hand_buffer(x, w)
        work = tc_dw_workspace(x, w) if need_dw else None
        csum = tc_csum_view(work, x, w) if (need_dw and logz is not None) else None   # filled by the G pre-pass of dX
        dx = tc_sum_backward_x(x, out, g, wimg, w, fhand, hand, csum)
        dw = tc_sum_backward_w(fhand, hand, x, out, g, w, logz, work, csum is not None) if need_dw else None
        return (dx if ctx.needs_input_grad[0] else None), dw, None, None


class RootSharedFn(torch.autograd.Function):
    """Root Sum (shared weights) fused with the last CrossProduct on the internal layout:
    x [d_in, R, B, c], lw [1, 1, R*c*c, C, 1] (normalised log-weights) -> out [B, C]."""

    @staticmethod
    def forward(ctx, x, lw):
        require_cuda(x, lw)
        x, lw = _f32c(x), _f32c(lw)
        d_in, R, B, c = x.shape
        C = lw.shape[3]
        assert lw.shape[2] == R * c * c, f"root weights {tuple(lw.shape)} do not fit the input {tuple(x.shape)}"
        part = torch.empty(R, B, C, device=x.device, dtype=torch.float32)
        out = torch.empty(B, C, device=x.device, dtype=torch.float32)
        call("spn_root_shared_fwd", ptr(x), d_in, ptr(lw), B, c, C, R, ptr(part), ptr(out), stream_ptr(x.device))
        ctx.save_for_backward(x, lw, out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, lw, out = ctx.saved_tensors
        d_in, R, B, c = x.shape
        C = lw.shape[3]
        g = _f32c(g)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dlw = torch.empty_like(lw) if ctx.needs_input_grad[1] else None
        call("spn_root_shared_bwd", ptr(x), d_in, ptr(lw), ptr(out), ptr(g), B, c, C, R, ptr(dx), ptr(dlw),
             stream_ptr(x.device))
        return dx, dlw


def root_shared_supported(c: int, C: int, R: int) -> bool:
    return bool(_abi.lib().spn_root_shared_supported(int(c), int(C), int(R)))


def root_forward_shared(x, lw):
    return RootSharedFn.apply(x.contiguous(), lw.contiguous())


class CrossProductFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        require_cuda(x)
        x = _f32c(x)
        *lead, w, d, c, R = x.shape
        B = x.numel() // (d * c * R)
        d_out = max(_pow2_ceil(d) // 2, 1)
        out = torch.empty(*lead, w, d_out, c * c, R, device=x.device, dtype=torch.float32)
        call("spn_xprod_fwd", ptr(x), B, d, c, R, ptr(out), stream_ptr(x.device))
        ctx.shape = (tuple(x.shape), B, d, c, R)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, B, d, c, R = ctx.shape
        g = _f32c(g)
        dx = torch.empty(shape, device=g.device, dtype=torch.float32)
        call("spn_xprod_bwd", ptr(g), B, d, c, R, ptr(dx), stream_ptr(g.device))
        return dx


def leaf_forward(x, mu, sigma, perm32, card, tanh_corr, layout="ref"):
    return LeafFn.apply(x, mu, sigma, perm32, card, tanh_corr, layout)


def sum_forward(x, lw, mode="sum", layout="ref"):
    return SumFn.apply(x, lw, mode, layout)


def sum_forward_tc(x, w, raw=False, cache=None):
    return SumTcFn.apply(x.contiguous(), w.contiguous(), raw, cache)


def cross_product(x):
    return CrossProductFn.apply(x)


# --------------------------------------------------------------------------------------------------------------
# Sampling kernels (no autograd: index mode)
# --------------------------------------------------------------------------------------------------------------
def _i32c(t):
    return t.to(torch.int32).contiguous() if t is not None else None


def sampling_weights(lw: torch.Tensor) -> torch.Tensor:
    """Shared log-weights [1, d, ic, oc, R] re-laid out [1, d, R, oc, ic] (in-channel contiguous) for the sampling
    kernels: every (sample, scope) reads ONE column (its parent's out-channel) over all in-channels, which is a 4-byte
    gather with stride oc*R in the parameter layout (1600 sectors per column at cfg 2) and one contiguous 6.4 KB segment
    in this one.  Pass the result to sample_sum(..., transposed=True)."""
    require_cuda(lw)
    return _f32c(lw).permute(0, 1, 4, 3, 2).contiguous()


def sample_sum(lw, root, Q, d, Rs, pa, pa_strides, unravel, rep, xc, Qe, gumbel, mpe, seed=0, q_offset=0,
               transposed=False, one_rep_per_group=False):
    """One Sum layer of the top-down pass (include/spn_b200.h: spn_sample_sum).  Returns int32 [Q, d, Rs].
    ``transposed``: ``lw`` comes from sampling_weights() ([W, d, R, oc, ic]).  ``one_rep_per_group``: kernel-variant
    selector (bit 1 of the ABI's ``mpe`` flags): same indices, one repetition per lane group instead of four."""
    require_cuda(lw, pa, rep, xc, gumbel)
    lw = _f32c(lw)
    Wsets = lw.shape[0]
    if transposed:
        _, dl, Rw, oc, ic = lw.shape
    elif root:
        _, _, ic, oc, _ = lw.shape
    else:
        _, dl, ic, oc, Rw = lw.shape
    w = Wsets if Wsets != 1 else 1
    c, xc_din, sxc, ic_top = 0, 0, (0, 0, 0, 0), 0
    if xc is not None:
        xc = _f32c(xc)
        xc_din, c, Rx = xc.shape[-3:]
        sxc = (xc_din * c * Rx, c * Rx, Rx, 1)
    if root:
        ic_top = ic // root  # root = number of repetitions; in-channel k = r * ic_top + i
        # (w, d, i, o, r) element strides; the repetition offset is applied through ic_top inside the kernel
        sw = (ic * oc, 0, 1, ic, ic_top) if transposed else (ic * oc, 0, oc, 1, ic_top * oc)
        if xc is None:
            c = int(round(math.sqrt(ic_top)))
    else:
        assert dl == d
        sw = (dl * Rw * oc * ic, Rw * oc * ic, 1, ic, oc * ic) if transposed else \
            (dl * ic * oc * Rw, ic * oc * Rw, oc * Rw, Rw, 1)
    out = torch.empty(Q, d, Rs, device=lw.device, dtype=torch.int32)
    hseed, dseed = _seed_args(seed)
    call("spn_sample_sum", ptr(lw), *sw, Wsets, w, Q, d, ic, oc, Rs, ptr(pa), *pa_strides, int(unravel), ptr(rep),
         ptr(xc), *sxc, xc_din, c, Qe, 1 if root else 0, ic_top, ptr(gumbel), int(mpe), hseed, q_offset, dseed,
         ptr(out), stream_ptr(lw.device))
    return out


def sample_leaf(mu, sigma, Q, card, Rs, pa, pa_strides, rep, eps, mpe, inv_perm32, evidence, ev_strides, Qe, squash,
                want_channels=True, seed=0, q_offset=0):
    """Leaf step + post-processing (include/spn_b200.h: spn_sample_leaf).  Returns (samples [Q,F,Rs], channels)."""
    require_cuda(mu, sigma, pa, rep, eps, inv_perm32, evidence)
    mu, sigma = _f32c(mu), _f32c(sigma)
    Wsets, F, I, R = mu.shape
    w = Wsets if Wsets != 1 else 1
    out = torch.empty(Q, F, Rs, device=mu.device, dtype=torch.float32)
    ch = torch.empty(Q, F, Rs, device=mu.device, dtype=torch.int32) if want_channels else None
    hseed, dseed = _seed_args(seed)
    call("spn_sample_leaf", ptr(mu), ptr(sigma), Wsets, w, Q, F, I, R, card, Rs, ptr(pa), *pa_strides, ptr(rep),
         ptr(eps), int(mpe), hseed, q_offset, dseed, ptr(inv_perm32), ptr(evidence), *ev_strides, Qe, int(squash), ptr(out),
         ptr(ch), stream_ptr(mu.device))
    return out, ch


def _sum_sample_geometry(lw, root, xc):
    Wsets = lw.shape[0]
    if root:
        _, _, KR, oc, _ = lw.shape
        ic, ic_top = KR, KR // root
        sw = (KR * oc, 0, oc, 1, ic_top * oc)
        c = int(round(math.sqrt(ic_top)))
    else:
        _, dl, ic, oc, R = lw.shape
        ic_top = 0
        sw = (dl * ic * oc * R, ic * oc * R, oc * R, R, 1)
        c = int(round(math.sqrt(ic)))
    xc_din, sxc = 0, (0, 0, 0, 0)
    if xc is not None:
        xc_din, cx, Rx = xc.shape[-3:]
        assert cx == c
        sxc = (xc_din * cx * Rx, cx * Rx, Rx, 1)
    return Wsets, ic, oc, ic_top, c, sw, xc_din, sxc


def sample_sum_backward(lw, root, Q, d, Rs, pa, pa_strides, unravel, rep, xc, Qe, gumbel, seed, gc, gc_d, dlw, dxc,
                        want_parent, gc_allreps=False, q_offset=0):
    """Straight-through gradient of one Sum layer of the onehot sampling pass (spn_sample_sum_bwd); accumulates into
    dlw / dxc and returns the gradient w.r.t. the parent one-hots [Q, d, oc, Rs] (None at the top layer)."""
    require_cuda(lw, pa, rep, xc, gumbel, gc, dlw, dxc)
    assert lw.is_contiguous() and dlw.is_contiguous() and gc.is_contiguous()
    Wsets, ic, oc, ic_top, c, sw, xc_din, sxc = _sum_sample_geometry(lw, root, xc)
    w = Wsets if Wsets != 1 else 1
    gp = torch.empty(Q, d, oc, Rs, device=lw.device, dtype=torch.float32) if want_parent and not root else None
    hseed, dseed = _seed_args(seed)
    call("spn_sample_sum_bwd", ptr(lw), *sw, Wsets, w, Q, d, ic, oc, Rs, ptr(pa), *pa_strides, int(unravel), ptr(rep),
         ptr(xc), *sxc, xc_din, c, Qe, 1 if root else 0, ic_top, ptr(gumbel), hseed, q_offset, dseed, ptr(gc), gc_d,
         int(gc_allreps), ptr(dlw), ptr(dxc), ptr(gp), stream_ptr(lw.device))
    return gp


def sample_leaf_backward(mu, sigma, Q, card, Rs, ch, rep, eps, mpe, seed, perm32, evidence, ev_strides, Qe, squash, y,
                         gy, want_g0, g0_allreps=False, q_offset=0):
    """Leaf step of the onehot sampling backward (spn_sample_leaf_bwd): returns (dmu, dsigma, g0 [Q, dP, I, Rs])."""
    require_cuda(mu, sigma, ch, rep, eps, perm32, evidence, y, gy)
    assert mu.is_contiguous() and sigma.is_contiguous() and gy.is_contiguous()
    Wsets, F, I, R = mu.shape
    w = Wsets if Wsets != 1 else 1
    dP = -(-F // card)
    dmu, dsigma = torch.zeros_like(mu), torch.zeros_like(sigma)
    g0 = torch.empty(Q, dP, I, R if g0_allreps else Rs, device=mu.device, dtype=torch.float32) if want_g0 else None
    hseed, dseed = _seed_args(seed)
    call("spn_sample_leaf_bwd", ptr(mu), ptr(sigma), Wsets, w, Q, F, I, R, card, Rs, dP, ptr(ch), ptr(rep), ptr(eps),
         int(mpe), hseed, q_offset, dseed, ptr(perm32), ptr(evidence), *ev_strides, Qe, int(squash), ptr(y), int(g0_allreps), ptr(gy),
         ptr(dmu), ptr(dsigma), ptr(g0), stream_ptr(mu.device))
    return dmu, dsigma, g0


def _seed_args(seed):
    """(host seed, device seed pointer): ``seed`` is an int, or a one-element int64 CUDA tensor read by the kernels."""
    if isinstance(seed, torch.Tensor):
        assert seed.is_cuda and seed.dtype == torch.int64 and seed.numel() == 1
        return 0, ptr(seed)
    return int(seed), None


def new_seed(device, mpe: bool = False):
    """Seed for the in-kernel noise of one sampling pass.  Eager: a host integer from torch's CPU generator
    (reproducible under torch.manual_seed, no device work).  While a CUDA graph is being captured: a device word drawn
    with torch's graph-safe CUDA generator, so every replay of the graph samples fresh noise."""
    if mpe:
        return 0
    if torch.device(device).type == "cuda" and torch.cuda.is_current_stream_capturing():
        return torch.randint(0, 2 ** 62, (1,), device=device, dtype=torch.int64)
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class EntropyCombineFn(torch.autograd.Function):
    """H = sum_k exp(lw) * (-lw + ce + lwd + own - ll) over the in-channels (and the repetitions when ``reduce_r``);
    see include/spn_b200.h: spn_entropy_combine_fwd.  ``lwd_has_grad``: the responsibility's log-weight term is not
    detached (aux_with_grad), so it cancels the weight-entropy term in the gradient as well as in the value."""

    @staticmethod
    def forward(ctx, lw, ce, own, ll, reduce_r, lwd_has_grad):
        require_cuda(lw, ce, own, ll)
        lw, ce, own, ll = _f32c(lw), _f32c(ce), _f32c(own), _f32c(ll)
        w, d, ic, oc, R = lw.shape
        assert ce.shape == (w, d, ic, R) and own.shape == (w, d, ic, R) and ll.shape == lw.shape
        H = torch.empty((w, d, oc) if reduce_r else (w, d, oc, R), device=lw.device, dtype=torch.float32)
        call("spn_entropy_combine_fwd", ptr(lw), ptr(ce), ptr(own), ptr(ll), w * d, ic, oc, R, int(reduce_r), ptr(H),
             stream_ptr(lw.device))
        ctx.save_for_backward(lw, ce, own, ll)
        ctx.flags = (bool(reduce_r), bool(lwd_has_grad))
        return H

    @staticmethod
    def backward(ctx, gH):
        lw, ce, own, ll = ctx.saved_tensors
        reduce_r, lwd_has_grad = ctx.flags
        w, d, ic, oc, R = lw.shape
        need = ctx.needs_input_grad
        dlw = torch.empty_like(lw) if need[0] else None
        dce = torch.empty_like(ce) if need[1] else None
        down = torch.empty_like(own) if need[2] else None
        dll = torch.empty_like(ll) if need[3] else None
        call("spn_entropy_combine_bwd", ptr(lw), ptr(ce), ptr(own), ptr(ll), ptr(_f32c(gH)), w * d, ic, oc, R,
             int(reduce_r), int(lwd_has_grad), ptr(dlw), ptr(dce), ptr(down), ptr(dll), stream_ptr(lw.device))
        return dlw, dce, down, dll, None, None


def entropy_combine(lw, ce, own, ll, reduce_r=False, lwd_has_grad=False):
    return EntropyCombineFn.apply(lw, ce, own, ll, reduce_r, lwd_has_grad)


def launches() -> int:
    return _abi.LAUNCHES


# --------------------------------------------------------------------------------------------------------------
# Tensor-core path (shared weights, internal activation layout [scope][repetition][batch][channel])
# --------------------------------------------------------------------------------------------------------------
def tc_supported(c: int, oc: int, R: int) -> bool:
    """Channel counts the tensor-core kernels take: c (per child) and oc multiples of 4, at most 64."""
    return bool(_abi.lib().spn_sum_tc_supported(int(c), int(oc), int(R)))


def _tc_dims(x, w):
    """(nsets, Bs, d_in, d_out, c, oc, R) of an internal-layout activation x [d_in, R, nsets*Bs, c] and weights
    w [nsets, d_out, c*c, oc, R]."""
    d_in, R, Btot, c = x.shape
    nsets, d_out, ic, oc, Rw = w.shape
    assert ic == c * c and Rw == R and _pow2_ceil(d_in) == 2 * d_out and Btot % nsets == 0, \
        f"input {tuple(x.shape)} does not fit weights {tuple(w.shape)}"
    return nsets, Btot // nsets, d_in, d_out, c, oc, R


def tc_prepare(w: torch.Tensor, raw: bool = False):
    """fp32 weights [nsets, d, c*c, oc, R] -> bf16 exp() operand images for the tensor-core kernels (uint8 buffer).
    raw = False: ``w`` are log-weights, returns the images.  raw = True: ``w`` are unnormalised parameters, returns
    (images of softmax(w), logz [nsets*d, oc, R])."""
    require_cuda(w)
    w = _f32c(w)
    nsets, d, ic, oc, R = w.shape
    c = int(round(math.sqrt(ic)))
    assert c * c == ic
    nsd = nsets * d
    wimg = torch.empty(_abi.lib().spn_sum_tc_image_bytes(nsd, c, oc, R), device=w.device, dtype=torch.uint8)
    logz = torch.empty(nsd, oc, R, device=w.device, dtype=torch.float32) if raw else None
    call("spn_sum_tc_prepare", ptr(w), nsd, c, oc, R, ptr(logz), ptr(wimg), stream_ptr(w.device))
    return (wimg, logz) if raw else wimg


def tc_prepare_cached(w: torch.Tensor, raw: bool, cache):
    """(wimg, logz) for ``w``; re-used from ``cache`` while the tensor (storage, version, parameter epoch) is unchanged."""
    if torch.cuda.is_current_stream_capturing():
        cache = None  # a captured graph must rebuild the image on every replay (the weights change between replays)
    key = (w.data_ptr(), w._version, tuple(w.shape), bool(raw), PARAM_EPOCH)
    if cache is not None and cache.get("key") == key:
        return cache["wimg"], cache["logz"]
    res = tc_prepare(w, raw)
    wimg, logz = res if raw else (res, None)
    if cache is not None:
        cache.update(key=key, wimg=wimg, logz=logz)
    return wimg, logz


def tc_fhand_buffer(x, w):
    """Scratch through which tc_sum_forward hands exp(x - max) (packed bf16) and the row shifts to the backward kernels."""
    nsets, Bs, d_in, d_out, c, oc, R = _tc_dims(x, w)
    return torch.empty(_abi.lib().spn_sum_tc_fhand_bytes(Bs, nsets * d_out, c, oc, R), device=x.device, dtype=torch.uint8)


def tc_sum_forward(x: torch.Tensor, wimg: torch.Tensor, lw: torch.Tensor, logz: torch.Tensor = None,
                   fhand: torch.Tensor = None) -> torch.Tensor:
    """x [d_in, R, nsets*Bs, c] -> out [d_out, R, nsets*Bs, oc] (fused CrossProduct + Sum on tcgen05)."""
    require_cuda(x, wimg, lw, logz, fhand)
    x, lw = _f32c(x), _f32c(lw)
    nsets, Bs, d_in, d_out, c, oc, R = _tc_dims(x, lw)
    out = torch.empty(d_out, R, nsets * Bs, oc, device=x.device, dtype=torch.float32)
    call("spn_sum_tc_fwd", ptr(x), ptr(wimg), ptr(lw), ptr(logz), Bs, nsets, d_in, d_out, c, oc, R, ptr(out), ptr(fhand),
         stream_ptr(x.device))
    return out


def tc_hand_buffer(x, w):
    """Scratch through which tc_sum_backward_x hands the packed G tiles to tc_sum_backward_w."""
    nsets, Bs, d_in, d_out, c, oc, R = _tc_dims(x, w)
    return torch.empty(_abi.lib().spn_sum_tc_hand_bytes(Bs, nsets * d_out, c, oc, R), device=x.device, dtype=torch.uint8)


def tc_dw_workspace(x, lw):
    """Scratch of tc_sum_backward_w: padded linear-domain gradient followed by the column sums of the softmax backward."""
    nsets, Bs, d_in, d_out, c, oc, R = _tc_dims(x, lw)
    return torch.empty(_abi.lib().spn_sum_tc_dw_workspace_floats(nsets * d_out, c, oc, R), device=x.device, dtype=torch.float32)


def tc_csum_view(work, x, lw):
    nsets, Bs, d_in, d_out, c, oc, R = _tc_dims(x, lw)
    n = nsets * d_out * oc * R
    return work[work.numel() - n:]


def tc_sum_backward_x(x, out, g, wimg, lw, fhand, hand=None, csum=None):
    """dL/dx of tc_sum_forward, [d_in, R, rows, c]: dE = G W^T as one small-K tensor-core GEMM per row tile, contracted with
    the children's exp vectors out of tensor memory.  ``fhand``: the forward's hand-off (required); ``hand``
    (tc_hand_buffer) receives the packed G tiles (written by a pre-pass, read by this kernel and by tc_sum_backward_w);
    ``csum`` (tc_csum_view of the dW workspace) receives the column sums of g for the fused softmax backward."""
    require_cuda(x, out, g, wimg, fhand, hand, csum)
    x, out, g = _f32c(x), _f32c(out), _f32c(g)
    nsets, Bs, d_in, d_out, c, oc, R = _tc_dims(x, lw)
    if hand is None:
        hand = tc_hand_buffer(x, lw)
    dx = torch.empty_like(x)
    call("spn_sum_tc_bwd_x", ptr(x), ptr(out), ptr(g), ptr(wimg), ptr(fhand), Bs, nsets, d_in, d_out, c, oc, R, ptr(dx),
         ptr(hand), ptr(csum), stream_ptr(x.device))
    return dx


def tc_sum_backward_w(fhand, hand, x, out, g, lw, logz=None, work=None, csum_ready=False):
    """Gradient of tc_sum_forward w.r.t. the weight tensor it was given, [nsets, d_out, c*c, oc, R] (tensor-core GEMM over
    the rows of each set) from the hand-off buffers filled by tc_sum_forward(..., fhand) and tc_sum_backward_x(..., hand);
    with ``logz`` the tensor is the unnormalised parameter and the softmax backward is fused."""
    require_cuda(fhand, hand, x, out, g, lw, logz)
    out, g, lw = _f32c(out), _f32c(g), _f32c(lw)
    nsets, Bs, d_in, d_out, c, oc, R = _tc_dims(x, lw)
    if work is None:
        work = tc_dw_workspace(x, lw)
    dlw = torch.empty_like(lw)
    call("spn_sum_tc_bwd_w", ptr(fhand), ptr(hand), ptr(out), ptr(g), ptr(lw), ptr(logz), Bs, nsets, d_in, d_out, c, oc, R,
         ptr(work), int(bool(csum_ready)), ptr(dlw), stream_ptr(x.device))
    return dlw


def tc_aborts() -> int:
    import ctypes
    v = ctypes.c_int(0)
    _abi.check(_abi.lib().spn_debug_tc_aborts(ctypes.byref(v)), "spn_debug_tc_aborts")
    return v.value
