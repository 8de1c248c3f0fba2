"""CPU oracle for the RAT-SPN / CSPN circuit-evaluation hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product package
(``conditional-ratspn-for-rl_b200``) never does; it fails loudly when its CUDA library is missing.

What it is: an independent restatement, in plain eager ``torch`` on the CPU (fp32 or fp64), of the algorithm
that the reference implements in ``rat_spn.py``, ``layers.py``, ``distributions.py`` and ``cspn.py``
(paths relative to the reference repository root).  torch rather than numpy/C because this is a floating point
path whose gradients are part of the parity contract: autograd of this restatement is what the hand-written
CUDA backward kernels are compared against.  Every function cites the reference lines it follows.

Parity pin: the reference ships no golden vectors and its three test files are stale, so the pin is the live
reference itself.  ``oracle/gen_golden.py`` imports the unmodified reference from ``/root/reference`` (build
container only), runs it on seeded inputs with recorded Gumbel/normal noise and stores inputs+outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this file against those fixtures.

The functions work on an explicit parameter dictionary (``params``) instead of module state:
    perm      int64 [F, R]            feature permutation per repetition      (rat_spn.py:143-153)
    mu, sigma float [W, F, I, R]      bounded leaf parameters                  (distributions.py:131-193)
    lw        list of float [W, d_j, ic_j, S, R], bottom -> top, log-normalised (layers.py:81-88)
    lw_root   float [W, 1, R*ic_top, C, 1]                                     (rat_spn.py:289-292)
W is 1 (RatSpn, shared weights) or w (CSPN, one weight set per conditional).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))
LOG_2 = math.log(2.0)


# ----------------------------------------------------------------------------------------------------------------
# Architecture bookkeeping (rat_spn.py:246-313, layers.py:256-274, layers.py:380-412, distributions.py:348-356)
# ----------------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Arch:
    F: int
    D: int
    R: int
    S: int
    I: int
    C: int
    tanh_squash: bool = False
    no_tanh_log_prob_correction: bool = False

    @property
    def card(self) -> int:  # rat_spn.py:305
        return -(-self.F // (2 ** self.D))

    @property
    def leaf_pad(self) -> int:  # distributions.py:348
        return (self.card - self.F % self.card) % self.card

    @property
    def d0(self) -> int:  # layers.py:273
        return -(-self.F // self.card)

    @property
    def P(self) -> int:  # layers.py:390-391 (scopes padded to a power of two)
        return 1 << max(0, (self.d0 - 1).bit_length())

    @property
    def num_sum_layers(self) -> int:  # rat_spn.py:272
        return self.D - 1

    @property
    def max_layer_index(self) -> int:  # rat_spn.py:117
        return 2 * self.D

    def sum_dims(self, j: int) -> Tuple[int, int]:
        """(number of scopes, in_channels) of inner sum layer S_j, j = 1..D-1, bottom -> top."""
        d = 2 ** (self.D - j)
        ic = self.I ** 2 if j == 1 else self.S ** 2
        return d, ic

    @property
    def ic_top(self) -> int:  # rat_spn.py:290
        return self.I ** 2 if self.D == 1 else self.S ** 2

    def xprod_channels(self, j: int) -> int:
        """Input channel count of CrossProduct X_j (j = 0..D-1)."""
        return self.I if j == 0 else self.S


def invert_permutation(perm: torch.Tensor) -> torch.Tensor:
    """inv[perm[f, r], r] = f   (rat_spn.py:15-23, applied per repetition :147-153)."""
    F, R = perm.shape
    inv = torch.empty_like(perm)
    ar = torch.arange(F, dtype=perm.dtype)
    for r in range(R):
        inv[perm[:, r], r] = ar
    return inv


# ----------------------------------------------------------------------------------------------------------------
# Parameter normalisation (what the reference recomputes on every call)
# ----------------------------------------------------------------------------------------------------------------
def bound_means(p: torch.Tensor, min_mean: Optional[float], max_mean: Optional[float]) -> torch.Tensor:
    """distributions.py:131-140 (sigmoid bound only when limits are set; tanh_squash => +-6, :118-122)."""
    if min_mean is None and max_mean is None:
        return p
    return min_mean + (max_mean - min_mean) * torch.sigmoid(p)


def bound_stds(p: torch.Tensor, min_sigma: float, max_sigma: float) -> torch.Tensor:
    """distributions.py:142-149 (linear-space, sigmoid-bounded std; the only mode the experiments use)."""
    return min_sigma + (max_sigma - min_sigma) * torch.sigmoid(p)


def params_from_ratspn_state(sd: Dict[str, torch.Tensor], arch: Arch, *, min_sigma: float = 0.0,
                             max_sigma: float = 1.0, min_mean: Optional[float] = None,
                             max_mean: Optional[float] = None, dtype=torch.float32) -> Dict:
    """Turn a RatSpn ``state_dict`` (key names of SURVEY.md section 5) into the normalised ``params`` dict.

    Sum weights: log_softmax over the in-channel dim (layers.py:85); leaf: distributions.py:156-193.
    """
    if arch.tanh_squash:
        min_mean, max_mean = -6.0, 6.0
    P = {"perm": sd["permutation"].long()}
    P["mu"] = bound_means(sd["_leaf.base_leaf.mean_param"].to(dtype), min_mean, max_mean)
    P["sigma"] = bound_stds(sd["_leaf.base_leaf.std_param"].to(dtype), min_sigma, max_sigma)
    P["lw"] = []
    for j in range(1, arch.D):
        P["lw"].append(torch.log_softmax(sd[f"_inner_layers.{2 * j - 1}.weight_param"].to(dtype), dim=2))
    P["lw_root"] = torch.log_softmax(sd["root.weight_param"].to(dtype), dim=2)
    return P


# ----------------------------------------------------------------------------------------------------------------
# Forward log-likelihood
# ----------------------------------------------------------------------------------------------------------------
def apply_permutation(x: torch.Tensor, perm: torch.Tensor) -> torch.Tensor:
    """xp[..., f, :, r] = x[..., perm[f, r], :, r or 0]   (rat_spn.py:155-170)."""
    R = perm.shape[1]
    if x.size(-1) == 1:
        x = x.expand(*x.shape[:-1], R)
    idx = perm.unsqueeze(-2).expand_as(x)
    return torch.gather(x, -3, idx)


def leaf_log_prob(xp: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor, arch: Arch) -> torch.Tensor:
    """Gaussian log density per (feature, leaf channel, repetition); NaN input => 0 (marginalised).

    distributions.py:199-246.  ``xp`` [*, w, F, 1|I, 1|R] is the permuted, *pre-squash* input;
    result [*, w, F, I, R].
    """
    nan = torch.isnan(xp)
    xs = torch.where(nan, torch.zeros_like(xp), xp)
    lp = -((xs - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - LOG_SQRT_2PI  # :235 / Normal.log_prob
    if arch.tanh_squash and not arch.no_tanh_log_prob_correction:
        lp = lp - 2 * (LOG_2 - xs - torch.nn.functional.softplus(-2 * xs))  # :220, :241
    return torch.where(nan.expand_as(lp), torch.zeros_like(lp), lp)  # :58-61, :243


def leaf_product(lp: torch.Tensor, arch: Arch) -> torch.Tensor:
    """Zero-pad the feature axis to d0*card and add up `card` consecutive features.

    distributions.py:366-382, layers.py:276-312.  [*, w, F, I, R] -> [*, w, d0, I, R].
    """
    if arch.leaf_pad:
        pad = lp.new_zeros(*lp.shape[:-3], arch.leaf_pad, *lp.shape[-2:])
        lp = torch.cat([lp, pad], dim=-3)
    lp = lp.reshape(*lp.shape[:-3], arch.d0, arch.card, *lp.shape[-2:])
    return lp.sum(dim=-3)


def cross_product(h: torch.Tensor) -> torch.Tensor:
    """out[..., s, l*c + r', r] = h[..., 2s, l, r] + h[..., 2s+1, r', r]; odd scope counts are zero padded.

    layers.py:419-440 and :545-563.  [*, w, d, c, R] -> [*, w, ceil_pow2(d)/2, c*c, R].
    """
    d, c, R = h.shape[-3:]
    P = 1 << max(0, (d - 1).bit_length())
    if P != d:
        h = torch.cat([h, h.new_zeros(*h.shape[:-3], P - d, c, R)], dim=-3)
    left = h[..., 0::2, :, :].unsqueeze(-2)   # [*, P/2, c, 1, R]
    right = h[..., 1::2, :, :].unsqueeze(-3)  # [*, P/2, 1, c, R]
    out = left + right
    return out.reshape(*h.shape[:-3], P // 2, c * c, R)


def sum_layer(h: torch.Tensor, lw: torch.Tensor) -> torch.Tensor:
    """out[..., s, o, r] = logsumexp_i(h[..., s, i, r] + lw[w, s, i, o, r])   (layers.py:110-128)."""
    return torch.logsumexp(h.unsqueeze(-2) + lw, dim=-3)


def root_layer(h: torch.Tensor, lw_root: torch.Tensor) -> torch.Tensor:
    """Repetitions become channels (index r*ic + i), then one Sum over R*ic   (rat_spn.py:230-236)."""
    z = h.transpose(-1, -2).flatten(-2, -1).unsqueeze(-1)  # [*, w, 1, R*ic, 1]
    return sum_layer(z, lw_root)


def forward(P: Dict, arch: Arch, x: torch.Tensor, layer_index: Optional[int] = None,
            x_needs_permutation: bool = True, input_cache: Optional[Dict[int, torch.Tensor]] = None,
            microbatch: Optional[int] = None) -> torch.Tensor:
    """Log-likelihood of ``x`` [*, w, F, 1|I, 1|R] at ``layer_index`` (default: root)  (rat_spn.py:189-244).

    ``input_cache`` (if given) receives the *input* of every sum layer keyed by its layer index, which is what
    evidence-conditioned sampling re-uses (layers.py:102-103, utils.py:52-86).
    ``microbatch`` evaluates the leading dim in slices (the reference cannot hold the big configs at once).
    """
    if microbatch is not None and x.shape[0] > microbatch and input_cache is None:
        outs = [forward(P, arch, x[i:i + microbatch], layer_index, x_needs_permutation)
                for i in range(0, x.shape[0], microbatch)]
        return torch.cat(outs, dim=0)
    if layer_index is None:
        layer_index = arch.max_layer_index
    assert x.dim() >= 5 and x.size(-3) == arch.F
    xp = apply_permutation(x, P["perm"]) if x_needs_permutation else x
    h = leaf_product(leaf_log_prob(xp, P["mu"], P["sigma"], arch), arch)
    if layer_index == 0:
        return h
    for li in range(1, min(layer_index, arch.max_layer_index - 1) + 1):
        if li % 2 == 1:
            h = cross_product(h)
        else:
            if input_cache is not None:
                input_cache[li] = h.detach().clone()
            h = sum_layer(h, P["lw"][li // 2 - 1])
    if layer_index == arch.max_layer_index:
        z = h.transpose(-1, -2).flatten(-2, -1).unsqueeze(-1)
        if input_cache is not None:
            input_cache[layer_index] = z.detach().clone()
        h = sum_layer(z, P["lw_root"])
    return h


# ----------------------------------------------------------------------------------------------------------------
# Ancestral sampling, index mode
# ----------------------------------------------------------------------------------------------------------------
class NoiseSource:
    """Hands out Gumbel / normal noise tensors in the order the reference draws them and records them.

    The reference draws: per Sum layer one ``F.gumbel_softmax`` call (layers.py:227) whose noise has the shape of
    the logits ``[nodes, *n, w, d, ic, r]``; the leaf one ``Normal.rsample`` (distributions.py:300-301).
    """

    def __init__(self, generator: Optional[torch.Generator] = None, replay: Optional[Sequence[torch.Tensor]] = None,
                 dtype=torch.float32):
        self.gen = generator
        self.replay = list(replay) if replay is not None else None
        self.record: List[torch.Tensor] = []
        self.dtype = dtype

    def _next(self, shape, kind):
        if self.replay is not None:
            t = self.replay.pop(0)
            assert tuple(t.shape) == tuple(shape), f"noise shape {tuple(t.shape)} != expected {tuple(shape)} ({kind})"
            t = t.to(self.dtype)
        elif kind == "gumbel":
            t = -torch.empty(shape, dtype=self.dtype).exponential_(generator=self.gen).log()
        else:
            t = torch.randn(shape, dtype=self.dtype, generator=self.gen)
        self.record.append(t)
        return t

    def gumbel(self, shape):
        return self._next(shape, "gumbel")

    def normal(self, shape):
        return self._next(shape, "normal")


def _argmax_first(t: torch.Tensor, dim: int) -> torch.Tensor:
    """First maximal index (torch CPU tie rule)."""
    m = t.max(dim=dim, keepdim=True).values
    n = t.shape[dim]
    ar_shape = [1] * t.dim()
    ar_shape[dim] = n
    ar = torch.arange(n).view(ar_shape)
    cand = torch.where(t == m, ar, torch.full_like(ar, n))
    return cand.min(dim=dim).values


def _select_logits(lw: torch.Tensor, parent: torch.Tensor, rep: Optional[torch.Tensor]) -> torch.Tensor:
    """logits[lead, w, s, i, r] = lw[w, s, i, parent[lead, w, s, r], rep or r]   (layers.py:166-180).

    parent [nodes, *n, w, d, r'] (r' = 1 when a repetition was selected), rep [nodes, *n, w] or None.
    Returns [nodes, *n, w, d, ic, r'].
    """
    W, d, ic, oc, R = lw.shape
    lead = parent.shape[:-2]
    w = lead[-1]
    lw_e = lw.expand(w, d, ic, oc, R) if W != w else lw
    lw_e = lw_e.expand(*lead[:-1], w, d, ic, oc, R)
    if rep is not None:
        ridx = rep.view(*rep.shape, 1, 1, 1, 1).expand(*lead, d, ic, oc, 1)
        lw_e = torch.gather(lw_e, -1, ridx)
    pidx = parent.unsqueeze(-2).unsqueeze(-2).expand(*lead, d, ic, 1, parent.shape[-1])
    return torch.gather(lw_e, -2, pidx).squeeze(-2)


def _unravel(parent: torch.Tensor, c: int, keep: int) -> torch.Tensor:
    """CrossProduct top-down step: channel p of scope s -> (p // c, p % c) at scopes (2s, 2s+1)  (layers.py:489-493,
    :505-508).  parent [lead, d, r] -> [lead, min(2d, keep), r]."""
    left = torch.div(parent, c, rounding_mode="floor")
    right = parent % c
    out = torch.stack([left, right], dim=-2)  # [lead, d, 2, r]
    out = out.reshape(*parent.shape[:-2], parent.shape[-2] * 2, parent.shape[-1])
    return out[..., :keep, :]


@dataclass
class OracleSample:
    sample: torch.Tensor
    parent_indices: Optional[torch.Tensor]
    repetition_indices: Optional[torch.Tensor]
    has_rep_dim: bool


def sample_index(P: Dict, arch: Arch, noise: NoiseSource, n: Tuple[int, ...] = (1,),
                 class_index: Optional[torch.Tensor] = None, evidence: Optional[torch.Tensor] = None,
                 is_mpe: bool = False, layer_index: Optional[int] = None, postprocess: bool = True,
                 invert_perm: bool = True) -> OracleSample:
    """Top-down ancestral sampling with integer paths (rat_spn.py:348-498, layers.py:130-237, :442-511,
    distributions.py:248-302, :384-393) followed by rat_spn.py:500-583 post-processing."""
    if isinstance(n, int):
        n = (n,)
    if layer_index is None:
        layer_index = arch.max_layer_index
    W = P["mu"].shape[0]
    cache: Dict[int, torch.Tensor] = {}
    if evidence is not None:
        n = (*n, *evidence.shape[:-4])  # rat_spn.py:425
        with torch.no_grad():
            forward(P, arch, evidence, input_cache=cache)
        w = evidence.shape[-4]
    else:
        w = W
    dt = P["mu"].dtype
    rep = None
    parent = None
    start = layer_index
    if layer_index == arch.max_layer_index:
        lwr = P["lw_root"]  # [W,1,K,C,1]
        K = lwr.shape[2]
        if class_index is not None:  # rat_spn.py:437-442
            w = class_index.shape[0]
            nodes = 1
            ci = class_index.view(1, *([1] * len(n)), w, 1, 1).expand(1, *n, w, 1, 1)
            logits = _select_logits(lwr, ci, None)
        else:
            nodes = arch.C
            lw_e = lwr.expand(w, 1, K, arch.C, 1)
            logits = lw_e.permute(3, 0, 1, 2, 4)  # [C, w, 1, K, 1]
            logits = logits.reshape(arch.C, *([1] * len(n)), w, 1, K, 1).expand(arch.C, *n, w, 1, K, 1)
        if layer_index in cache:
            logits = logits + cache[layer_index]  # layers.py:189-203 (root: no repetition gather)
        if not is_mpe:
            logits = logits + noise.gumbel(logits.shape)
        k0 = _argmax_first(logits, dim=-2)  # [nodes,*n,w,1,1]
        rep = torch.div(k0, arch.ic_top, rounding_mode="floor").squeeze(-1).squeeze(-1)  # rat_spn.py:459-461
        parent = k0 % arch.ic_top
        start = layer_index - 1
        has_rep_dim = False
    else:
        has_rep_dim = True
        if layer_index == 0:
            parent = None
        elif layer_index % 2 == 1:  # CrossProduct root: every node of the layer (layers.py:457-474)
            j = (layer_index - 1) // 2
            c = arch.xprod_channels(j)
            d_out = 2 ** (arch.D - 1 - j)
            k = torch.arange(c * c).view(c * c, *([1] * len(n)), 1, 1, 1).expand(c * c, *n, w, d_out, arch.R)
            parent = k
        else:  # Sum root: every sum node, all scopes and repetitions (layers.py:160-164)
            j = layer_index // 2
            lw = P["lw"][j - 1]
            Wj, d, ic, oc, R = lw.shape
            logits = lw.expand(w, d, ic, oc, R).permute(3, 0, 1, 2, 4)
            logits = logits.reshape(oc, *([1] * len(n)), w, d, ic, R).expand(oc, *n, w, d, ic, R)
            if layer_index in cache:
                logits = logits + cache[layer_index]
            if not is_mpe:
                logits = logits + noise.gumbel(logits.shape)
            parent = _argmax_first(logits, dim=-2)
            start = layer_index - 1

    # walk down the remaining inner layers (rat_spn.py:475-485)
    for li in range(start, 0, -1):
        if li % 2 == 1:
            j = (li - 1) // 2
            keep = arch.d0 if j == 0 else 2 ** (arch.D - j)
            parent = _unravel(parent, arch.xprod_channels(j), keep)
        else:
            lw = P["lw"][li // 2 - 1]
            logits = _select_logits(lw, parent, rep)
            if li in cache:
                off = cache[li]
                if rep is not None:  # layers.py:192-197
                    off = off.expand(*logits.shape[:-1], off.shape[-1])
                    ridx = rep.view(*rep.shape, 1, 1, 1).expand(*logits.shape[:-1], 1)
                    off = torch.gather(off, -1, ridx)
                logits = logits + off
            if not is_mpe:
                logits = logits + noise.gumbel(logits.shape)
            parent = _argmax_first(logits, dim=-2)

    mu, sigma = P["mu"], P["sigma"]
    if layer_index == 0:  # every leaf samples itself (distributions.py:257-260, rat_spn.py:491-492)
        assert len(n) == 1
        mu_e = mu.expand(w, *mu.shape[1:]).expand(*n, w, *mu.shape[1:])
        if is_mpe:
            y = mu_e.clone()
        else:
            y = mu_e + sigma.expand_as(mu_e) * noise.normal(mu_e.shape)
        y = y.permute(3, 0, 1, 2, 4)  # 'nwFIR -> InwFR'
        leaf_parent = None
    else:
        # fan out scopes to features and strip the leaf padding (layers.py:343-362, distributions.py:385-390)
        ch = torch.repeat_interleave(parent, arch.card, dim=-2)[..., :arch.F, :]  # [lead, F, r']
        lead = ch.shape[:-2]
        mu_e = mu.expand(w, *mu.shape[1:]).expand(*lead[:-1], w, *mu.shape[1:])  # [lead, F, I, R]
        sg_e = sigma.expand(w, *mu.shape[1:]).expand(*lead[:-1], w, *mu.shape[1:])
        if rep is not None:
            ridx = rep.view(*rep.shape, 1, 1, 1).expand(*lead, arch.F, arch.I, 1)
            mu_e = torch.gather(mu_e, -1, ridx)
            sg_e = torch.gather(sg_e, -1, ridx)
        mu_s = torch.gather(mu_e, -2, ch.unsqueeze(-2)).squeeze(-2)  # [lead, F, r']
        sg_s = torch.gather(sg_e, -2, ch.unsqueeze(-2)).squeeze(-2)
        if rep is not None:
            mu_s, sg_s = mu_s.squeeze(-1), sg_s.squeeze(-1)
        y = mu_s.clone() if is_mpe else mu_s + sg_s * noise.normal(mu_s.shape)
        leaf_parent = ch
    out = OracleSample(sample=y, parent_indices=leaf_parent, repetition_indices=rep, has_rep_dim=has_rep_dim)
    if postprocess:
        out = sample_postprocess(P, arch, out, evidence, n, invert_perm=invert_perm)
    return out


def sample_postprocess(P: Dict, arch: Arch, s: OracleSample, evidence: Optional[torch.Tensor], n: Tuple[int, ...],
                       invert_perm: bool = True) -> OracleSample:
    """Inverse permutation of the selected repetition, evidence fill-in, tanh squashing (rat_spn.py:538-580)."""
    y = s.sample
    inv = invert_permutation(P["perm"])
    has_rep = s.has_rep_dim
    if invert_perm:
        if s.repetition_indices is not None:
            sel = inv.t()[s.repetition_indices]  # [lead, F]
            y = torch.gather(y, -1, sel)
        else:
            sel = inv.view(*([1] * (y.dim() - 2)), *inv.shape).expand_as(y)
            y = torch.gather(y, -2, sel)
        if not has_rep:
            y = y.unsqueeze(-1)
            has_rep = True
    if evidence is not None:
        ev = evidence.expand(*n, *evidence.shape[-4:])  # [*n, w, F, 1, r]
        ev = ev.movedim(-2, 0)  # '...or -> o...r'
        assert ev.shape == y.shape, (ev.shape, y.shape)
        y = torch.where(torch.isnan(ev), y, ev)
    if arch.tanh_squash:
        y = torch.tanh(y.clamp(-6.0, 6.0))
    return OracleSample(sample=y, parent_indices=s.parent_indices, repetition_indices=s.repetition_indices,
                        has_rep_dim=has_rep)


# ----------------------------------------------------------------------------------------------------------------
# Differentiable ("onehot") sampling: same forward values, straight-through gradients
# ----------------------------------------------------------------------------------------------------------------
def _st_onehot(logits: torch.Tensor, g: Optional[torch.Tensor]) -> torch.Tensor:
    """hard one-hot of argmax(logits + g) with softmax(logits + g) gradients (F.gumbel_softmax(hard=True), tau=1;
    layers.py:206-227).  MPE (g is None): plain one-hot of the argmax, no gradient (layers.py:208-214)."""
    z = logits if g is None else logits + g
    idx = _argmax_first(z, dim=-2)
    hard = torch.zeros_like(z).scatter_(-2, idx.unsqueeze(-2), 1.0)
    if g is None:
        return hard
    soft = torch.softmax(z, dim=-2)
    return hard - soft.detach() + soft


def sample_onehot(P: Dict, arch: Arch, noise: NoiseSource, n: Tuple[int, ...] = (1,),
                  evidence: Optional[torch.Tensor] = None, is_mpe: bool = False,
                  layer_index: Optional[int] = None, postprocess: bool = True) -> OracleSample:
    """Differentiable sampling; paths are float one-hots [nodes, *n, w, d, ic, r]  (layers.py:181-186, :495-503,
    distributions.py:281-295).  The leaf-padding slice follows the *intended* behaviour (feature dim), i.e. the
    one-line fix of the defect at distributions.py:390 documented in SURVEY.md section 0 / 8c."""
    if isinstance(n, int):
        n = (n,)
    if layer_index is None:
        layer_index = arch.max_layer_index
    W = P["mu"].shape[0]
    cache: Dict[int, torch.Tensor] = {}
    if evidence is not None:
        n = (*n, *evidence.shape[:-4])
        forward(P, arch, evidence, input_cache=cache)
        w = evidence.shape[-4]
    else:
        w = W
    start = layer_index
    has_rep_dim = True
    if layer_index == arch.max_layer_index:
        lwr = P["lw_root"]
        K = lwr.shape[2]
        logits = lwr.expand(w, 1, K, arch.C, 1).permute(3, 0, 1, 2, 4)
        logits = logits.reshape(arch.C, *([1] * len(n)), w, 1, K, 1).expand(arch.C, *n, w, 1, K, 1)
        if layer_index in cache:
            logits = logits + cache[layer_index]
        oh = _st_onehot(logits, None if is_mpe else noise.gumbel(logits.shape))  # [C,*n,w,1,K,1]
        oh = oh.reshape(*oh.shape[:-2], arch.R, arch.ic_top).transpose(-1, -2)  # rat_spn.py:466-467
        start = layer_index - 1
        has_rep_dim = False
    elif layer_index == 0:
        oh = None
    elif layer_index % 2 == 1:
        j = (layer_index - 1) // 2
        c = arch.xprod_channels(j)
        d_out = 2 ** (arch.D - 1 - j)
        eye = torch.eye(c * c, dtype=P["mu"].dtype)
        oh = eye.view(c * c, *([1] * len(n)), 1, 1, c * c, 1).expand(c * c, *n, w, d_out, c * c, arch.R)
    else:
        j = layer_index // 2
        lw = P["lw"][j - 1]
        Wj, d, ic, oc, R = lw.shape
        logits = lw.expand(w, d, ic, oc, R).permute(3, 0, 1, 2, 4)
        logits = logits.reshape(oc, *([1] * len(n)), w, d, ic, R).expand(oc, *n, w, d, ic, R)
        if layer_index in cache:
            logits = logits + cache[layer_index]
        oh = _st_onehot(logits, None if is_mpe else noise.gumbel(logits.shape))
        start = layer_index - 1

    for li in range(start, 0, -1):
        if li % 2 == 1:  # CrossProduct (layers.py:495-508)
            j = (li - 1) // 2
            c = arch.xprod_channels(j)
            keep = arch.d0 if j == 0 else 2 ** (arch.D - j)
            t = oh.reshape(*oh.shape[:-2], c, c, oh.shape[-1])  # [lead, d, l, r', r]
            left = t.sum(dim=-2)
            right = t.sum(dim=-3)
            oh = torch.stack([left, right], dim=-3)  # [lead, d, 2, c, r]
            oh = oh.reshape(*oh.shape[:-4], oh.shape[-4] * 2, c, oh.shape[-1])[..., :keep, :, :]
        else:  # Sum (layers.py:181-186, :198-203, :233-234)
            lw = P["lw"][li // 2 - 1]  # [W,d,ic,oc,R]
            logits = (lw * oh.unsqueeze(-3)).sum(-2)  # [lead, d, ic, r]
            clear = (oh.sum(-2, keepdim=True) == 0.0).expand_as(logits)
            if li in cache:
                logits = logits + (cache[li].unsqueeze(-2) * oh.unsqueeze(-3)).sum(-2)
            new = _st_onehot(logits, None if is_mpe else noise.gumbel(logits.shape))
            oh = torch.where(clear, torch.zeros_like(new), new)

    mu, sigma = P["mu"], P["sigma"]
    if layer_index == 0:
        assert len(n) == 1
        mu_e = mu.expand(w, *mu.shape[1:]).expand(*n, w, *mu.shape[1:])
        y = mu_e.clone() if is_mpe else mu_e + sigma.expand_as(mu_e) * noise.normal(mu_e.shape)
        y = y.permute(3, 0, 1, 2, 4)
    else:
        oh_f = torch.repeat_interleave(oh, arch.card, dim=-3)[..., :arch.F, :, :]  # [lead, F, I, r]
        mu_s = (mu * oh_f).sum(-2)
        sg_s = (sigma * oh_f).sum(-2)
        if not has_rep_dim:
            mu_s, sg_s = mu_s.sum(-1), sg_s.sum(-1)
        y = mu_s if is_mpe else mu_s + sg_s * noise.normal(mu_s.shape)
    rep_sel = None
    if not has_rep_dim:
        rep_sel = oh.detach().sum(-2).sum(-2).argmax(-1)  # which repetition carries the path
    out = OracleSample(sample=y, parent_indices=oh if layer_index != 0 else None, repetition_indices=rep_sel,
                       has_rep_dim=has_rep_dim)
    if postprocess:
        out = sample_postprocess(P, arch, out, evidence, n)
    return out


# ----------------------------------------------------------------------------------------------------------------
# Recursive entropy approximation (rat_spn.py:591-797)
# ----------------------------------------------------------------------------------------------------------------
def _detached(P: Dict) -> Dict:
    return {"perm": P["perm"], "mu": P["mu"].detach(), "sigma": P["sigma"].detach(),
            "lw": [t.detach() for t in P["lw"]], "lw_root": P["lw_root"].detach()}


def recursive_entropy(P: Dict, arch: Arch, noise: NoiseSource, sample_size: int = 10, aux_with_grad: bool = False,
                      marginal_mask: Optional[torch.Tensor] = None, differentiable: bool = False) -> torch.Tensor:
    """Layer-by-layer entropy estimate of every root node, returned flattened as [w*C]  (rat_spn.py:776-797).

    ``differentiable`` picks onehot (straight-through) vs index sampling, which the reference derives from
    ``th.is_grad_enabled()`` (rat_spn.py:656, :677)."""
    sampler = sample_onehot if differentiable else sample_index
    w = P["mu"].shape[0]
    mask_p = None
    if marginal_mask is not None:  # rat_spn.py:647-651
        mm = marginal_mask.clone().bool().unsqueeze(-1).unsqueeze(-1)
        mask_p = apply_permutation(mm, P["perm"])  # [w, F, 1, R]
    Paux = P if aux_with_grad else _detached(P)

    # layer 0 (rat_spn.py:653-668): every leaf scores its own samples
    s0 = sampler(P, arch, noise, n=(sample_size,), layer_index=0, postprocess=False).sample  # [I,n,w,F,R]
    x0 = s0.permute(1, 2, 3, 0, 4)  # 'InwFR -> nwFIR'
    if mask_p is not None:
        x0 = torch.where(mask_p.expand_as(x0), torch.full_like(x0, float("nan")), x0)
    H = -forward(P, arch, x0, layer_index=0, x_needs_permutation=False).mean(dim=0)  # [w,d0,I,R]

    for li in range(1, arch.max_layer_index + 1):
        if li % 2 == 1:
            H = cross_product(H)  # rat_spn.py:672-673 (entropies of independent factors add)
            continue
        ys = sampler(P, arch, noise, n=(sample_size,), layer_index=li - 1, postprocess=False).sample  # [K,n,w,F,R]
        xs = ys.unsqueeze(-2)  # [K,n,w,F,1,R]
        if mask_p is not None:
            xs = torch.where(mask_p.expand_as(xs), torch.full_like(xs, float("nan")), xs)
        cll = forward(Paux, arch, xs, layer_index=li - 1, x_needs_permutation=False).mean(dim=1)  # [K,w,d,K,R]
        if li == arch.max_layer_index:  # rat_spn.py:694-705, :331-335
            lwr = P["lw_root"]
            Wr, _, KR, C, _ = lwr.shape
            lw_w = lwr.view(Wr, 1, arch.R, KR // arch.R, C).permute(0, 1, 3, 4, 2)  # [W,1,K,C,R]
            lw_aux = lw_w if aux_with_grad else lw_w.detach()
            ll = torch.logsumexp(cll.unsqueeze(-2) + lw_aux, dim=-3)  # [K,w,1,C,R]
        else:
            lw_w = P["lw"][li // 2 - 1]
            lw_aux = lw_w if aux_with_grad else lw_w.detach()
            ll = sum_layer(cll, lw_aux)  # [K,w,d,S,R]
        ll = ll.permute(1, 2, 0, 3, 4)  # 'iwdoR -> wdioR'
        diag = torch.diagonal(cll, dim1=0, dim2=3)  # [w,d,R,K]
        diag = diag.permute(0, 1, 3, 2).unsqueeze(-2)  # [w,d,K,1,R]
        resp = lw_aux + diag - ll  # rat_spn.py:728
        wts = lw_w.exp()
        w_ent = -(wts * lw_w).sum(dim=2)  # rat_spn.py:611
        inner = H.unsqueeze(3) + resp  # rat_spn.py:731-736
        if li < arch.max_layer_index:
            H = w_ent + (wts * inner).sum(dim=2)
        else:  # rat_spn.py:615-621: sum over channels, then over repetitions; weight entropy from the flat root
            lwr = P["lw_root"]
            w_ent_flat = -(lwr.exp() * lwr).sum(dim=2).squeeze(-1)  # [W,1,C]
            H = w_ent_flat + (wts * inner).sum(dim=2).sum(dim=-1)  # [w,1,C]
    return H.flatten()


# ----------------------------------------------------------------------------------------------------------------
# CSPN gating network -> per-conditional parameters (cspn.py:254-301)
# ----------------------------------------------------------------------------------------------------------------
def cspn_params(sd: Dict[str, torch.Tensor], arch: Arch, condition: torch.Tensor, *, min_sigma: float,
                max_sigma: float, min_mean: Optional[float] = None, max_mean: Optional[float] = None,
                negative_slope: float = 0.01) -> Dict:
    """Evaluate the gating MLP of a CSPN ``state_dict`` on ``condition`` [B, F_cond] and return ``params`` with
    W = B.  Supports the Linear+LeakyReLU stacks the experiments build (cspn.py:197-247)."""
    if arch.tanh_squash:
        min_mean, max_mean = -6.0, 6.0

    def mlp(x, prefix, last_identity):
        idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith(prefix + ".") and k.endswith(".weight")})
        for t, i in enumerate(idx):
            x = torch.nn.functional.linear(x, sd[f"{prefix}.{i}.weight"], sd[f"{prefix}.{i}.bias"])
            if not (last_identity and t == len(idx) - 1):
                x = torch.nn.functional.leaky_relu(x, negative_slope)
        return x

    B = condition.shape[0]
    feat = mlp(condition, "feat_layers", False).flatten(1)
    hs = mlp(feat, "sum_layers", True)
    hd = mlp(feat, "dist_layers", True)
    P = {"perm": sd["permutation"].long(), "lw": []}
    for j in range(1, arch.D):
        d, ic = arch.sum_dims(j)
        raw = torch.nn.functional.linear(hs, sd[f"sum_param_heads.{j - 1}.weight"], sd[f"sum_param_heads.{j - 1}.bias"])
        P["lw"].append(torch.log_softmax(raw.view(B, d, ic, arch.S, arch.R), dim=2))
    h = arch.D - 1
    raw = torch.nn.functional.linear(hs, sd[f"sum_param_heads.{h}.weight"], sd[f"sum_param_heads.{h}.bias"])
    P["lw_root"] = torch.log_softmax(raw.view(B, 1, arch.R * arch.ic_top, arch.C, 1), dim=2)
    shp = (B, arch.F, arch.I, arch.R)
    P["mu"] = bound_means(torch.nn.functional.linear(hd, sd["dist_mean_head.weight"], sd["dist_mean_head.bias"]).view(shp),
                          min_mean, max_mean)
    P["sigma"] = bound_stds(torch.nn.functional.linear(hd, sd["dist_std_head.weight"], sd["dist_std_head.bias"]).view(shp),
                            min_sigma, max_sigma)
    return P


# ----------------------------------------------------------------------------------------------------------------
# Entropy lower bound of Huber et al. 2008 (rat_spn.py:825-951), restated per repetition pair
# ----------------------------------------------------------------------------------------------------------------
def huber_entropy_lb(P: Dict, arch: Arch, layer_index: Optional[int] = None, detach_weights: bool = False,
                     add_sub_weight_ent: bool = False, marginal_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """H(sum_k w_k p_k) >= -sum_k w_k log sum_j w_j z_kj with z_kj = int p_k p_j dx.

    The reference carries log z for every pair of nodes and every pair of repetitions.  Leaves (rat_spn.py:849-861):
    int N(x; m1, v1) N(x; m2, v2) dx = N(m1; m2, v1 + v2), evaluated between slot f of repetition d and the slot of
    repetition c that holds the same original feature (the reference inverts the permutation per repetition and
    re-applies the one of the second repetition index); masked original features contribute log z = 0.
    Product / CrossProduct add the logs of the joined scopes (:864-887), a Sum mixes linearly (:888-921):
        M[a, o, c, d]   = log sum_b w[b, o, d] z[a, b, c, d]
        bound[o, c]     = -sum_a w[a, o, c] M[a, o, c, c]                 (root: M also summed over d, bound over c)
        z'[o1, o2, c, d] = sum_a w[a, o2, c] exp(M[a, o1, c, d])
    Returns the bound of the nodes of ``layer_index`` ([W, d, oc, R]; root [W, 1, C])."""
    if layer_index is None:
        layer_index = arch.max_layer_index
    assert 1 <= layer_index <= arch.max_layer_index
    mu, var, perm = P["mu"], P["sigma"] ** 2, P["perm"]
    inv = invert_permutation(perm)
    W, F, I, R = mu.shape
    assert F % arch.card == 0, "the reference reshapes [F] -> [F / card, card] without padding (rat_spn.py:866-868)"
    per_pair = []
    for c in range(R):
        row = []
        for d in range(R):
            slot_c = inv[perm[:, d], c]                                    # slot of repetition c with the same feature
            m1, v1 = mu[:, slot_c, :, c], var[:, slot_c, :, c]             # [W, F, I]  (index a)
            m2, v2 = mu[:, :, :, d], var[:, :, :, d]                       # [W, F, I]  (index b)
            v = v1.unsqueeze(3) + v2.unsqueeze(2)
            lz = -0.5 * (m1.unsqueeze(3) - m2.unsqueeze(2)) ** 2 / v - 0.5 * torch.log(2.0 * math.pi * v)
            if marginal_mask is not None:
                masked = marginal_mask.bool()[:, perm[:, d]]               # [W, F]: mask of the original feature
                lz = torch.where(masked.view(*masked.shape, 1, 1), torch.zeros_like(lz), lz)
            row.append(lz)
        per_pair.append(torch.stack(row, dim=-1))                           # [W, F, a, b, d]
    L = torch.stack(per_pair, dim=-2)                                       # [W, F, a, b, c, d]
    L = L.reshape(W, F // arch.card, arch.card, I, I, R, R).sum(2)

    bound = None
    for li in range(1, layer_index + 1):
        if li % 2 == 1:  # CrossProduct: pad the scopes to an even count with log z = 0, pair (2j, 2j + 1)
            if L.shape[1] % 2 == 1:
                L = torch.cat((L, torch.zeros_like(L[:, :1])), dim=1)
            left, right = L[:, 0::2], L[:, 1::2]
            o = left.shape[2]
            # new node (a', a): right child a', left child a  (index order of rat_spn.py:872-887)
            L = (right[:, :, :, None, :, None] + left[:, :, None, :, None, :]).reshape(W, left.shape[1], o * o, o * o, R, R)
            continue
        is_root = li == arch.max_layer_index
        if is_root:
            lw_full = P["lw_root"]                                           # [W, 1, R * ic, C, 1], k = r * ic + i
            ic = lw_full.shape[2] // R
            lw = lw_full.view(lw_full.shape[0], 1, R, ic, -1).permute(0, 1, 3, 4, 2)   # [W, 1, ic, C, R]
            w_ent = -(lw_full * lw_full.exp()).sum(2)                       # [W, 1, C, 1]
        else:
            lw = P["lw"][li // 2 - 1]                                        # [W, d, ic, oc, R]
            w_ent = -(lw * lw.exp()).sum(2)                                  # [W, d, oc, R]
        wgt = lw.detach() if detach_weights else lw
        # M[w, s, a, o, c, d] = logsumexp_b (L[w, s, a, b, c, d] + wgt[w, s, b, o, d])
        M = torch.logsumexp(L.unsqueeze(4) + wgt[:, :, None, :, :, None, :], dim=3)
        if add_sub_weight_ent:
            t = w_ent.unsqueeze(2).unsqueeze(-1)                             # broadcast as rat_spn.py:903-905
            M = M - t + t.detach()
        if is_root:
            M = torch.logsumexp(M, dim=-1)                                   # [W, 1, a, C, c]
            bound = -(wgt.exp() * M).sum(dim=2).sum(-1)                      # [W, 1, C]
        else:
            own = torch.diagonal(M, dim1=-2, dim2=-1)                        # [W, d, a, o, c]
            bound = -(wgt.exp() * own).sum(dim=2)                            # [W, d, o, R]
        if li < layer_index:
            # z'[w, s, o1, o2, c, d] = sum_a w[a, o2, c] exp(M[a, o1, c, d])
            L = torch.logsumexp(M.unsqueeze(4) + wgt[:, :, :, None, :, :, None], dim=2)
    return bound
