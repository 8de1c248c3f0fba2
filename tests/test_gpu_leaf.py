"""GPU: register-tiled shared-parameter leaf kernels (forward + parameter gradients) against the CPU oracle in fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_leaf_kernels():
    """This file pins the exact fp32 kernels (tolerance 2e-6); the tensor-core leaf has its own file and tolerance."""
    from ratspn_b200 import ops
    ops.LEAF_TENSOR_CORES = False
    yield
    ops.LEAF_TENSOR_CORES = True


def _oracle_leaf(x, perm, mu, sigma, arch):
    from oracle import spn_oracle as O
    xp = O.apply_permutation(x.double(), perm)
    return O.leaf_product(O.leaf_log_prob(xp, mu, sigma, arch), arch)


@pytest.mark.parametrize("F,D,I,R,B,tanh,layout,nan", [
    (64, 3, 5, 3, 200, False, "ref", False),      # card 8, I not a multiple of 4
    (64, 3, 5, 3, 200, False, "drbc", True),
    (784, 3, 10, 5, 256, False, "drbc", False),   # cfg 1 leaf: card 98 (four feature chunks), row-split backward
    (784, 5, 40, 4, 300, False, "drbc", True),    # cfg 2 leaf shape (fewer repetitions): card 25, padded last scope
    (17, 4, 3, 3, 64, True, "ref", True),         # tanh correction + leaf padding (SAC action space)
    (100, 2, 33, 2, 130, False, "drbc", False),   # card 25, I = 33
])
def test_tiled_leaf_forward_and_param_grads_vs_oracle(F, D, I, R, B, tanh, layout, nan):
    from oracle import spn_oracle as O
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(F * 7 + B)
    arch = O.Arch(F=F, D=D, R=R, S=2, I=I, C=1, tanh_squash=tanh)
    card, d0 = arch.card, arch.d0
    x = torch.randn(B, 1, F, 1, 1, generator=g)
    if nan:
        x[3, 0, ::3] = float("nan")
        x[B - 1, 0, :] = float("nan")
        x[17, 0, 5] = float("nan")
    perm = torch.stack([torch.randperm(F, generator=g) for _ in range(R)], dim=-1)
    mu = torch.randn(1, F, I, R, generator=g)
    sigma = torch.rand(1, F, I, R, generator=g) * 0.9 + 0.05
    go = torch.randn(B, 1, d0, I, R, generator=g)

    mu_o, sg_o = mu.double().requires_grad_(True), sigma.double().requires_grad_(True)
    ref = _oracle_leaf(x, perm, mu_o, sg_o, arch)                    # [B, 1, d0, I, R]
    (ref * go.double()).sum().backward()

    mu_t, sg_t = mu.to(DEV).requires_grad_(True), sigma.to(DEV).requires_grad_(True)
    out = ops.leaf_forward(x.to(DEV), mu_t, sg_t, perm.to(torch.int32).to(DEV), card, tanh, layout)
    if layout == "drbc":
        assert out.shape == (d0, R, B, I)
        got = out.permute(2, 0, 3, 1).unsqueeze(1)                   # -> [B, 1, d0, I, R]
    else:
        got = out
    (got * go.to(DEV)).sum().backward()
    scale = ref.abs().max().item()
    assert (got.cpu().double() - ref).abs().max().item() <= 2e-6 * scale + 1e-5
    for name, a, b in (("dmu", mu_t.grad, mu_o.grad), ("dsigma", sg_t.grad, sg_o.grad)):
        err = (a.cpu().double() - b).abs().max().item()
        assert err <= 2e-5 * b.abs().max().item() + 1e-6, f"{name}: {err} vs scale {b.abs().max().item()}"


def test_tiled_leaf_all_nan_rows_are_fully_marginalised():
    from ratspn_b200 import ops
    F, I, R, B, card = 32, 4, 2, 64, 4
    x = torch.full((B, 1, F, 1, 1), float("nan"), device=DEV)
    mu = torch.randn(1, F, I, R, device=DEV)
    sigma = torch.rand(1, F, I, R, device=DEV) + 0.1
    out = ops.leaf_forward(x, mu, sigma, None, card, False, "ref")
    assert torch.equal(out, torch.zeros_like(out))


@pytest.mark.parametrize("n,w,F,D,I,R,tanh,nan", [
    (1, 600, 64, 3, 16, 8, False, False),   # vectorised forward and backward (enough parameters to fill the GPU)
    (3, 300, 40, 2, 8, 4, True, True),      # several rows per weight set, tanh correction, NaN marginalisation
    (2, 5, 17, 4, 3, 4, True, True),        # small: vectorised forward, scalar backward; leaf padding
])
def test_per_conditional_leaf_vectorised_kernels_vs_oracle(n, w, F, D, I, R, tanh, nan):
    """CSPN leaf (one parameter set per conditional, R % 4 == 0): the 16-byte-vectorised streaming kernels against
    the fp64 oracle; x [n, w, F, 1, 1]."""
    from oracle import spn_oracle as O
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(F * 3 + w)
    arch = O.Arch(F=F, D=D, R=R, S=2, I=I, C=1, tanh_squash=tanh)
    card, d0 = arch.card, arch.d0
    x = torch.randn(n, w, F, 1, 1, generator=g)
    if nan:
        x[0, 1, ::3] = float("nan")
        x[n - 1, w - 1, :] = float("nan")
    perm = torch.stack([torch.randperm(F, generator=g) for _ in range(R)], dim=-1)
    mu = torch.randn(w, F, I, R, generator=g)
    sigma = torch.rand(w, F, I, R, generator=g) * 0.9 + 0.05
    go = torch.randn(n, w, d0, I, R, generator=g)
    mu_o, sg_o = mu.double().requires_grad_(True), sigma.double().requires_grad_(True)
    ref = _oracle_leaf(x, perm, mu_o, sg_o, arch)
    (ref * go.double()).sum().backward()
    mu_t, sg_t = mu.to(DEV).requires_grad_(True), sigma.to(DEV).requires_grad_(True)
    out = ops.leaf_forward(x.to(DEV), mu_t, sg_t, perm.to(torch.int32).to(DEV), card, tanh, "ref")
    assert out.shape == ref.shape
    (out * go.to(DEV)).sum().backward()
    scale = ref.abs().max().item()
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-6 * scale + 1e-5
    for name, a, b in (("dmu", mu_t.grad, mu_o.grad), ("dsigma", sg_t.grad, sg_o.grad)):
        err = (a.cpu().double() - b).abs().max().item()
        assert err <= 2e-5 * b.abs().max().item() + 1e-6, f"{name}: {err} vs scale {b.abs().max().item()}"


def test_per_conditional_leaf_with_channel_and_repetition_dims_in_x():
    """Entropy path (rat_spn.py:654-659): leaf samples x [n, w, F, I, R] scored without permutation, per-conditional
    parameters, R % 4 == 0 -> vectorised kernels with strided x."""
    from oracle import spn_oracle as O
    from ratspn_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, w, F, D, I, R = 3, 40, 24, 2, 6, 4
    arch = O.Arch(F=F, D=D, R=R, S=2, I=I, C=1, tanh_squash=False)
    x = torch.randn(n, w, F, I, R, generator=g)
    x[1, 2, 5] = float("nan")
    mu = torch.randn(w, F, I, R, generator=g)
    sigma = torch.rand(w, F, I, R, generator=g) * 0.9 + 0.05
    go = torch.randn(n, w, arch.d0, I, R, generator=g)
    mu_o, sg_o = mu.double().requires_grad_(True), sigma.double().requires_grad_(True)
    ref = O.leaf_product(O.leaf_log_prob(x.double(), mu_o, sg_o, arch), arch)
    (ref * go.double()).sum().backward()
    mu_t, sg_t = mu.to(DEV).requires_grad_(True), sigma.to(DEV).requires_grad_(True)
    out = ops.leaf_forward(x.to(DEV), mu_t, sg_t, None, arch.card, False, "ref")
    (out * go.to(DEV)).sum().backward()
    assert (out.cpu().double() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item() + 1e-5
    for name, a, b in (("dmu", mu_t.grad, mu_o.grad), ("dsigma", sg_t.grad, sg_o.grad)):
        err = (a.cpu().double() - b).abs().max().item()
        assert err <= 2e-5 * b.abs().max().item() + 1e-6, f"{name}: {err}"
