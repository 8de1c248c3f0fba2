"""Golden fixture at the REAL shape of BASELINE config 3 (build container only; TEST INFRASTRUCTURE).

    python oracle/gen_golden_cfg3.py

Runs the UNMODIFIED reference CSPN exactly as ``CspnActor.__init__`` builds it for Humanoid (sb3.py:89-109, CLI defaults
joint_fail_sac_sb3.py:440-444): F_cond = (376,), F = 17 actions, R = I = S = 3, D = int(log2 17) = 4, feat / sum / dist
hidden layers [64], LeakyReLU, tanh squashing, sigma in [1e-3, 1], on a 16-conditional slice of the 256-row policy batch
(the circuit is independent per conditional; the recorded entropy
noise is what makes the fixture 4 MiB).  Stored in tests/golden/cspn_cfg3.npz:

  * log-likelihood of a batch of actions + gradients w.r.t. every gating-network parameter         (sb3.py:336-344)
  * rollout: index sampling with evidence                                                           (sb3.py:196-203)
  * policy update: onehot sample with evidence + recursive entropy (n = 5, marginal mask) + gradients (sb3.py:146-184)

together with the Gumbel / normal noise the reference drew, in draw order.  The oracle's deviation from the reference on
every stored quantity is appended to tests/golden/PIN_REPORT.txt.  Needs the one-line onehot / leaf-padding patch of
oracle/ref_shim.py (the reference crashes at distributions.py:390 for every action dim that is not a power of two).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import spn_oracle as O  # noqa: E402
from oracle.gen_golden import GOLD, REPORT, grads_of, noise_arrays, note, save, sd_arrays, zero_grads  # noqa: E402
from oracle.ref_shim import load_reference, recorded_noise  # noqa: E402

CASE = "cspn_cfg3"
B, F_COND, A = 16, 376, 17


def main():
    ref_rat_spn, ref_cspn, ref_layers, ref_dist, ref_utils = load_reference()
    CSPN, CspnConfig = ref_cspn.CSPN, ref_cspn.CspnConfig
    torch.manual_seed(33)
    np.random.seed(33)
    c = CspnConfig()
    c.F_cond = (F_COND,)
    c.C = 1
    c.feat_layers = [64]
    c.sum_param_layers = [64]
    c.dist_param_layers = [64]
    c.F, c.R, c.D, c.I, c.S = A, 3, int(np.log2(A)), 3, 3
    c.dropout = 0.0
    c.leaf_base_class = ref_dist.RatNormal
    c.tanh_squash = True
    c.cond_layers_inner_act = torch.nn.LeakyReLU
    c.leaf_base_kwargs = {"min_sigma": 0.001, "max_sigma": 1.0}
    m = CSPN(c)
    arch = O.Arch(F=A, D=c.D, R=3, S=3, I=3, C=1, tanh_squash=True)
    arrays = dict(sd_arrays(m))
    arrays["meta"] = np.array([A, c.D, 3, 3, 3, 1, 1, F_COND, 64])
    obs = torch.randn(B, F_COND)
    failed = torch.rand(B, A) < 0.05
    failed[0, 3] = failed[5, 16] = True          # at least two failed joints whatever the seed
    ev = torch.atanh(torch.rand(1, B, A, 1, 1) * 1.8 - 0.9)
    ev[~failed.view(1, B, A, 1, 1)] = float("nan")
    x = torch.atanh(torch.rand(1, B, A, 1, 1) * 1.8 - 0.9)
    arrays.update(cond=obs.numpy(), failed=failed.numpy(), evidence=ev.numpy(), x=x.numpy())

    def oracle_params(requires_grad=False):
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        raw = {}
        if requires_grad:
            for k, p in m.named_parameters():
                if p.requires_grad:
                    sd[k] = sd[k].requires_grad_(True)
                    raw[k] = sd[k]
        return O.cspn_params(sd, arch, obs, min_sigma=0.001, max_sigma=1.0), raw

    # ---- log-likelihood of replayed actions + gradients (actor/critic losses evaluate log pi(a|s))
    zero_grads(m)
    ll = m(x, condition=obs)
    (-ll.mean()).backward()
    arrays["ll"] = ll.detach().numpy()
    g = grads_of(m)
    arrays.update({f"ll__{k}": v for k, v in g.items()})
    Pc, raw = oracle_params(True)
    o = O.forward(Pc, arch, x)
    (-o.mean()).backward()
    note(CASE, "forward", ll.detach(), o.detach())
    for k, t in raw.items():
        note(CASE, f"ll grad {k}", g["grad__" + k], t.grad)

    # ---- rollout: index sampling with evidence, and its MPE variant (deterministic actions at evaluation time)
    for tag, mpe, seed in (("ixev", False, 34), ("ixevmpe", True, 35)):
        rec = []
        torch.manual_seed(seed)
        with torch.no_grad(), recorded_noise(ref_layers, ref_dist, rec):
            ctx = m.sample(mode="index", condition=obs, evidence=ev, is_mpe=mpe)
        arrays.update(noise_arrays(tag, rec))
        arrays[f"{tag}__sample"] = ctx.sample.numpy()
        Pc0, _ = oracle_params(False)
        o = O.sample_index(Pc0, arch, O.NoiseSource(replay=rec), n=1, evidence=ev, is_mpe=mpe)
        note(CASE, f"{tag}: index sample + evidence", ctx.sample, o.sample)

    # ---- policy update: onehot sample with evidence + recursive entropy with the failed-joint mask, one backward
    rec = []
    torch.manual_seed(36)
    zero_grads(m)
    with recorded_noise(ref_layers, ref_dist, rec):
        ctx = m.sample(mode="onehot", condition=obs, evidence=ev)
        H, _ = m.recursive_entropy_approx(condition=obs, sample_size=5, marginal_mask=failed)
    act = ctx.sample.squeeze(0).squeeze(0).squeeze(-1)
    (-(0.1 * H + act.sum(-1)).mean()).backward()
    arrays.update(noise_arrays("sac", rec))
    arrays["sac__sample"] = ctx.sample.detach().numpy()
    arrays["sac__H"] = H.detach().numpy()
    g = grads_of(m)
    arrays.update({f"sac__{k}": v for k, v in g.items()})
    Pc, raw = oracle_params(True)
    ns = O.NoiseSource(replay=rec)
    os_ = O.sample_onehot(Pc, arch, ns, n=1, evidence=ev)
    oH = O.recursive_entropy(Pc, arch, ns, sample_size=5, marginal_mask=failed, differentiable=True)
    (-(0.1 * oH + os_.sample.squeeze(0).squeeze(0).squeeze(-1).sum(-1)).mean()).backward()
    note(CASE, "SAC sample", ctx.sample.detach(), os_.sample.detach())
    note(CASE, "SAC entropy", H.detach(), oH.detach())
    for k, t in raw.items():
        note(CASE, f"SAC grad {k}", g["grad__" + k], t.grad)

    # noise tensors dominate the fixture: half precision storage would change the draws, so keep fp32 but compressed
    save(CASE, arrays)
    with open(os.path.join(GOLD, "PIN_REPORT.txt"), "a") as f:
        f.write("\noracle/gen_golden_cfg3.py (BASELINE config 3 at its real shape: obs 376, 17 actions, R=I=S=3, D=4, [64] "
                f"hidden, {B}-conditional slice)\n")
        f.write("\n".join(REPORT) + "\n")


if __name__ == "__main__":
    main()
