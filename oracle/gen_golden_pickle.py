"""Golden fixture for checkpoint interchange: whole-model pickles written by the UNMODIFIED reference
(th.save(model), as rat_spn.py:1046-1047 / cspn.py:324-327 do) plus inputs and the reference's log-likelihoods.

TEST INFRASTRUCTURE, build container only (imports /root/reference through oracle/ref_shim.py).  Writes
tests/golden/ref_pickle_ratspn.pt, ref_pickle_cspn.pt and ref_pickle_io.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    ref_rat_spn, ref_cspn, ref_layers, ref_dist, ref_utils = load_reference()
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = ref_rat_spn.RatSpnConfig()
    cfg.F, cfg.D, cfg.R, cfg.S, cfg.I, cfg.C = 16, 3, 3, 4, 5, 2
    cfg.dropout, cfg.leaf_base_class, cfg.leaf_base_kwargs, cfg.tanh_squash = 0.0, ref_dist.RatNormal, {}, False
    m = ref_rat_spn.RatSpn(cfg)
    x = torch.randn(6, 1, 16, 1, 1)
    with torch.no_grad():
        ll = m(x)
    torch.save(m, os.path.join(OUT, "ref_pickle_ratspn.pt"))

    ccfg = ref_cspn.CspnConfig()
    ccfg.F_cond = (5,)
    ccfg.F, ccfg.R, ccfg.D, ccfg.I, ccfg.S, ccfg.C = 8, 2, 2, 3, 3, 1
    ccfg.dropout = 0.0
    ccfg.leaf_base_class = ref_dist.RatNormal
    ccfg.leaf_base_kwargs = {"min_sigma": 0.05, "max_sigma": 1.0}
    ccfg.tanh_squash = True
    ccfg.feat_layers, ccfg.sum_param_layers, ccfg.dist_param_layers = [6], [6], [6]
    ccfg.cond_layers_inner_act = torch.nn.LeakyReLU
    c = ref_cspn.CSPN(ccfg)
    cond = torch.randn(4, 5)
    xc = torch.rand(3, 4, 8, 1, 1) * 1.6 - 0.8
    with torch.no_grad():
        llc = c(xc, condition=cond)
    c.save(os.path.join(OUT, "ref_pickle_cspn.pt"))   # cspn.py:324-327
    np.savez(os.path.join(OUT, "ref_pickle_io.npz"), x=x.numpy(), ll=ll.numpy(), cond=cond.numpy(), xc=xc.numpy(),
             llc=llc.numpy())
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
