"""Checkpoint interchange with the reference: whole-model pickles written by the UNMODIFIED reference
(th.save(model): rat_spn.py:1046-1047; CSPN.save: cspn.py:324-327; fixtures from oracle/gen_golden_pickle.py) load into
this package once its modules are registered under the reference's flat names, and evaluate to the reference's own
log-likelihoods.  The loading half runs on CPU (no kernels involved); the evaluation half needs the GPU."""
import io
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    import ratspn_b200 as rb
    rb.install_as_reference_modules()
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def test_reference_pickles_load_into_the_mirror_classes():
    import ratspn_b200 as rb
    m = _load("ref_pickle_ratspn.pt")
    assert type(m) is rb.RatSpn and type(m.root) is rb.Sum and type(m._leaf.base_leaf) is rb.RatNormal
    fresh = rb.RatSpn(m.config)
    assert list(fresh.state_dict().keys()) == list(m.state_dict().keys())
    assert all(a.shape == b.shape for a, b in zip(fresh.state_dict().values(), m.state_dict().values()))
    c = _load("ref_pickle_cspn.pt")
    assert type(c) is rb.CSPN
    fresh = rb.CSPN(c.config)
    assert list(fresh.state_dict().keys()) == list(c.state_dict().keys())
    # every host-side attribute of a freshly built model exists on the loaded one (defaults filled in by __setstate__)
    def walk(a, b, prefix=""):
        for k in b.__dict__:
            assert k in a.__dict__, f"loaded model lacks {prefix}{k}"
        for (n1, c1), (_, c2) in zip(a.named_children(), b.named_children()):
            walk(c1, c2, prefix + n1 + ".")
    walk(c, fresh)


def test_own_pickles_round_trip():
    import ratspn_b200 as rb
    m = _load("ref_pickle_ratspn.pt")
    buf = io.BytesIO()
    torch.save(m, buf)            # rat_spn.py:1046-1047 semantics: the whole module is pickled
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert type(m2) is rb.RatSpn
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    c = _load("ref_pickle_cspn.pt")
    buf = io.BytesIO()
    c.save(buf)                   # CSPN.save: a fresh instance carrying only the state_dict
    buf.seek(0)
    c2 = torch.load(buf, weights_only=False)
    for (k, a), (_, b) in zip(c.state_dict().items(), c2.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.gpu
def test_loaded_reference_models_reproduce_the_reference_log_likelihoods():
    io_ = np.load(os.path.join(GOLD, "ref_pickle_io.npz"))
    m = _load("ref_pickle_ratspn.pt").to("cuda")
    ll = m(torch.from_numpy(io_["x"]).to("cuda")).detach().cpu().numpy()
    np.testing.assert_allclose(ll, io_["ll"], rtol=1e-5, atol=1e-4)
    c = _load("ref_pickle_cspn.pt").to("cuda")
    llc = c(torch.from_numpy(io_["xc"]).to("cuda"), condition=torch.from_numpy(io_["cond"]).to("cuda")).detach().cpu().numpy()
    np.testing.assert_allclose(llc, io_["llc"], rtol=1e-5, atol=1e-4)
