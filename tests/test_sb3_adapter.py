"""SAC actor adapter (ratspn_b200.sb3_adapter.CspnActor, mirror of sac_rl_experiments/sb3.py:29-203).

CPU tests: the host logic around the CSPN calls (observation split, evidence construction, sampling-mode choice, entropy
objective dispatch) with the CSPN replaced by a recorder.  GPU test: the real actor against direct CSPN calls."""
import math
from contextlib import contextmanager
from types import SimpleNamespace

import pytest
import torch

A, FEAT = 6, 10


class _Space:
    def __init__(self, *shape):
        self.shape = tuple(shape)


def _actor(objective="recursive", failure_info=True, **kw):
    from ratspn_b200.sb3_adapter import CspnActor
    args = dict(observation_space=_Space(FEAT + (2 * A if failure_info else 0)), action_space=_Space(A), net_arch=[],
                features_extractor=None, features_dim=FEAT + (2 * A if failure_info else 0), R=3, D=2, I=3, S=3, dropout=0.0,
                feat_layers=[8], sum_param_layers=[8], dist_param_layers=[8], entropy_objective=objective,
                recurs_ent_approx_sample_size=4, naive_ent_approx_sample_size=7, joint_failure_info_in_obs=failure_info)
    args.update(kw)
    return CspnActor(**args)


class _Recorder(torch.nn.Module):
    """Stands in for the CSPN: records the calls, returns tensors of the documented shapes."""
    dtype = torch.float32

    def __init__(self):
        super().__init__()
        self.calls = []

    @contextmanager
    def conditioned(self, condition):
        self.calls.append(("conditioned", condition))
        yield self

    def sample(self, mode=None, condition=None, is_mpe=False, evidence=None):
        self.calls.append(("sample", dict(mode=mode, condition=condition, is_mpe=is_mpe, evidence=evidence)))
        w = condition.shape[0]
        s = torch.arange(w * A, dtype=torch.float32).reshape(1, 1, w, A, 1)
        if evidence is not None:
            s = s.unsqueeze(0)
        return SimpleNamespace(sample=s)

    def recursive_entropy_approx(self, **kw):
        self.calls.append(("recursive", kw))
        return torch.zeros(kw["condition"].shape[0]), {"r": 1}

    def naive_entropy_approx(self, **kw):
        self.calls.append(("naive", kw))
        return torch.ones(kw["condition"].shape[0])

    def huber_entropy_lb(self, **kw):
        self.calls.append(("huber", kw))
        return torch.full((kw["condition"].shape[0],), 2.0), {"h": 1}


def _obs(B=5):
    g = torch.Generator().manual_seed(0)
    feats = torch.randn(B, FEAT, generator=g)
    failed = (torch.rand(B, A, generator=g) < 0.4).float()
    values = torch.rand(B, A, generator=g) * 1.6 - 0.8
    return feats, failed, values, torch.cat([feats, failed, values], dim=1)


def test_constructor_builds_the_cspn_of_the_reference_actor():
    import ratspn_b200 as rb
    actor = _actor()
    cfg = actor.cspn.config
    assert isinstance(actor.cspn, rb.CSPN)
    assert (cfg.F, cfg.F_cond, cfg.C, cfg.R, cfg.D, cfg.I, cfg.S) == (A, (FEAT,), 1, 3, 2, 3, 3)   # failure columns are not features
    assert cfg.tanh_squash and cfg.leaf_base_kwargs == {"min_sigma": 0.001, "max_sigma": 1.0}
    assert _actor(D=None, failure_info=False).cspn.config.D == int(math.log2(A))                   # sb3.py:101
    params = actor._get_constructor_parameters()
    assert params["features_dim"] == FEAT + 2 * A and params["entropy_objective"] == "recursive"


def test_observation_split_and_evidence_construction():
    actor = _actor()
    feats, failed, values, obs = _obs()
    f, info = actor.extract_features(obs)
    assert torch.equal(f, feats) and torch.equal(info, torch.cat([failed, values], 1))
    ev = actor.evidence_from_failure_info(info)
    assert ev.shape == (1, 5, A, 1, 1)
    flat = ev[0, :, :, 0, 0]
    assert torch.equal(torch.isnan(flat), failed == 0)                                  # healthy joints are sampled
    assert torch.allclose(flat[failed == 1], torch.atanh(values[failed == 1]))          # failed joints: unsquashed evidence
    f2, info2 = _actor(failure_info=False).extract_features(feats)
    assert torch.equal(f2, feats) and info2 is None


def test_sampling_mode_follows_grad_mode_and_shapes_are_squeezed():
    actor = _actor()
    actor.cspn = rec = _Recorder()
    _, failed, _, obs = _obs()
    a = actor(obs)                                   # gradients enabled: differentiable sampling
    assert a.shape == (5, A)
    kind, kw = rec.calls[1]
    assert rec.calls[0][0] == "conditioned" and kind == "sample" and kw["mode"] == "onehot" and not kw["is_mpe"]
    assert kw["condition"] is rec.calls[0][1] and kw["evidence"].shape == (1, 5, A, 1, 1)
    rec.calls.clear()
    with torch.no_grad():
        a = actor._predict(obs, deterministic=True)  # rollout: index sampling, MPE
    assert a.shape == (5, A) and rec.calls[1][1]["mode"] == "index" and rec.calls[1][1]["is_mpe"]
    plain = _actor(failure_info=False)
    plain.cspn = rec2 = _Recorder()
    assert plain(torch.randn(3, FEAT)).shape == (3, A) and rec2.calls[1][1]["evidence"] is None


@pytest.mark.parametrize("objective,call,expect", [
    ("recursive", "recursive", dict(sample_size=4, verbose=True)),
    ("naive", "naive", dict(sample_size=7)),
    ("huber", "huber", dict(detach_weights=False, add_sub_weight_ent=False, verbose=True)),
    ("augmented_huber", "huber", dict(detach_weights=True, add_sub_weight_ent=True, verbose=True)),
])
def test_entropy_objective_dispatch(objective, call, expect):
    actor = _actor(objective)
    actor.cspn = rec = _Recorder()
    _, failed, _, obs = _obs()
    action, ent, log = actor.action_entropy(obs, log_ent_metrics=True, compute_entropy=True)
    assert action.shape == (5, A) and ent.shape == (5,) and isinstance(log, dict)
    kind, kw = rec.calls[-1]
    assert kind == call and torch.equal(kw["marginal_mask"], failed) and kw["condition"] is rec.calls[0][1]
    for k, v in expect.items():
        assert kw[k] == v, (k, kw[k])
    rec.calls.clear()
    _, ent, log = actor.action_entropy(obs)          # compute_entropy defaults to False: sample only
    assert ent is None and log is None and [c[0] for c in rec.calls] == ["conditioned", "sample"]


def test_unknown_entropy_objective_raises():
    actor = _actor("made_up")
    actor.cspn = _Recorder()
    with pytest.raises(ValueError):
        actor.action_entropy(_obs()[3], compute_entropy=True)


@pytest.mark.gpu
def test_actor_on_gpu_matches_direct_cspn_calls():
    torch.manual_seed(0)
    actor = _actor("recursive").to("cuda")
    _, failed, _, obs = _obs(B=9)
    obs, failed = obs.cuda(), failed.cuda()
    feats, info = actor.extract_features(obs)
    ev = actor.evidence_from_failure_info(info)
    with torch.no_grad():
        got = actor(obs, deterministic=True)
        ref = actor.cspn.sample(mode="index", condition=feats, is_mpe=True, evidence=ev).sample
    assert got.shape == (9, A) and torch.equal(got, ref.squeeze(0).squeeze(0).squeeze(0).squeeze(-1))
    # failed joints carry their evidence (squashed back), healthy ones are inside the tanh range
    assert torch.allclose(got[failed == 1], info[:, A:][failed == 1], atol=1e-5) and got.abs().max() <= 1
    # policy-update pattern: differentiable sample + entropy, one backward reaches the gating network
    for objective in ("recursive", "huber"):
        actor.entropy_objective = objective
        actor.zero_grad(set_to_none=True)
        action, ent, _ = actor.action_entropy(obs, compute_entropy=True)
        # one entropy per conditional (the reference keeps the root's [w, 1, 1] shape)
        assert action.shape == (9, A) and ent.numel() == 9 and torch.isfinite(ent).all() and torch.isfinite(action).all()
        (-(0.2 * ent.reshape(9) + action.sum(-1))).mean().backward()
        grads = [p.grad for p in actor.cspn.parameters() if p.grad is not None]
        assert grads and all(torch.isfinite(g).all() for g in grads) and any(g.abs().max() > 0 for g in grads)
