"""SAC actor around the CSPN: mirror of ``sac_rl_experiments/sb3.py:29-203`` (``CspnActor``), the caller of the hot path in
the reference's RL experiments.

With ``stable_baselines3`` installed the class derives from its ``BasePolicy`` exactly like the reference, so
``CspnPolicy.make_actor`` (sb3.py:215-220) can construct it unchanged.  Without it (this image has neither SB3 nor gym) a
small stand-in base keeps the same constructor arguments and ``extract_features`` contract, which is all the actor uses of
``BasePolicy``; spaces then only need a ``.shape``.

What the adapter does per call (sb3.py:125-203):
  * ``extract_features``: optionally splits the trailing ``2 * action_dim`` joint-failure columns off the observation
    (first half: 0/1 failure flags, second half: the squashed action values of the failed joints);
  * ``_sample``: evidence = atanh of the failed joints' values (NaN = to be sampled), differentiable ``onehot`` sampling when
    gradients are enabled and ``index`` sampling otherwise (rollouts), MPE when ``deterministic``;
  * ``action_entropy``: the sample plus one of the entropy objectives (``recursive`` / ``naive`` / ``huber`` /
    ``augmented_huber``) with the failure flags as marginalisation mask.
One deviation, value-neutral: sample and entropy of one call run inside ``CSPN.conditioned(features)``, so the gating
network is evaluated once instead of twice (tests/test_gpu_parity.py::test_cspn_conditioned_block_... pins the equality).
"""
from typing import Any, Dict, List, Optional, Tuple, Type

import numpy as np
import torch as th
from torch import nn

from .cspn import CSPN, CspnConfig
from .distributions import RatNormal

try:  # pragma: no cover - not installed in this image
    from stable_baselines3.common.policies import BasePolicy as _PolicyBase
    from stable_baselines3.common.preprocessing import get_action_dim
    HAVE_SB3 = True
except Exception:  # ImportError, or a broken optional dependency of SB3
    HAVE_SB3 = False

    def get_action_dim(action_space) -> int:
        """Box action spaces only (what SAC supports): the number of action dimensions."""
        return int(np.prod(action_space.shape))

    class _PolicyBase(nn.Module):
        """The part of ``stable_baselines3.common.policies.BasePolicy`` the actor relies on."""

        def __init__(self, observation_space, action_space, features_extractor: Optional[nn.Module] = None,
                     normalize_images: bool = True, squash_output: bool = False, **_unused):
            super().__init__()
            self.observation_space = observation_space
            self.action_space = action_space
            self.features_extractor = features_extractor
            self.normalize_images = normalize_images
            self._squash_output = squash_output

        @property
        def squash_output(self) -> bool:
            return self._squash_output

        def extract_features(self, obs: th.Tensor) -> th.Tensor:
            obs = obs.float()
            if self.features_extractor is None:
                return obs.flatten(1)
            return self.features_extractor(obs)

        def set_training_mode(self, mode: bool) -> None:
            self.train(mode)

        def _get_constructor_parameters(self) -> Dict[str, Any]:
            return dict(observation_space=self.observation_space, action_space=self.action_space,
                        normalize_images=self.normalize_images)

ENTROPY_OBJECTIVES = ("recursive", "naive", "huber", "augmented_huber")


class CspnActor(_PolicyBase):
    """Actor network (policy) for SAC whose action distribution is a CSPN conditioned on the observation features;
    constructor arguments as in sb3.py:43-69."""

    def __init__(self, observation_space, action_space, net_arch: List[int], features_extractor: Optional[nn.Module],
                 features_dim: int, R: int, D: Optional[int], I: int, S: int, dropout: float, feat_layers,
                 sum_param_layers, dist_param_layers, entropy_objective: str, recurs_ent_approx_sample_size: int,
                 naive_ent_approx_sample_size: int, joint_failure_info_in_obs: bool,
                 cond_layers_inner_act: Type[nn.Module] = nn.ReLU, squash_output: bool = True, min_std: float = 0.001,
                 max_std: float = 1.0, activation_fn: Type[nn.Module] = nn.ReLU, normalize_images: bool = True, **kwargs):
        super().__init__(observation_space, action_space, features_extractor=features_extractor,
                         normalize_images=normalize_images, squash_output=True)
        action_dim = get_action_dim(self.action_space)
        if joint_failure_info_in_obs:
            features_dim -= action_dim * 2
        self.net_arch = net_arch
        self.features_dim = features_dim
        self.recurs_ent_approx_sample_size = recurs_ent_approx_sample_size
        self.naive_ent_approx_sample_size = naive_ent_approx_sample_size
        self.entropy_objective = entropy_objective
        self.joint_failure_info_in_obs = joint_failure_info_in_obs
        self.action_dim = action_dim

        config = CspnConfig()
        config.F_cond = (features_dim,)
        config.C = 1
        config.feat_layers = feat_layers
        config.sum_param_layers = sum_param_layers
        config.dist_param_layers = dist_param_layers
        config.F = action_dim
        config.R = R
        config.D = D if D is not None else int(np.log2(action_dim))
        config.I = I
        config.S = S
        config.dropout = dropout
        config.leaf_base_class = RatNormal
        config.tanh_squash = squash_output
        config.cond_layers_inner_act = cond_layers_inner_act
        config.leaf_base_kwargs = {'min_sigma': min_std, 'max_sigma': max_std}
        self.config = config
        self.cspn = CSPN(config)

    def _get_constructor_parameters(self) -> Dict[str, Any]:
        # the reference's version (sb3.py:113-123) reads attributes the class never sets; this one lists what __init__ takes
        data = super()._get_constructor_parameters()
        cfg = self.config
        data.update(dict(
            net_arch=self.net_arch, features_extractor=self.features_extractor,
            features_dim=self.features_dim + (2 * self.action_dim if self.joint_failure_info_in_obs else 0),
            R=cfg.R, D=cfg.D, I=cfg.I, S=cfg.S, dropout=cfg.dropout, feat_layers=cfg.feat_layers,
            sum_param_layers=cfg.sum_param_layers, dist_param_layers=cfg.dist_param_layers,
            entropy_objective=self.entropy_objective, recurs_ent_approx_sample_size=self.recurs_ent_approx_sample_size,
            naive_ent_approx_sample_size=self.naive_ent_approx_sample_size,
            joint_failure_info_in_obs=self.joint_failure_info_in_obs, cond_layers_inner_act=cfg.cond_layers_inner_act,
            squash_output=cfg.tanh_squash, min_std=cfg.leaf_base_kwargs['min_sigma'],
            max_std=cfg.leaf_base_kwargs['max_sigma']))
        return data

    # ------------------------------------------------------------------------------------------ pieces of one call
    def extract_features(self, obs: th.Tensor) -> Tuple[th.Tensor, Optional[th.Tensor]]:
        """obs [B, features (+ 2 * action_dim)] -> (features [B, features_dim], joint failure info [B, 2 * action_dim] | None)
        (sb3.py:186-194)."""
        joint_failure_info = None
        if self.joint_failure_info_in_obs:
            split = obs.shape[1] - 2 * self.action_dim
            obs, joint_failure_info = obs[:, :split], obs[:, split:]
        obs = super().extract_features(obs)
        if joint_failure_info is not None:
            joint_failure_info = joint_failure_info.to(dtype=obs.dtype)
        return obs, joint_failure_info

    def evidence_from_failure_info(self, joint_failure_info: th.Tensor) -> th.Tensor:
        """[B, 2 A] (flags | squashed values) -> evidence [1, B, A, 1, 1] in the leaves' unsquashed support: atanh of the
        failed joints' values, NaN (= sample it) for the healthy ones (sb3.py:127-132)."""
        info = joint_failure_info.clone().to(dtype=self.cspn.dtype)
        failed_joints, evidence = info[:, :self.action_dim], info[:, self.action_dim:].clone()
        evidence[failed_joints == 0.0] = float("nan")
        return evidence.unsqueeze(0).unsqueeze(-1).unsqueeze(-1).atanh()

    def _sample(self, condition: th.Tensor, is_mpe: bool, joint_failure_info: Optional[th.Tensor]) -> th.Tensor:
        evidence = None
        if joint_failure_info is not None:
            evidence = self.evidence_from_failure_info(joint_failure_info)
        action = self.cspn.sample(mode='onehot' if th.is_grad_enabled() else 'index', condition=condition,
                                  is_mpe=is_mpe, evidence=evidence).sample
        if joint_failure_info is not None:
            action = action.squeeze(0)          # the evidence's extra batch dimension
        # [nodes = 1, samples per node = 1, conditionals, action_dim, 1] -> [conditionals, action_dim]
        return action.squeeze(0).squeeze(0).squeeze(-1)

    def _entropy(self, features: th.Tensor, failed_joints: Optional[th.Tensor], log_ent_metrics: bool):
        obj = self.entropy_objective
        if obj == 'recursive':
            return self.cspn.recursive_entropy_approx(condition=features, verbose=log_ent_metrics,
                                                      sample_size=self.recurs_ent_approx_sample_size,
                                                      marginal_mask=failed_joints)
        if obj == 'naive':
            return self.cspn.naive_entropy_approx(condition=features, sample_size=self.naive_ent_approx_sample_size,
                                                  marginal_mask=failed_joints), {}
        if obj == 'huber':
            return self.cspn.huber_entropy_lb(condition=features, verbose=log_ent_metrics, detach_weights=False,
                                              add_sub_weight_ent=False, marginal_mask=failed_joints)
        if obj == 'augmented_huber':
            return self.cspn.huber_entropy_lb(condition=features, verbose=log_ent_metrics, detach_weights=True,
                                              add_sub_weight_ent=True, marginal_mask=failed_joints)
        raise ValueError(f"entropy_objective {obj} unknown!")

    # ------------------------------------------------------------------------------------------ public calls
    def action_entropy(self, observation: th.Tensor, deterministic: bool = False, log_ent_metrics: bool = False,
                       compute_entropy: bool = False) -> Tuple[th.Tensor, Optional[th.Tensor], Optional[dict]]:
        """Action [B, action_dim] (tanh-squashed when the leaves are), entropy estimate [B] or None, logging dict
        (sb3.py:146-184)."""
        if compute_entropy and self.entropy_objective not in ENTROPY_OBJECTIVES:
            raise ValueError(f"entropy_objective {self.entropy_objective} unknown!")
        features, joint_failure_info = self.extract_features(observation)
        failed_joints = joint_failure_info[:, :self.action_dim] if joint_failure_info is not None else None
        entropy = ent_log = None
        with self.cspn.conditioned(features):   # one pass of the gating network for the sample and the entropy
            action = self._sample(condition=features, is_mpe=deterministic, joint_failure_info=joint_failure_info)
            if compute_entropy:
                entropy, ent_log = self._entropy(features, failed_joints, log_ent_metrics)
        return action, entropy, ent_log

    def forward(self, obs: th.Tensor, deterministic: bool = False) -> th.Tensor:
        action, _, _ = self.action_entropy(observation=obs, deterministic=deterministic, log_ent_metrics=False,
                                           compute_entropy=False)
        return action

    def _predict(self, observation: th.Tensor, deterministic: bool = False) -> th.Tensor:
        return self(observation, deterministic)
