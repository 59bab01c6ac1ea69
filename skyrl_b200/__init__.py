"""skyrl_b200 — B200-native policy-update hot path behind skyrl_train's plugin API.

Public surface (same names as the reference where one exists):

* registries: ``AdvantageEstimatorRegistry``, ``PolicyLossRegistry``, ``register_advantage_estimator``,
  ``register_policy_loss``, ``repopulate_all_registries``, ``install_into_skyrl_train``
* advantages: ``compute_grpo_outcome_advantage``, ``compute_rloo_outcome_advantage``,
  ``compute_gae_advantage_return``, ``compute_advantages_and_returns``, ``compute_approx_kl``,
  ``apply_reward_kl_penalty``
* losses: ``ppo_policy_loss``, ``gspo_policy_loss``, ``policy_update_loss``
* logits helpers: ``logprobs_from_logits``, ``chunked_entropy_from_logits``, ``masked_mean``
* fused head: ``fused_lm_head``, ``fused_lm_head_logprobs``
* critic / TP: ``value_head``, ``action_values``, ``vocab_parallel_lm_head``; trainer-side batch helpers in
  ``skyrl_b200.trainer_ops``
* the step: ``policy_update_step`` / ``PolicyUpdateConfig`` (skyrl_b200.policy_update)

All arithmetic runs in ``libskyb200.so`` (hand-written sm_100a CUDA, C ABI in include/skyb200.h).  Importing this
package does not need a GPU; calling any op without the built library or without a CUDA device raises.
"""
from .advantages import (  # noqa: F401
    apply_reward_kl_penalty,
    compute_advantages_and_returns,
    compute_approx_kl,
    compute_gae_advantage_return,
    compute_grpo_outcome_advantage,
    compute_reinforce_plus_plus_outcome_advantage,
    normalize_advantages,
    normalize_advantages_dict,
    compute_rloo_outcome_advantage,
)
from .fused_head import fused_lm_head, fused_lm_head_logprobs  # noqa: F401
from .hf_wrapper import FusedHeadPolicy, ValueHeadCritic  # noqa: F401
from .losses import gspo_policy_loss, policy_update_loss, ppo_critic_loss, ppo_policy_loss  # noqa: F401
from .model_wrapper import (  # noqa: F401
    action_log_probs_from_hidden,
    action_log_probs_from_packed_hidden,
    action_log_probs_from_sp_hidden,
    action_log_probs_from_sp_packed_hidden,
)
from .policy_update import (PolicyUpdateOutput, fused_policy_loss, pack_by_lengths, policy_update_step,  # noqa: F401
                            step_gemm_times)
from .registry import (  # noqa: F401
    AdvantageEstimator,
    AdvantageEstimatorRegistry,
    PolicyLossRegistry,
    PolicyLossType,
    install_into_skyrl_train,
    register_advantage_estimator,
    register_policy_loss,
    repopulate_all_registries,
)
from .tp_head import vocab_parallel_lm_head  # noqa: F401
from .value_head import action_values, value_head  # noqa: F401
from .torch_utils import (  # noqa: F401
    chunked_entropy_from_logits,
    logprobs_from_logits,
    logprobs_from_logits_v2,
    masked_mean,
    safe_exp_delta,
)

repopulate_all_registries()

__version__ = "0.1.0"
