"""Drop-in for ``HFModelWrapper.forward`` (skyrl-train/skyrl_train/model_wrapper.py:261-375) on a Hugging Face causal
LM: the transformer trunk runs as before, but the ``lm_head`` GEMM, ``logits.div_(T)``, the flash-attn cross-entropy
and the chunked entropy are replaced by the fused head on the ``num_actions`` response rows of every sequence -
``[B, S, V]`` logits are never built.

    policy = FusedHeadPolicy(AutoModelForCausalLM.from_pretrained(...))       # what HFModelWrapper holds as .model
    action_log_probs, output = policy(sequences, num_actions, attention_mask, temperature=T,
                                      return_output=True, compute_entropy=True, entropy_requires_grad=use_entropy_loss)

Same call signature and return convention as the reference wrapper: ``action_log_probs`` is ``[B, A]`` and
``output["entropy"]`` is ``[B, S]`` so that the worker's own slice ``[:, -A-1:-1]`` (worker.py:855) keeps working
(positions outside the response window hold zeros - the worker never reads them).

FSDP: the wrapper reads ``lm_head.weight`` directly instead of calling ``lm_head.forward``.  With FSDP2
(``fully_shard``) the weight is a ``DTensor``: it is all-gathered with ``full_tensor()`` for the fused head, and autograd
reduce-scatters the ``d_weight`` the head produces back onto the shards.  With FSDP1 (``FullyShardedDataParallel``)
a unit's parameters only exist un-sharded inside that unit's own ``forward``; reading the weight from outside would
see the flat shard, so the wrapper refuses with a clear error unless the weight is whole (``use_orig_params`` views of
a root-unit lm_head inside the root forward, or ``summon_full_params``).  Tied embeddings follow the same rule.

Padded layout, one trunk per GPU.  A trunk that runs on sample-packed or Ulysses-sliced inputs (both need the
reference's attention patches around the HF model) calls the matching glue function of :mod:`skyrl_b200.model_wrapper`
on its own hidden states instead: ``action_log_probs_from_packed_hidden`` / ``_from_sp_hidden`` / ``_from_sp_packed_hidden``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import model_wrapper as _glue


def _whole_weight(weight: torch.Tensor, hidden_size: int) -> torch.Tensor:
    """The un-sharded ``[V, H]`` lm_head weight, or a clear error: the fused head bypasses ``lm_head.forward``, so the
    FSDP hooks that un-shard a parameter for its own module's forward do not fire for it."""
    if hasattr(weight, "full_tensor"):      # DTensor (FSDP2 / TP): gather; the backward reduce-scatters d_weight
        weight = weight.full_tensor()
    if weight.dim() != 2 or weight.shape[1] != hidden_size or weight.numel() == 0:
        raise RuntimeError(
            f"FusedHeadPolicy needs the whole [V, H] lm_head weight but sees a tensor of shape {tuple(weight.shape)} "
            f"(hidden size {hidden_size}): under FSDP1 a wrapped unit's parameters are only un-sharded inside its own "
            "forward.  Keep lm_head (and tied embeddings) in the ROOT FSDP unit with use_orig_params=True, use FSDP2 "
            "(fully_shard: the DTensor weight is gathered here), or call the wrapper under summon_full_params().")
    return weight


class FusedHeadPolicy(nn.Module):
    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model
        trunk = model.get_decoder() if hasattr(model, "get_decoder") else getattr(model, "model", None)
        head = model.get_output_embeddings() if hasattr(model, "get_output_embeddings") else getattr(model, "lm_head", None)
        if trunk is None or head is None or getattr(head, "bias", None) is not None:
            raise ValueError("FusedHeadPolicy needs a causal LM with a decoder trunk and a bias-free lm_head")
        self._trunk, self._head = [trunk], [head]  # lists: not registered twice as sub-modules

    def forward(self, sequences: torch.LongTensor, num_actions, attention_mask: Optional[torch.Tensor] = None,
                temperature: float = 1.0, return_output: bool = False, compute_entropy: bool = False,
                entropy_requires_grad: bool = True):
        if isinstance(num_actions, (list, tuple)):  # model_wrapper.py:363-367
            if len(num_actions) != 1:
                raise ValueError("per-sample num_actions is not supported by the fused head (use the padded batch)")
            num_actions = num_actions[0]
        A = int(num_actions)
        if attention_mask is None:
            attention_mask = torch.ones_like(sequences)
        position_ids = attention_mask.long().cumsum(-1) - 1   # model_wrapper.py:272-273
        position_ids.masked_fill_(attention_mask == 0, 1)
        hidden = self._trunk[0](sequences, attention_mask=attention_mask, position_ids=position_ids).last_hidden_state
        weight = _whole_weight(self._head[0].weight, hidden.shape[-1])
        lp, ent = _glue.action_log_probs_from_hidden(hidden, weight, sequences, A, attention_mask, temperature,
                                                     compute_entropy)
        if not return_output:
            return lp
        output = {}
        if compute_entropy:
            if not entropy_requires_grad:
                ent = ent.detach()
            S = sequences.shape[1]
            output["entropy"] = F.pad(ent, (S - A - 1, 1))    # [B, S]; the worker slices [:, -A-1:-1]
        return lp, output


class ValueHeadCritic(nn.Module):
    """Drop-in for the reference's ``CriticModel.forward`` (model_wrapper.py:425-500, padded layout): the trunk's last
    hidden state goes through the CUDA value head on the ``num_actions`` response rows only, returning what the
    reference returns as ``values[:, :-1][:, -num_actions:]``.

    ``trunk`` is the base model (``AutoModel``), ``value_head`` the reference's ``nn.Linear(hidden_size, 1, bias=False)``
    (a bias is honoured if present)."""

    def __init__(self, trunk: nn.Module, value_head: nn.Linear):
        super().__init__()
        self.trunk, self.value_head = trunk, value_head

    def forward(self, input_ids: torch.LongTensor, num_actions=None, attention_mask: Optional[torch.Tensor] = None,
                return_output: bool = False):
        from .value_head import action_values

        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        position_ids = attention_mask.long().cumsum(-1) - 1
        position_ids.masked_fill_(attention_mask == 0, 1)
        outputs = self.trunk(input_ids, attention_mask=attention_mask, position_ids=position_ids)
        if num_actions is None:
            assert return_output
            return outputs
        if isinstance(num_actions, (list, tuple)):
            if len(num_actions) != 1:
                raise ValueError("per-sample num_actions is not supported (use the padded batch)")
            num_actions = num_actions[0]
        values = action_values(outputs.last_hidden_state, self.value_head.weight, int(num_actions),
                               getattr(self.value_head, "bias", None))
        return (values, outputs) if return_output else values
