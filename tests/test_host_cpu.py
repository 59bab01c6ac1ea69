"""Host-side logic that needs no GPU: registries (same behaviour as the reference's), uid grouping, sharding,
config validation, and the loud failure when no CUDA device / library is present."""
import numpy as np
import pytest
import torch

import skyrl_b200
from skyrl_b200 import advantages, dist as sdist, losses
from skyrl_b200.registry import (AdvantageEstimator, AdvantageEstimatorRegistry, PolicyLossRegistry,
                                 register_advantage_estimator, register_policy_loss)
from oracle import policy_oracle as O


def test_builtin_names_registered():
    assert set(AdvantageEstimatorRegistry.list_available()) >= {"grpo", "rloo", "gae"}
    assert set(PolicyLossRegistry.list_available()) >= {"regular", "dual_clip", "gspo", "sapo", "cispo", "clip_cov",
                                                        "kl_cov", "cross_entropy", "importance_sampling"}
    assert set(AdvantageEstimatorRegistry.list_available()) >= {"reinforce++"}
    assert PolicyLossRegistry.get("regular") is losses.ppo_policy_loss
    assert PolicyLossRegistry.get("dual_clip") is losses.ppo_policy_loss
    assert AdvantageEstimatorRegistry.get(AdvantageEstimator.GRPO) is advantages.compute_grpo_outcome_advantage


def test_register_get_unregister_roundtrip():  # reference: tests/cpu/utils/test_ppo_utils.py:287-403
    def fn(**kw):
        return 1

    AdvantageEstimatorRegistry.register("unit_test_est", fn)
    try:
        assert AdvantageEstimatorRegistry.get("unit_test_est") is fn
        assert "unit_test_est" in AdvantageEstimatorRegistry.list_available()
        with pytest.raises(ValueError, match="already registered"):
            AdvantageEstimatorRegistry.register("unit_test_est", fn)
    finally:
        AdvantageEstimatorRegistry.unregister("unit_test_est")
    with pytest.raises(ValueError, match="Unknown advantage estimator"):
        AdvantageEstimatorRegistry.get("unit_test_est")
    with pytest.raises(ValueError, match="not registered"):
        AdvantageEstimatorRegistry.unregister("unit_test_est")
    with pytest.raises(ValueError, match="Unknown policy loss"):
        PolicyLossRegistry.get("nope")


def test_decorators_register():
    @register_policy_loss("unit_test_loss")
    def my_loss(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None):
        return log_probs.sum(), {}

    @register_advantage_estimator("unit_test_adv")
    def my_adv(**kw):
        return 0, 0

    try:
        assert PolicyLossRegistry.get("unit_test_loss").__name__ == "my_loss"
        assert AdvantageEstimatorRegistry.get("unit_test_adv")() == (0, 0)
    finally:
        PolicyLossRegistry.unregister("unit_test_loss")
        AdvantageEstimatorRegistry.unregister("unit_test_adv")


def test_repopulate_is_idempotent():
    before = list(PolicyLossRegistry.list_available())
    skyrl_b200.repopulate_all_registries()
    assert list(PolicyLossRegistry.list_available()) == before


def test_registered_functions_pickle_by_reference():
    import pickle

    for fn in (losses.ppo_policy_loss, losses.gspo_policy_loss, advantages.compute_grpo_outcome_advantage,
               advantages.compute_gae_advantage_return):
        assert pickle.loads(pickle.dumps(fn)) is fn


def test_group_index_matches_dict_semantics():
    uids = np.array(["b", "a", "b", "c", "a", "b"])
    gor, gstart, grows, ng = advantages.build_group_index(uids)
    assert ng == 3
    assert gor.tolist() == [0, 1, 0, 2, 1, 0]  # dense ids in order of first appearance
    assert gstart.tolist() == [0, 3, 5, 6]
    assert grows.tolist() == [0, 2, 5, 1, 4, 3]  # stable inside a group
    # arbitrary hashables, python list input, numpy ints equal to python ints
    gor2, _, _, ng2 = advantages.build_group_index([np.int64(7), 7, (1, 2), (1, 2), "7"])
    assert ng2 == 3 and gor2.tolist() == [0, 0, 1, 1, 2]
    # group assignment agrees bit-exactly with the oracle's
    gid, n = O._group_ids(uids)
    assert n == ng and gid.tolist() == gor.tolist()


def test_shard_prompt_groups_keeps_groups_whole_and_balanced():
    rng = np.random.default_rng(0)
    uids = np.repeat([f"p{i}" for i in range(32)], 16)
    rng.shuffle(uids)
    lens = rng.integers(128, 16384, size=len(uids))
    for ws in (2, 4, 8):
        rows = sdist.shard_prompt_groups(uids, lens, ws)
        flat = sorted(i for r in rows for i in r)
        assert flat == list(range(len(uids)))
        owner = {}
        for r, rr in enumerate(rows):
            for i in rr:
                assert owner.setdefault(uids[i], r) == r, "a prompt group was split across ranks"
        loads = [int(lens[r].sum()) for r in rows]
        assert max(loads) <= 1.15 * (sum(loads) / ws)


def test_sharded_group_advantages_equal_the_whole_batch():
    """N-GPU == 1-GPU for the advantage stage: with whole prompt groups per rank (shard_prompt_groups) the GRPO / RLOO
    advantages each rank computes from ITS rows are the rows of the whole-batch result - no collective needed.
    (Arithmetic by the oracle; what is under test is the partition.)  A row-contiguous split, the reference's
    dispatch (dispatch.py:186-205), does not have that property for shuffled uids."""
    from oracle import policy_oracle as O

    g = torch.Generator().manual_seed(0)
    B, A, G = 48, 20, 6
    uids = np.array([f"p{i // G}" for i in range(B)])[torch.randperm(B, generator=g).numpy()]
    lens = torch.randint(1, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    r = torch.zeros(B, A)
    r[torch.arange(B), lens - 1] = torch.randn(B, generator=g)
    for fn in (O.grpo_advantage, O.rloo_advantage):
        whole, _ = fn(r, mask, uids)
        for ws in (2, 3, 8):
            for rows in sdist.shard_prompt_groups(uids, lens.tolist(), ws):
                part, _ = fn(r[rows], mask[rows], uids[rows])
                assert torch.equal(part, whole[rows])
        split = [list(range(0, B // 2)), list(range(B // 2, B))]
        assert any(not torch.equal(fn(r[rows], mask[rows], uids[rows])[0], whole[rows]) for rows in split)


def test_config_validation_matches_reference_messages():
    lp = torch.zeros(2, 3)
    with pytest.raises(AssertionError, match="loss_type must be either"):
        losses.ppo_policy_loss(lp, lp, lp, O.make_cfg(policy_loss_type="gspo"))
    with pytest.raises(AssertionError, match="loss_reduction must be either"):
        losses.ppo_policy_loss(lp, lp, lp, O.make_cfg(loss_reduction="bogus"))
    with pytest.raises(AssertionError, match="max_seq_len must be provided"):
        losses.ppo_policy_loss(lp, lp, lp, O.make_cfg(loss_reduction="seq_mean_token_sum_norm", max_seq_len=None))
    with pytest.raises(ValueError, match="Invalid KL estimator type"):
        advantages.compute_approx_kl(lp, lp, None, "k9")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_ops_fail_loudly_without_cuda():
    lp = torch.zeros(2, 4, requires_grad=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        losses.ppo_policy_loss(lp, lp.detach(), lp.detach(), O.make_cfg())
    with pytest.raises(RuntimeError, match="CUDA"):
        advantages.compute_grpo_outcome_advantage(torch.zeros(2, 4), torch.ones(2, 4), ["a", "a"])
    with pytest.raises(RuntimeError, match="CUDA"):
        skyrl_b200.fused_lm_head(torch.zeros(4, 8, dtype=torch.bfloat16), torch.zeros(16, 8, dtype=torch.bfloat16),
                                 torch.zeros(4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        skyrl_b200.logprobs_from_logits(torch.zeros(2, 3, 5), torch.zeros(2, 3, dtype=torch.long))


def test_product_code_never_imports_the_oracle():
    import os
    import re

    pkg = os.path.dirname(skyrl_b200.__file__)
    for d in (pkg, os.path.join(os.path.dirname(pkg), "scripts")):  # measurement scripts are not checkers either
        for name in os.listdir(d):
            if name.endswith(".py"):
                src = open(os.path.join(d, name)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{name} imports the oracle"


def test_wrapper_glue_selects_the_response_window_rows(monkeypatch):
    """Host logic of model_wrapper.py (reference model_wrapper.py:278-291, 353-370) with the CUDA head stubbed out:
    which hidden rows and which labels reach the head, and where the results land in [B, A]."""
    import torch

    from skyrl_b200 import model_wrapper as mw

    seen = {}

    def fake_head(h, w, labels, temperature=1.0, entropy_mask=None, compute_entropy=True, chunk=0):
        seen["rows"], seen["labels"], seen["emask"] = h[:, 0].clone(), labels.clone(), entropy_mask
        return h[:, 0] * 10.0, h[:, 0] * 100.0   # "log-prob" and "entropy" encode the row that produced them

    monkeypatch.setattr(mw, "fused_lm_head", fake_head)
    B, S, A = 3, 8, 3
    sequences = torch.arange(B * S).view(B, S) + 1
    am = torch.tensor([[1, 1, 1, 1, 1, 1, 1, 1], [0, 0, 1, 1, 1, 1, 1, 0], [1, 1, 1, 1, 1, 0, 0, 0]])
    # ---- padded layout: hidden[b, s, 0] = b * S + s identifies the row
    hidden = torch.arange(B * S, dtype=torch.float32).view(B, S, 1)
    lp, ent = mw.action_log_probs_from_hidden(hidden, None, sequences, A, attention_mask=am, compute_entropy=True)
    want_rows = torch.tensor([[4, 5, 6], [12, 13, 14], [20, 21, 22]], dtype=torch.float32)  # positions S-A-1 .. S-2
    assert torch.equal(lp, want_rows * 10.0) and torch.equal(ent, want_rows * 100.0)
    assert torch.equal(seen["labels"].view(B, A), sequences[:, S - A:S])                   # next-token labels
    assert torch.equal(seen["emask"].view(B, A), am[:, S - A - 1:S - 1])
    # ---- sample packing: packed row j is the j-th valid token; its label is the NEXT packed token
    valid_pos = torch.nonzero(am.reshape(-1)).squeeze(-1)
    packed_hidden = torch.arange(valid_pos.numel(), dtype=torch.float32).view(-1, 1)
    lp, ent = mw.action_log_probs_from_packed_hidden(packed_hidden, None, sequences, am, A, compute_entropy=True)
    packed_of = {int(p): j for j, p in enumerate(valid_pos)}
    packed_tokens = sequences.reshape(-1)[valid_pos]
    for b in range(B):
        for k, s in enumerate(range(S - A - 1, S - 1)):
            if am[b, s]:
                j = packed_of[b * S + s]
                assert lp[b, k].item() == j * 10.0 and ent[b, k].item() == j * 100.0
                assert j in seen["rows"].long().tolist()
                where = seen["rows"].long().tolist().index(j)
                assert seen["labels"][where] == packed_tokens[(j + 1) % len(packed_tokens)]
            else:
                assert lp[b, k].item() == 0.0 and ent[b, k].item() == 0.0   # pad_input leaves zeros
    assert seen["rows"].numel() == int(am[:, S - A - 1:S - 1].sum())      # nothing else reached the head


def test_kl_controllers_match_reference_update_rule():  # reference: utils/ppo_utils.py:45-70
    from skyrl_b200 import trainer_ops

    ctrl = trainer_ops.AdaptiveKLController(0.05, target=0.1, horizon=100)
    ctrl.update(current=0.2, n_steps=10)          # error clipped to +0.2 -> x (1 + 0.2 * 10 / 100)
    assert ctrl.value == pytest.approx(0.05 * 1.02)
    ctrl.update(current=0.095, n_steps=20)        # error -0.05 -> x (1 - 0.05 * 20 / 100)
    assert ctrl.value == pytest.approx(0.05 * 1.02 * 0.99)
    ctrl.update(current=0.0, n_steps=100)         # clipped to -0.2 -> x 0.8
    assert ctrl.value == pytest.approx(0.05 * 1.02 * 0.99 * 0.8)
    fixed = trainer_ops.FixedKLController(0.3)
    fixed.update(current=123.0, n_steps=7)
    assert fixed.value == 0.3


def test_pack_by_lengths_drops_right_padding_and_builds_cu_seqlens():
    import torch

    from skyrl_b200 import pack_by_lengths

    B, A, H = 4, 6, 3
    lens = torch.tensor([6, 0, 2, 5])
    x = torch.arange(B * A * H, dtype=torch.float32).view(B, A, H)
    y = torch.arange(B * A).view(B, A)
    cu, (xp, none, yp) = pack_by_lengths(lens, x, None, y)
    assert cu.dtype == torch.int32 and cu.tolist() == [0, 6, 6, 8, 13]
    assert none is None and xp.shape == (13, H) and yp.shape == (13,)
    for b in range(B):
        assert torch.equal(xp[cu[b]:cu[b + 1]], x[b, :lens[b]])
        assert torch.equal(yp[cu[b]:cu[b + 1]], y[b, :lens[b]])


def test_reduce_metrics_mean_max_min():  # reference: workers/worker_utils.py:8-22
    from skyrl_b200.dist import reduce_metrics

    out = reduce_metrics({"loss": [1.0, 2.0, 6.0], "ratio_max": [1.0, 5.0, 2.0], "ratio_min": [0.5, 0.1, 0.9]})
    assert out == {"loss": 3.0, "ratio_max": 5.0, "ratio_min": 0.1}
    with pytest.raises(AssertionError):
        reduce_metrics({"loss": []})
    with pytest.raises(AssertionError):
        reduce_metrics({"loss": [1.0, "x"]})


def test_install_into_the_live_reference_registries(monkeypatch):
    """install_into_skyrl_train() against the UNMODIFIED reference registries (build container only): every name
    the reference registers ends up bound to the skyrl_b200 callable through the reference's own
    unregister/register API, the model_wrapper helpers are rebound, and the reference's dispatcher
    (compute_advantages_and_returns, ppo_utils.py:1190-1214) then reaches our code - which refuses CPU tensors
    instead of falling back."""
    import sys
    import types

    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("/root/reference not present (GPU box)")
    pu, _ = ref_loader.load()
    saved = {reg: dict(reg._functions) for reg in (pu.AdvantageEstimatorRegistry, pu.PolicyLossRegistry)}
    fake_mw = types.ModuleType("skyrl_train.model_wrapper")
    fake_mw.logprobs_from_logits = fake_mw.chunked_entropy_from_logits = None
    monkeypatch.setitem(sys.modules, "skyrl_train.model_wrapper", fake_mw)
    try:
        done = skyrl_b200.install_into_skyrl_train()
        assert set(done) >= {"grpo", "rloo", "gae", "reinforce++", "regular", "dual_clip", "gspo", "sapo", "cispo",
                             "clip_cov", "kl_cov", "cross_entropy", "importance_sampling", "logprobs_from_logits"}
        for name in ("grpo", "rloo", "gae", "reinforce++"):
            assert pu.AdvantageEstimatorRegistry.get(name).__module__ == "skyrl_b200.advantages", name
        for name in pu.PolicyLossRegistry.list_available():
            assert pu.PolicyLossRegistry.get(name).__module__ == "skyrl_b200.losses", name
        assert fake_mw.logprobs_from_logits is skyrl_b200.logprobs_from_logits
        assert fake_mw.chunked_entropy_from_logits is skyrl_b200.chunked_entropy_from_logits
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError, match="CUDA"):
                pu.compute_advantages_and_returns(
                    token_level_rewards=torch.zeros(2, 3), response_mask=torch.ones(2, 3), index=["a", "a"],
                    adv_estimator="grpo", config=None)
    finally:
        for reg, fns in saved.items():
            reg._functions.clear()
            reg._functions.update(fns)


def test_fused_head_policy_is_the_hf_logits_path(monkeypatch):
    """FusedHeadPolicy (drop-in for HFModelWrapper.forward, model_wrapper.py:261-375) on a tiny random Qwen2 with the
    CUDA head replaced by the fp32 oracle head: action log-probs and the worker's entropy slice must equal the
    reference recipe on the model's own logits - logits / T, log_softmax + gather against roll(sequences, -1),
    entropy * attention_mask, slice [:, -A-1:-1] - and gradients must reach the trunk and the lm_head."""
    from transformers import Qwen2Config, Qwen2ForCausalLM

    from oracle import policy_oracle as O
    from skyrl_b200 import model_wrapper as mw

    def oracle_head(h, w, labels, temperature=1.0, entropy_mask=None, compute_entropy=True, chunk=0):
        lp, ent = O.head_logprobs_entropy(h.float(), w.float(), labels, temperature)
        return lp, (ent * entropy_mask.float() if entropy_mask is not None else ent)

    monkeypatch.setattr(mw, "fused_lm_head", oracle_head)
    torch.manual_seed(0)
    cfg = Qwen2Config(vocab_size=300, hidden_size=32, intermediate_size=64, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64,
                      tie_word_embeddings=False)
    model = Qwen2ForCausalLM(cfg).float()
    policy = skyrl_b200.FusedHeadPolicy(model)
    B, S, A, T = 3, 12, 5, 0.7
    sequences = torch.randint(0, 300, (B, S))
    am = torch.ones(B, S, dtype=torch.long)
    am[0, :3] = 0   # left padding of the prompt
    am[2, -2:] = 0  # right padding of the response
    lp, out = policy(sequences, A, am, temperature=T, return_output=True, compute_entropy=True)
    pos = am.cumsum(-1) - 1
    pos.masked_fill_(am == 0, 1)
    logits = model(sequences, attention_mask=am, position_ids=pos).logits / T
    want_lp = O.logprobs_from_logits(logits, torch.roll(sequences, -1, 1))[:, -A - 1:-1]
    want_ent = O.entropy_from_logits(logits, am)[:, -A - 1:-1]
    assert lp.shape == (B, A) and out["entropy"].shape == (B, S)
    assert torch.allclose(lp, want_lp, atol=1e-5)
    assert torch.allclose(out["entropy"][:, -A - 1:-1], want_ent, atol=1e-5)
    (lp.sum() + out["entropy"][:, -A - 1:-1].sum()).backward()
    g_fused = [model.lm_head.weight.grad.clone(), model.model.layers[0].mlp.up_proj.weight.grad.clone()]
    model.zero_grad()
    (want_lp.sum() + want_ent.sum()).backward()
    assert torch.allclose(g_fused[0], model.lm_head.weight.grad, atol=1e-4)
    assert torch.allclose(g_fused[1], model.model.layers[0].mlp.up_proj.weight.grad, atol=1e-4)
    assert isinstance(policy(sequences, [A], am), torch.Tensor)        # return_output=False, list form of num_actions
    _, out2 = policy(sequences, A, am, return_output=True, compute_entropy=True, entropy_requires_grad=False)
    assert not out2["entropy"].requires_grad


def test_value_head_critic_is_the_reference_slice(monkeypatch):
    """ValueHeadCritic (drop-in for CriticModel.forward, model_wrapper.py:425-500) on a tiny random Qwen2 trunk with the
    CUDA value head replaced by a matmul: values[:, :-1][:, -num_actions:] of Linear(hidden, 1) over all positions."""
    from transformers import Qwen2Config, Qwen2Model

    import importlib

    vh = importlib.import_module("skyrl_b200.value_head")  # the package re-exports the function under this name

    def matmul_head(hidden, weight, bias=None):
        v = hidden.float() @ weight.float().reshape(-1)
        return v + (bias.float().reshape(-1)[0] if bias is not None else 0.0)

    monkeypatch.setattr(vh, "value_head", matmul_head)
    torch.manual_seed(0)
    cfg = Qwen2Config(vocab_size=300, hidden_size=32, intermediate_size=64, num_hidden_layers=1,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=64)
    trunk = Qwen2Model(cfg).float()
    head = torch.nn.Linear(32, 1, bias=False)
    critic = skyrl_b200.ValueHeadCritic(trunk, head)
    B, S, A = 2, 11, 4
    ids = torch.randint(0, 300, (B, S))
    am = torch.ones(B, S, dtype=torch.long)
    am[0, :2] = 0
    values, outputs = critic(ids, A, am, return_output=True)
    want = head(outputs.last_hidden_state).squeeze(-1)[:, :-1][:, -A:]
    assert values.shape == (B, A) and torch.allclose(values, want, atol=1e-6)
    assert critic(ids, None, am, return_output=True).last_hidden_state.shape == (B, S, 32)


# ------------------------------------------------------------------------------------------ Ray mirror of the registries
class _FakeRay:
    """The slice of the Ray API the registries use (reference: ppo_utils.py:245-262, 335-361), in process: named
    actors created with ``ray.remote(cls).options(name=..., get_if_exists=True).remote()``, method calls through
    ``.remote()`` returning handles that ``ray.get`` resolves, ``get_actor`` raising ValueError for unknown names."""

    def __init__(self):
        self.actors = {}

    def is_initialized(self):
        return True

    def remote(self, cls):
        fake = self

        class _Remote:
            def options(self, name=None, get_if_exists=False, **_):
                self.name, self.get_if_exists = name, get_if_exists
                return self

            def remote(self, *a, **k):
                if self.name in fake.actors:
                    if not self.get_if_exists:
                        raise ValueError(f"actor {self.name} exists")
                    return fake.actors[self.name]
                handle = _Handle(cls(*a, **k))
                fake.actors[self.name] = handle
                return handle

        return _Remote()

    def get(self, ref):
        return ref

    def get_actor(self, name):
        if name not in self.actors:
            raise ValueError(f"Failed to look up actor with name '{name}'")
        return self.actors[name]

    def kill(self, handle):
        for k, v in list(self.actors.items()):
            if v is handle:
                del self.actors[k]


class _Handle:
    def __init__(self, obj):
        self._obj = obj

    def __getattr__(self, name):
        method = getattr(self._obj, name)

        class _M:
            @staticmethod
            def remote(*a, **k):
                return method(*a, **k)

        return _M()


def _module_level_estimator(token_level_rewards, response_mask, index, **kwargs):
    return token_level_rewards, token_level_rewards


def test_registries_mirror_through_a_named_actor_when_ray_is_live(monkeypatch):
    """With a Ray session the registries keep a named actor in sync (cloudpickled callables), so a function registered
    in the driver process is found by the registry of another process (reference: test_ppo_utils.py:405-557).  Ray is
    not installed here: the test drives the same code with an in-process fake of the Ray calls it makes."""
    from skyrl_b200 import registry as reg

    fake = _FakeRay()
    monkeypatch.setattr(reg, "_ray", lambda: fake)

    class Driver(reg.BaseFunctionRegistry):
        _actor_name = "test_registry_actor"
        _function_type = "Advantage estimator"

    class Worker(reg.BaseFunctionRegistry):       # the registry object of a different process: same actor name
        _actor_name = "test_registry_actor"
        _function_type = "Advantage estimator"

    Driver.register("mine", _module_level_estimator)
    assert "test_registry_actor" in fake.actors
    assert Worker.get("mine") is _module_level_estimator           # pickled by reference, found through the actor
    assert sorted(Worker.list_available()) == ["mine"]
    with pytest.raises(ValueError, match="already registered"):
        Driver.register("mine", _module_level_estimator)
    Worker.register("theirs", skyrl_b200.compute_grpo_outcome_advantage)   # library functions pickle by reference too
    assert Driver.get("theirs") is skyrl_b200.compute_grpo_outcome_advantage
    Driver.unregister("mine")
    assert "mine" not in fake.actors["test_registry_actor"]._obj.table
    with pytest.raises(ValueError, match="Unknown advantage estimator 'nope'"):
        Driver.get("nope")
    with pytest.raises(ValueError, match="not registered"):
        Driver.unregister("nope")
    # a new Ray session (the actor is gone): the next call re-creates it and re-publishes the local table
    fake.actors.clear()
    assert Driver.get("theirs") is skyrl_b200.compute_grpo_outcome_advantage
    assert "theirs" in fake.actors["test_registry_actor"]._obj.table
    Driver.reset()
    assert Driver._functions == {} and "test_registry_actor" not in fake.actors


def test_fused_head_policy_refuses_a_sharded_lm_head_weight():
    """FusedHeadPolicy reads lm_head.weight without going through lm_head.forward, so FSDP's un-shard hooks do not fire
    for it: a flat / sharded view must be refused with a clear error, a DTensor-like weight is gathered."""
    import torch

    from skyrl_b200.hf_wrapper import _whole_weight

    w = torch.zeros(10, 8)
    assert _whole_weight(w, 8) is w
    with pytest.raises(RuntimeError, match="FSDP1"):
        _whole_weight(torch.zeros(40), 8)          # FSDP1 flat shard
    with pytest.raises(RuntimeError, match="FSDP1"):
        _whole_weight(torch.zeros(0), 8)           # freed original parameter

    class FakeDTensor:
        def full_tensor(self):
            return w

    assert _whole_weight(FakeDTensor(), 8) is w
