"""Name -> callable registries with the interface of skyrl_train's plugin API.

Mirrors (interface and error behaviour, not code) ``BaseFunctionRegistry``, ``AdvantageEstimatorRegistry``,
``PolicyLossRegistry`` and the two decorators of skyrl-train/skyrl_train/utils/ppo_utils.py:196-545:

* ``register(name, fn)``  -> ``ValueError("... already registered")`` on duplicates (ppo_utils.py:328-329)
* ``get(name)``           -> ``ValueError("Unknown ... Available: [...]")`` on a miss (ppo_utils.py:359-361)
* ``unregister(name)``    -> ``ValueError("... not registered")`` (ppo_utils.py:403-404)
* names may be ``StrEnum`` members or plain strings
* when a Ray session is live the table is mirrored through a named actor holding cloudpickled callables, so
  functions registered on the driver are visible inside the GPU workers (ppo_utils.py:245-262, 335-340).
  Every function this package registers is a module-level function that loads ``libskyb200.so`` lazily, so
  it pickles by reference.

Use :func:`install_into_skyrl_train` to put the B200 implementations into a real skyrl_train install.
"""
from __future__ import annotations

from enum import StrEnum
from functools import wraps
from typing import Callable, Dict, List, Optional, Union


def _ray():
    """The ray module if it is importable *and* a session is initialised, else None."""
    try:
        import ray  # type: ignore
    except Exception:  # noqa: BLE001 - ray is optional here
        return None
    try:
        return ray if ray.is_initialized() else None
    except Exception:  # noqa: BLE001
        return None


def _pickler():
    try:
        import cloudpickle

        return cloudpickle
    except ImportError:  # pragma: no cover
        import pickle

        return pickle


class _RegistryStore:
    """Plain-Python body of the named Ray actor that mirrors a registry across processes."""

    def __init__(self):
        self.table: Dict[str, bytes] = {}

    def register(self, name: str, blob: bytes) -> None:
        self.table[name] = blob

    def get(self, name: str) -> Optional[bytes]:
        return self.table.get(name)

    def list_available(self) -> List[str]:
        return list(self.table)

    def unregister(self, name: str):
        return self.table.pop(name, None)


def _key(name: Union[str, StrEnum]) -> str:
    return name.value if isinstance(name, StrEnum) else name


class BaseFunctionRegistry:
    _actor_name: Optional[str] = None
    _function_type: str = "Function"

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        cls._functions: Dict[str, Callable] = {}
        cls._ray_actor = None
        cls._synced_to_actor = False

    # ------------------------------------------------------------------ ray mirror
    @classmethod
    def _actor(cls):
        ray = _ray()
        if ray is None:
            raise Exception("Ray is not initialized, cannot create registry actor")
        if cls._ray_actor is None:
            remote_cls = ray.remote(_RegistryStore)
            cls._ray_actor = remote_cls.options(name=cls._actor_name, get_if_exists=True).remote()
        return cls._ray_actor

    @classmethod
    def sync_with_actor(cls) -> None:
        ray = _ray()
        if ray is None:
            raise Exception("Ray is not initialized, cannot sync with actor")
        try:
            ray.get_actor(cls._actor_name)
        except ValueError:  # stale handle from an earlier ray session
            cls._ray_actor = None
            cls._synced_to_actor = False
        actor = cls._actor()
        pk = _pickler()
        if not cls._synced_to_actor:
            for name, fn in cls._functions.items():
                ray.get(actor.register.remote(name, pk.dumps(fn)))
            cls._synced_to_actor = True
        for name in ray.get(actor.list_available.remote()):
            if name not in cls._functions:
                blob = ray.get(actor.get.remote(name))
                if blob is not None:
                    cls._functions[name] = pk.loads(blob)

    # ------------------------------------------------------------------ public API
    @classmethod
    def register(cls, name: Union[str, StrEnum], func: Callable) -> None:
        name = _key(name)
        if name in cls._functions:
            raise ValueError(f"{cls._function_type} '{name}' already registered")
        cls._functions[name] = func
        ray = _ray()
        if ray is not None:
            ray.get(cls._actor().register.remote(name, _pickler().dumps(func)))

    @classmethod
    def get(cls, name: Union[str, StrEnum]) -> Callable:
        name = _key(name)
        if _ray() is not None:
            cls.sync_with_actor()
        if name not in cls._functions:
            raise ValueError(
                f"Unknown {cls._function_type.lower()} '{name}'. Available: {list(cls._functions.keys())}"
            )
        return cls._functions[name]

    @classmethod
    def list_available(cls) -> List[str]:
        if _ray() is not None:
            cls.sync_with_actor()
        return list(cls._functions.keys())

    @classmethod
    def unregister(cls, name: Union[str, StrEnum]) -> None:
        name = _key(name)
        ray = _ray()
        if ray is not None:
            cls.sync_with_actor()
        found = name in cls._functions
        if found:
            del cls._functions[name]
        if ray is not None:
            actor = cls._actor()
            if name in ray.get(actor.list_available.remote()):
                found = True
                ray.get(actor.unregister.remote(name))
        if not found:
            raise ValueError(f"{cls._function_type} '{name}' not registered")

    @classmethod
    def reset(cls) -> None:
        ray = _ray()
        if ray is not None and cls._ray_actor is not None:
            try:
                ray.kill(ray.get_actor(cls._actor_name))
            except ValueError:
                pass
        cls._functions.clear()
        cls._ray_actor = None
        cls._synced_to_actor = False


class AdvantageEstimator(StrEnum):
    GAE = "gae"
    GRPO = "grpo"
    RLOO = "rloo"
    REINFORCE_PP = "reinforce++"


class PolicyLossType(StrEnum):
    REGULAR = "regular"
    DUAL_CLIP = "dual_clip"
    GSPO = "gspo"
    CISPO = "cispo"
    SAPO = "sapo"
    CLIP_COV = "clip_cov"
    KL_COV = "kl_cov"
    CROSS_ENTROPY = "cross_entropy"
    IMPORTANCE_SAMPLING = "importance_sampling"


class AdvantageEstimatorRegistry(BaseFunctionRegistry):
    """B200 estimators: ``grpo``, ``rloo``, ``gae``, ``reinforce++`` (the names ppo_utils.py:425-460 registers)."""

    _actor_name = "advantage_estimator_registry"
    _function_type = "advantage estimator"

    @classmethod
    def repopulate_registry(cls) -> None:
        from . import advantages as adv

        have = set(cls.list_available())
        for name, fn in (
            ("grpo", adv.compute_grpo_outcome_advantage),
            ("gae", adv.compute_gae_advantage_return),
            ("rloo", adv.compute_rloo_outcome_advantage),
            ("reinforce++", adv.compute_reinforce_plus_plus_outcome_advantage),
        ):
            if name not in have:
                cls.register(name, fn)


class PolicyLossRegistry(BaseFunctionRegistry):
    """B200 policy losses: ``regular``, ``dual_clip``, ``gspo``, ``sapo``, ``cispo``, ``clip_cov``, ``kl_cov``,
    ``cross_entropy``, ``importance_sampling`` — every name ppo_utils.py:463-508 registers."""

    _actor_name = "policy_loss_registry"
    _function_type = "policy loss"

    @classmethod
    def repopulate_registry(cls) -> None:
        from . import losses

        have = set(cls.list_available())
        for name, fn in (
            ("regular", losses.ppo_policy_loss),
            ("dual_clip", losses.ppo_policy_loss),
            ("gspo", losses.gspo_policy_loss),
            ("sapo", losses.sapo_policy_loss),
            ("cispo", losses.compute_policy_loss_cispo),
            ("clip_cov", losses.compute_policy_loss_clip_cov),
            ("kl_cov", losses.compute_policy_loss_kl_cov),
            ("cross_entropy", losses.cross_entropy_loss),
            ("importance_sampling", losses.importance_sampling_loss),
        ):
            if name not in have:
                cls.register(name, fn)


def register_advantage_estimator(name: Union[str, AdvantageEstimator]):
    """Decorator form of ``AdvantageEstimatorRegistry.register`` (ppo_utils.py:511-522)."""

    def deco(func: Callable):
        @wraps(func)
        def wrapper(*a, **k):
            return func(*a, **k)

        AdvantageEstimatorRegistry.register(name, wrapper)
        return wrapper

    return deco


def register_policy_loss(name: Union[str, PolicyLossType]):
    """Decorator form of ``PolicyLossRegistry.register`` (ppo_utils.py:525-536)."""

    def deco(func: Callable):
        @wraps(func)
        def wrapper(*a, **k):
            return func(*a, **k)

        PolicyLossRegistry.register(name, wrapper)
        return wrapper

    return deco


def repopulate_all_registries() -> None:
    PolicyLossRegistry.repopulate_registry()
    AdvantageEstimatorRegistry.repopulate_registry()


def sync_registries() -> None:
    if _ray() is None:
        raise ValueError("Ray is not initialized, cannot sync registries")
    PolicyLossRegistry.sync_with_actor()
    AdvantageEstimatorRegistry.sync_with_actor()


def install_into_skyrl_train(patch_model_wrapper: bool = True) -> List[str]:
    """Swap the B200 kernels into a live ``skyrl_train`` install (the drop-in step of INTEGRATION.md).

    Replaces every name skyrl_train registers - the ``grpo`` / ``rloo`` / ``gae`` / ``reinforce++`` estimators and the
    ``regular`` / ``dual_clip`` / ``gspo`` / ``sapo`` / ``cispo`` / ``clip_cov`` / ``kl_cov`` / ``cross_entropy`` /
    ``importance_sampling`` losses - in skyrl_train's own registries (through their ``unregister`` / ``register``, so a
    live Ray mirror follows) and points ``skyrl_train.model_wrapper.logprobs_from_logits`` /
    ``chunked_entropy_from_logits`` at the CUDA versions.  Returns the list of names it replaced.
    """
    import importlib

    pu = importlib.import_module("skyrl_train.utils.ppo_utils")
    from . import advantages as adv
    from . import losses
    from . import torch_utils as tu

    done = []
    for reg, table in (
        (pu.AdvantageEstimatorRegistry, {"grpo": adv.compute_grpo_outcome_advantage,
                                         "rloo": adv.compute_rloo_outcome_advantage,
                                         "gae": adv.compute_gae_advantage_return,
                                         "reinforce++": adv.compute_reinforce_plus_plus_outcome_advantage}),
        (pu.PolicyLossRegistry, {"regular": losses.ppo_policy_loss, "dual_clip": losses.ppo_policy_loss,
                                 "gspo": losses.gspo_policy_loss, "sapo": losses.sapo_policy_loss,
                                 "cispo": losses.compute_policy_loss_cispo,
                                 "clip_cov": losses.compute_policy_loss_clip_cov,
                                 "kl_cov": losses.compute_policy_loss_kl_cov,
                                 "cross_entropy": losses.cross_entropy_loss,
                                 "importance_sampling": losses.importance_sampling_loss}),
    ):
        for name, fn in table.items():
            if name in reg.list_available():
                reg.unregister(name)
            reg.register(name, fn)
            done.append(name)
    if patch_model_wrapper:
        mw = importlib.import_module("skyrl_train.model_wrapper")
        mw.logprobs_from_logits = tu.logprobs_from_logits
        mw.chunked_entropy_from_logits = tu.chunked_entropy_from_logits
        done += ["logprobs_from_logits", "chunked_entropy_from_logits"]
    return done
