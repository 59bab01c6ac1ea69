"""Load the UNMODIFIED reference hot-path modules: from /root/reference in the build container, else from the copy
of the same four files that ``__graft_entry__.build()`` vendors into ``baseline/_ref/`` (git-ignored, travels to the
GPU box with the repo snapshot - ``baseline/ref_path.py``).

Used by ``oracle/make_golden.py``, by the CPU tests that cross-check the oracle against the real reference, and by the
baseline arms of ``bench.py`` (``--impl reference``, ``cpu_baseline``, ``ref_gpu_path``).  Never by the product.
"""
from __future__ import annotations

import dataclasses
import importlib.util
import os
import sys
import types

_CANDIDATES = ["/root/reference/skyrl-train/skyrl_train",
               os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref",
                            "skyrl_train")]
REF_ROOT = os.environ.get("SKYRL_REF_ROOT") or next(
    (c for c in _CANDIDATES if os.path.isfile(os.path.join(c, "utils", "ppo_utils.py"))), _CANDIDATES[0])


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "ppo_utils.py"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _remote(*a, **k):  # supports @ray.remote and @ray.remote(...)
    if len(a) == 1 and callable(a[0]) and not k:
        return a[0]
    return lambda c: c


_cache = None


def load():
    """Returns (ppo_utils, torch_utils) modules of the reference. Stubs ray / omegaconf / skyrl_train.config."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise FileNotFoundError(f"reference not found under {REF_ROOT}")
    saved = {k: sys.modules.get(k) for k in ("ray", "omegaconf")}
    _stub("ray", remote=_remote, is_initialized=lambda: False)
    _stub("omegaconf", DictConfig=type("DictConfig", (dict,), {}), OmegaConf=object)
    _stub("skyrl_train").__path__ = [REF_ROOT]
    _stub("skyrl_train.utils").__path__ = [REF_ROOT + "/utils"]
    _stub("skyrl_train.config", AlgorithmConfig=dataclasses.make_dataclass("AlgorithmConfig", []))

    def ld(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    ld("skyrl_train.training_batch", "training_batch.py")
    tu = ld("skyrl_train.utils.torch_utils", "utils/torch_utils.py")
    ld("skyrl_train.utils.off_policy_correction_utils", "utils/off_policy_correction_utils.py")
    pu = ld("skyrl_train.utils.ppo_utils", "utils/ppo_utils.py")
    for k, v in saved.items():  # do not leave the stubs where the product code could see them
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    _cache = (pu, tu)
    return _cache
