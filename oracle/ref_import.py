"""Import the reference's own ``models.Generator`` from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so this
module is used solely by ``oracle/gen_golden.py`` and by CPU tests that skip when
the reference tree is absent.  Recipe = SURVEY.md Appendix B: models.py:6 imports
utils, utils.py:3-7 imports matplotlib (absent here) -> register empty stand-ins.
"""
from __future__ import annotations

import json
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("M2S_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models.py"))


class _AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def load_config():
    with open(os.path.join(REFERENCE_ROOT, "config_custom.json"), "r", encoding="utf-8") as f:
        return _AttrDict(json.load(f))


def import_reference_models():
    """Return the reference ``models`` module (never cached under the name ``models``)."""
    for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    saved_path = list(sys.path)
    saved = {n: sys.modules.pop(n, None) for n in ("models", "utils", "env")}
    try:
        sys.path.insert(0, REFERENCE_ROOT)
        import models as ref_models  # noqa: WPS433 (the reference module)
    finally:
        sys.path[:] = saved_path
        for n in ("models", "utils", "env"):
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]
    return ref_models


def reference_generator(seed: int = 1234, **overrides):
    """Reference Generator with its own default random init under ``seed``
    (= config_custom.json:9); ``overrides`` replace config entries (e.g. resblock="2")."""
    import warnings
    import torch
    ref_models = import_reference_models()
    h = load_config()
    h.update(overrides)
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = ref_models.Generator(h).eval()
    return g, h
