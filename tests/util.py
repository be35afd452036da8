"""Shared helpers for the test-suite."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def load_config():
    with open(os.path.join(ROOT, "config_custom.json"), "r", encoding="utf-8") as f:
        return AttrDict(json.load(f))


def multi_tap_reference(a, w, shifts, l_out, a_rows=None):
    """fp64 restatement of the engine contract: out[b,q,n] = sum_j sum_c A[b,q+shift_j,c] W[j,n,c]."""
    B, L, C = a.shape
    a = a.double().cpu()
    w = w.double().cpu()
    if a_rows is not None:
        a = a.clone()
        a[:, a_rows:] = 0
    lo = max(-min(shifts), 0)
    hi = max(max(shifts), 0) + max(l_out - L, 0)
    ap = torch.nn.functional.pad(a, (0, 0, lo, hi))
    out = torch.zeros(B, l_out, w.shape[1], dtype=torch.float64)
    for j, s in enumerate(shifts):
        out += ap[:, lo + s: lo + s + l_out, :] @ w[j].t()
    return out
