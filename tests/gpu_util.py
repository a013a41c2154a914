"""Helpers shared by the `-m gpu` parity tests (test infrastructure)."""
import importlib.util
import os
import sys

import numpy as np
import torch

from conftest import ROOT


def dev():
    return torch.device("cuda:0")


def t(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).to(dev())


def n(x):
    return x.detach().cpu().numpy()


def load_ref_ext(subdir, name):
    """Import a reference-built extension from oracle/_ref/<subdir>/ (None when it was not built)."""
    d = os.path.join(ROOT, "oracle", "_ref", subdir)
    if not os.path.isdir(d):
        return None
    for f in os.listdir(d):
        if f.startswith(name + ".") and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(d, f))
            mod = importlib.util.module_from_spec(spec)
            try:
                spec.loader.exec_module(mod)
            except ImportError:
                return None
            return mod
    return None


def dot(a, b):
    return float((a.double() * b.double()).sum())
