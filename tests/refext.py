"""Loader for the UNMODIFIED reference extensions prebuilt into oracle/_ref (test infrastructure).
They exist only where oracle/build_ref.py has run (the build container) and travel to the GPU box
with gpurun; tests that want them skip when they are absent."""
import importlib.util
import os

import pytest

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
_cache = {}


def load(name):
    """name in raymarching | gridencoder | shencoder | freqencoder | ffmlp -> module or pytest.skip"""
    if name in _cache:
        return _cache[name]
    path = os.path.join(_REF, "_ref_%s.so" % name)
    if not os.path.exists(path):
        pytest.skip("reference extension %s not built (oracle/build_ref.py)" % name)
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod
