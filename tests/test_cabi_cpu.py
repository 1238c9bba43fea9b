"""CPU-only checks of the C ABI: the library builds, loads and exports every declared symbol,
and fails loudly (no CPU fallback) when no sm_100 device is present."""
import os

import pytest


def test_library_builds_and_exports_every_header_symbol():
    from mkids_sdr_b200 import _lib, build
    build.build()
    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    undeclared = [s for s in syms if s not in _lib._SIGNATURES]
    assert not undeclared, undeclared
    assert b'sm_100a' in lib.mkid_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from mkids_sdr_b200 import _lib
    with pytest.raises(_lib.MkidError) as e:
        _lib.Context(0)
    assert e.value.code == _lib.MKID_ENODEV


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, 'mkids_sdr_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f
