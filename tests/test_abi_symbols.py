"""The C-ABI library loads on a CPU-only box, exports every symbol include/binest.h declares, and refuses
to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "binest.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(binest_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bayesianinference_b200 import _lib
    lib = _lib.load()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/binest.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) <= set(names), "ctypes binds symbols the header does not declare"
    assert lib.binest_version() == 100


def test_options_struct_layout_and_defaults():
    from bayesianinference_b200 import _lib
    o = _lib.Options()
    _lib.load().binest_default_options(ctypes.byref(o))
    # Options[nestedSampling] BS:837-851
    assert (o.pool_size, o.mc_steps, o.max_iter, o.min_iter, o.term_frac) == (100, 200, 10000, 100, 0.01)
    assert (o.acc_min, o.acc_max, o.batch_k, o.n_runs) == (0.0, 1.0, 1, 1)
    assert o.loglmax != o.loglmax  # "LogLikelihoodMaximum" -> Automatic travels as NaN (BS:847)
    assert ctypes.sizeof(_lib.Options) == 12 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bayesianinference_b200 import _lib, engine
    with pytest.raises(_lib.BinestError) as e:
        engine.init()
    assert e.value.code == 7  # BINEST_ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bayesianinference_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".wl")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "libbinest_oracle" not in src, f
