"""CPU tests of the C-ABI boundary: the library loads and exports every symbol include/zvx.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "zvx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zvx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(zvx):
    from zerovox_cpp_b200 import capi
    lib = capi.load_library()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"libzvx.so does not export {n}"
    assert sorted(capi.EXPORTS) == names


def test_default_config_matches_reference_constructor_arguments(zvx):
    """zerovox.cpp:117-138: dims 528/528/64/80, hop 300, k 7, scales 5,5,4,3, 3 blocks x dilations 1,3,5."""
    from zerovox_cpp_b200 import capi
    lib = capi.load_library()
    c = capi.Config()
    lib.zvx_default_config(ctypes.byref(c))
    assert (c.dim_in, c.style_dim, c.residual_dim, c.num_mels, c.hop_size, c.kernel_size) == (528, 528, 64, 80, 300, 7)
    assert list(c.upsample_scales)[:4] == [5, 5, 4, 3] and c.num_upsamples == 4
    assert list(c.resblock_dilations)[:9] == [1, 3, 5] * 3


def test_no_cpu_fallback(zvx, weights):
    """Without a CUDA device zvx_create must fail loudly -- there is no host path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from zerovox_cpp_b200 import capi
    with pytest.raises(capi.ZvxError, match="no CUDA device"):
        capi.Context({"hifigan.mean": weights["hifigan.mean"]})
