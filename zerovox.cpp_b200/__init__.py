"""zerovox.cpp_b200 -- B200-native (sm_100a) decoder + vocoder hot path of zerovox.cpp.

The directory name contains a dot, so it is imported through `zvxload.py` at the repo
root (`from zvxload import zvx`), which registers it as module `zerovox_cpp_b200`.
"""
from . import gguf_io, synth, sharding, convert  # noqa: F401
