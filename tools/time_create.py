import sys, time, os
sys.path.insert(0, os.getcwd())
from zvxload import zvx
from zerovox_cpp_b200 import capi
gguf = zvx.synth.write_model(zvx.synth.default_model_path())
t0 = time.perf_counter(); _, W = zvx.gguf_io.read_gguf(gguf); t1 = time.perf_counter()
c = capi.Context(W, device=0); t2 = time.perf_counter()
c.close()
c = capi.Context(W, device=0); t3 = time.perf_counter()
print(f"read_gguf {t1-t0:.3f}s  zvx_create (first, incl. CUDA init) {t2-t1:.3f}s  zvx_create (second) {t3-t2:.3f}s")
