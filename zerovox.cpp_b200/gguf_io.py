"""Minimal GGUF v3 writer / reader for the zerovox weight layout.

The on-disk contract is the one `utils/zv2gguf.py` of the reference produces
(/root/reference/utils/zv2gguf.py:113-139 for the KV block, :156-180 for the
tensor naming / dtype rule) and `gguf_init_from_file` consumes
(/root/reference/ggml/src/ggml.c:6620): little-endian, magic "GGUF", version 3,
tensor data aligned to 32 bytes, dims stored fastest-first (ggml `ne` order, i.e.
the reverse of the numpy shape).

Only what the hot path needs is implemented: F32 / F16 tensors, uint32 and
string KVs.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

GGUF_MAGIC = 0x46554747  # "GGUF"
GGUF_VERSION = 3
ALIGNMENT = 32

GGML_TYPE_F32 = 0
GGML_TYPE_F16 = 1

_KV_UINT32 = 4
_KV_STRING = 8

_NP2GGML = {np.dtype(np.float32): GGML_TYPE_F32, np.dtype(np.float16): GGML_TYPE_F16}
_GGML2NP = {GGML_TYPE_F32: np.dtype(np.float32), GGML_TYPE_F16: np.dtype(np.float16)}


def _wstr(b: bytearray, s: str) -> None:
    raw = s.encode("utf-8")
    b += struct.pack("<Q", len(raw))
    b += raw


def write_gguf(path: str, kv: Dict[str, object], tensors: Dict[str, np.ndarray]) -> None:
    """Write `tensors` (numpy arrays, C order; shape is reversed into ggml ne) and `kv`
    (int -> uint32, str -> string) to `path`."""
    head = bytearray()
    head += struct.pack("<IIQQ", GGUF_MAGIC, GGUF_VERSION, len(tensors), len(kv))
    for key, val in kv.items():
        _wstr(head, key)
        if isinstance(val, str):
            head += struct.pack("<I", _KV_STRING)
            _wstr(head, val)
        else:
            head += struct.pack("<II", _KV_UINT32, int(val))
    offset = 0
    layout = []
    for name, arr in tensors.items():
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in _NP2GGML:
            raise ValueError(f"{name}: unsupported dtype {arr.dtype}")
        _wstr(head, name)
        ne = tuple(reversed(arr.shape))
        head += struct.pack("<I", len(ne))
        for d in ne:
            head += struct.pack("<Q", d)
        head += struct.pack("<IQ", _NP2GGML[arr.dtype], offset)
        layout.append((offset, arr))
        offset += (arr.nbytes + ALIGNMENT - 1) // ALIGNMENT * ALIGNMENT
    pad = (-len(head)) % ALIGNMENT
    with open(path, "wb") as f:
        f.write(head)
        f.write(b"\0" * pad)
        pos = 0
        for off, arr in layout:
            if off != pos:
                f.write(b"\0" * (off - pos))
                pos = off
            f.write(arr.tobytes())
            pos += arr.nbytes
        f.write(b"\0" * ((-pos) % ALIGNMENT))


def read_gguf(path: str) -> Tuple[Dict[str, object], Dict[str, np.ndarray]]:
    """Parse a GGUF v3 file written by `write_gguf` / gguf-py. Returns (kv, tensors);
    tensors are numpy arrays in numpy (reversed-ne) shape, memory-mapped read-only."""
    data = np.memmap(path, dtype=np.uint8, mode="r")
    buf = memoryview(data)
    pos = 0

    def rd(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos)
        pos += struct.calcsize(fmt)
        return v

    def rstr():
        nonlocal pos
        (n,) = rd("<Q")
        s = bytes(buf[pos:pos + n]).decode("utf-8")
        pos += n
        return s

    magic, version, n_tensors, n_kv = rd("<IIQQ")
    if magic != GGUF_MAGIC:
        raise ValueError("not a GGUF file")
    if version not in (2, 3):
        raise ValueError(f"unsupported GGUF version {version}")
    scalar = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<?", 10: "<Q", 11: "<q", 12: "<d"}
    kv: Dict[str, object] = {}
    for _ in range(n_kv):
        key = rstr()
        (vt,) = rd("<I")
        if vt == _KV_STRING:
            kv[key] = rstr()
        elif vt in scalar:
            (kv[key],) = rd(scalar[vt])
        elif vt == 9:  # array
            (et,) = rd("<I")
            (cnt,) = rd("<Q")
            if et == _KV_STRING:
                kv[key] = [rstr() for _ in range(cnt)]
            else:
                kv[key] = [rd(scalar[et])[0] for _ in range(cnt)]
        else:
            raise ValueError(f"unsupported kv type {vt}")
    infos = []
    for _ in range(n_tensors):
        name = rstr()
        (nd,) = rd("<I")
        ne = rd("<" + "Q" * nd)
        ttype, off = rd("<IQ")
        infos.append((name, ne, ttype, off))
    align = int(kv.get("general.alignment", ALIGNMENT))
    base = (pos + align - 1) // align * align
    tensors: Dict[str, np.ndarray] = {}
    for name, ne, ttype, off in infos:
        if ttype not in _GGML2NP:
            raise ValueError(f"{name}: unsupported ggml type {ttype}")
        dt = _GGML2NP[ttype]
        n = int(np.prod(ne))
        arr = np.frombuffer(buf, dtype=dt, count=n, offset=base + off).reshape(tuple(reversed(ne)))
        tensors[name] = arr
    return kv, tensors
