"""Reader/writer for the ``.mocflat`` named-array container.

Same layout as ``mocc_b200/host/arrayfile.hpp`` (the C++ side that writes the
flattened ray data).  ``.gz`` files are handled transparently so that golden
fixtures can be committed compressed.
"""
import gzip
import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"MOCFLAT1"
_DTYPES = {0: np.dtype("<f8"), 1: np.dtype("<i4"), 2: np.dtype("<i8"), 3: np.dtype("<u4")}
_CODES = {v: k for k, v in _DTYPES.items()}


def _open(path, mode):
    path = str(path)
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


def load_arrays(path):
    """Return an OrderedDict name -> numpy array (1-element arrays for scalars)."""
    with _open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a MOCFLAT1 container")
    (n,) = struct.unpack_from("<I", buf, 8)
    pos = 12
    out = OrderedDict()
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = buf[pos:pos + ln].decode()
        pos += ln
        dt, nd = struct.unpack_from("<II", buf, pos)
        pos += 8
        dims = struct.unpack_from(f"<{nd}Q", buf, pos)
        pos += 8 * nd
        pos += (-pos) % 8
        dtype = _DTYPES[dt]
        count = int(np.prod(dims, dtype=np.int64)) if nd else 1
        arr = np.frombuffer(buf, dtype=dtype, count=count, offset=pos).reshape(dims).copy()
        pos += count * dtype.itemsize
        pos += (-pos) % 8
        out[name] = arr
    return out


def save_arrays(path, arrays):
    chunks = [MAGIC, struct.pack("<I", len(arrays))]
    pos = 12

    def pad():
        nonlocal pos
        r = (-pos) % 8
        if r:
            chunks.append(b"\0" * r)
            pos += r

    for name, arr in arrays.items():
        arr = np.ascontiguousarray(arr)
        if arr.dtype not in _CODES:
            raise TypeError(f"{name}: unsupported dtype {arr.dtype}")
        nm = name.encode()
        hdr = struct.pack("<I", len(nm)) + nm + struct.pack("<II", _CODES[arr.dtype], arr.ndim)
        hdr += struct.pack(f"<{arr.ndim}Q", *arr.shape)
        chunks.append(hdr)
        pos += len(hdr)
        pad()
        raw = arr.tobytes()
        chunks.append(raw)
        pos += len(raw)
        pad()
    with _open(path, "wb") as f:
        f.write(b"".join(chunks))


def scalar(arrays, name):
    return arrays[name].reshape(-1)[0].item()
