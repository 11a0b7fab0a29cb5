"""Reader/writer for the DGD1 named-array container (see csrc/common/dgd_dump.h).

A DGD1 file carries stage outputs (panel model, minimizer index, anchor hits,
levelized expanded graph, DP results) between the C++ host glue, the reference
hook under oracle/, tests and bench.py.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_DTYPES = {0: np.uint8, 1: np.int32, 2: np.uint32, 3: np.int64, 4: np.uint64, 5: np.float64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def load(path: str) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"DGD1":
        raise ValueError(f"{path}: not a DGD1 file")
    pos = 4
    n = len(buf)
    while pos < n:
        (nl,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = buf[pos:pos + nl].decode()
        pos += nl
        dtype, count = struct.unpack_from("<IQ", buf, pos)
        pos += 12
        dt = np.dtype(_DTYPES[dtype])
        nbytes = dt.itemsize * count
        out[name] = np.frombuffer(buf, dtype=dt, count=count, offset=pos)
        pos += nbytes
    return out


def save(path: str, arrays: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(b"DGD1")
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            code = _CODES[a.dtype]
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<IQ", code, a.size))
            f.write(a.tobytes())


def ragged(d: Dict[str, np.ndarray], name: str):
    """Return (off, val) of a ragged list stored as <name>.off / <name>.val."""
    return d[name + ".off"], d[name + ".val"]
