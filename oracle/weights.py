"""Reader for the AIRFEW01 weight containers (see tools/make_weights.py).  Test infrastructure."""
import os
import struct
import numpy as np

WEIGHT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights")
_DT = {0: np.float32, 1: np.float16, 2: np.int32}


def load_container(path, as_float32=True):
    with open(path, "rb") as fh:
        buf = fh.read()
    magic, count, _ = struct.unpack_from("<8sII", buf, 0)
    assert magic == b"AIRFEW01", magic
    out = {}
    for i in range(count):
        name, dt, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<120sII4IQQ", buf, 16 + 160 * i)
        name = name.rstrip(b"\0").decode()
        dims = [d0, d1, d2, d3][:nd]
        arr = np.frombuffer(buf, dtype=_DT[dt], count=int(np.prod(dims)), offset=off).reshape(dims)
        out[name] = arr.astype(np.float32) if (as_float32 and dt == 1) else arr.copy()
    return out


def load(model):
    """model in {'superpoint','plnet','lightglue','superglue_indoor','superglue_outdoor'}."""
    return load_container(os.path.join(WEIGHT_DIR, model + ".afw"))
