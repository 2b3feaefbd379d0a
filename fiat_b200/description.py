"""Store / load element descriptions (nested dicts of arrays, scalars and lists) as one .npz.

Descriptions are produced by `fiat_b200.extract.describe_element` from a reference-built FIAT
element.  Keeping them as files lets the GPU box (where FIAT is not installed) run parity tests
and the benchmark on exactly the constants the reference computed.
"""
import io
import json

import numpy

__all__ = ["save", "load", "dumps", "loads"]


def _split(node, path, arrays):
    if isinstance(node, dict):
        return {"__dict__": {k: _split(v, f"{path}/{k}", arrays) for k, v in node.items()}}
    if isinstance(node, numpy.ndarray):
        arrays[path] = node
        return {"__array__": path}
    if isinstance(node, (list, tuple)):
        return [_split(v, f"{path}/{i}", arrays) for i, v in enumerate(node)]
    if isinstance(node, (numpy.integer,)):
        return int(node)
    if isinstance(node, (numpy.floating,)):
        return float(node)
    if isinstance(node, (numpy.bool_,)):
        return bool(node)
    return node


def _join(node, arrays):
    if isinstance(node, dict):
        if "__array__" in node:
            return arrays[node["__array__"]]
        return {k: _join(v, arrays) for k, v in node["__dict__"].items()}
    if isinstance(node, list):
        return [_join(v, arrays) for v in node]
    return node


def dumps(desc):
    arrays = {}
    meta = _split(desc, "", arrays)
    buf = io.BytesIO()
    numpy.savez_compressed(buf, __meta__=numpy.frombuffer(json.dumps(meta).encode(), dtype=numpy.uint8),
                           **{k.replace("/", "|"): v for k, v in arrays.items()})
    return buf.getvalue()


def loads(data):
    with numpy.load(io.BytesIO(data), allow_pickle=False) as z:
        meta = json.loads(bytes(z["__meta__"]).decode())
        arrays = {k.replace("|", "/"): z[k] for k in z.files if k != "__meta__"}
    return _join(meta, arrays)


def save(path, desc):
    with open(path, "wb") as f:
        f.write(dumps(desc))


def load(path):
    with open(path, "rb") as f:
        return loads(f.read())
