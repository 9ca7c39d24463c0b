import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    info = json.load(open(os.path.join(GOLDEN, f"{name}.json")))
    image = np.fromfile(os.path.join(GOLDEN, f"{name}.bin"), dtype=np.uint8)
    return info, image


def meta_to_json(m, tensor_meta_types, config_types=()):
    """Same schema as tests/golden/make_golden.py:meta_to_json."""
    if isinstance(m, tensor_meta_types):
        return {"__tensor__": True, "shape": list(m.shape), "dtype": str(m.dtype),
                "element_size": m.element_size, "numel": m.numel, "offset": m.offset}
    if config_types and isinstance(m, config_types):
        return {"__config__": True, "step": m.step, "rank": m.rank, "paths": m.paths}
    if isinstance(m, dict):
        return {"__dict__": [[repr(k) if not isinstance(k, str) else k,
                              meta_to_json(v, tensor_meta_types, config_types)]
                             for k, v in m.items()]}
    if isinstance(m, list):
        return {"__list__": [meta_to_json(v, tensor_meta_types, config_types) for v in m]}
    if isinstance(m, tuple):
        return {"__tuple__": list(m)}
    return {"__leaf__": repr(m)}


def strip_config(meta_json):
    """Drop the _DLORVER_CKPT_CONFIG entry (paths are temp dirs)."""
    return {"__dict__": [kv for kv in meta_json["__dict__"] if kv[0] != "_DLORVER_CKPT_CONFIG"]}


def tree_equal(a, b):
    if isinstance(a, dict):
        return isinstance(b, dict) and list(a.keys()) == list(b.keys()) and all(
            tree_equal(a[k], b[k]) for k in a)
    if isinstance(a, list):
        return isinstance(b, list) and len(a) == len(b) and all(
            tree_equal(x, y) for x, y in zip(a, b))
    if torch.is_tensor(a):
        return (torch.is_tensor(b) and a.dtype == b.dtype and tuple(a.shape) == tuple(b.shape)
                and torch.equal(a.cpu(), b.cpu()))
    return a == b


def to_device(tree, device):
    if isinstance(tree, dict):
        return type(tree)((k, to_device(v, device)) for k, v in tree.items())
    if isinstance(tree, list):
        return [to_device(v, device) for v in tree]
    if torch.is_tensor(tree):
        return tree.to(device)
    return tree
