"""Deterministic fixture state_dicts shared by make_golden.py (which feeds them
to the REFERENCE implementation) and by the tests (which feed them to ours and
to the oracle).  Values come from integer arithmetic only — no RNG, no libm —
so the same bytes are produced on any box and torch version."""

from collections import OrderedDict

import torch


def pat(shape, dtype, seed):
    n = 1
    for s in shape:
        n *= s
    idx = torch.arange(n, dtype=torch.int64)
    v = (idx * 2654435761 + seed * 40503) % 251
    if dtype == torch.bool:
        t = (v % 2 == 0)
    elif dtype.is_floating_point:
        t = ((v - 125).to(torch.float32) / 64.0).to(dtype)  # exact in bf16/fp16
    else:
        t = v.to(dtype)
    return t.reshape(shape)


def simplenet_state():
    """SimpleNet = Linear(64,32)+Linear(32,10) fp32 (reference tests'
    fixture, dlrover/trainer/tests/torch/checkpoint_egine_test.py:79-92)."""
    return OrderedDict([
        ("fc1.weight", pat((32, 64), torch.float32, 1)),
        ("fc1.bias", pat((32,), torch.float32, 2)),
        ("fc2.weight", pat((10, 32), torch.float32, 3)),
        ("fc2.bias", pat((10,), torch.float32, 4)),
    ])


def sgd_empty_state():
    return {"state": {}, "param_groups": [{"lr": 0.01, "momentum": 0, "dampening": 0,
                                           "weight_decay": 0, "nesterov": False,
                                           "maximize": False, "foreach": None,
                                           "differentiable": False, "fused": None,
                                           "params": [0, 1, 2, 3]}]}


def fixture_simplenet_sgd():
    """-> 9640 bytes (8192+128+1280+40): checkpoint_egine_test.py:251-252."""
    return {"model": simplenet_state(), "optimizer": sgd_empty_state(), "step": 100}


def fixture_toymodel_adam():
    """ToyModel = Linear(16,16)+Linear(16,8) -> 1632 bytes
    (checkpoint_backup_test.py:91,99)."""
    model = OrderedDict([
        ("net1.weight", pat((16, 16), torch.float32, 5)),
        ("net1.bias", pat((16,), torch.float32, 6)),
        ("net2.weight", pat((8, 16), torch.float32, 7)),
        ("net2.bias", pat((8,), torch.float32, 8)),
    ])
    optim = {"state": {}, "param_groups": [{"lr": 0.001, "betas": (0.9, 0.999), "eps": 1e-8,
                                            "weight_decay": 0, "amsgrad": False,
                                            "params": [0, 1, 2, 3]}]}
    return {"model": model, "optimizer": optim}


def fixture_adamw_stepped():
    """AdamW state after a step: 4-byte fp32 `step` scalars make every later
    offset only 4-B aligned (no padding in the layout)."""
    model = simplenet_state()
    state = {}
    for i, (k, p) in enumerate(model.items()):
        state[i] = {"step": torch.tensor(float(i + 1)),
                    "exp_avg": pat(tuple(p.shape), torch.float32, 20 + i),
                    "exp_avg_sq": pat(tuple(p.shape), torch.float32, 30 + i)}
    optim = {"state": state, "param_groups": [{"lr": 3e-4, "betas": (0.9, 0.95), "eps": 1e-8,
                                               "weight_decay": 0.1, "params": [0, 1, 2, 3]}]}
    return {"model": model, "optimizer": optim, "epoch": 3}


def fixture_mixed():
    """Every edge the traversal/layout has: nested dict+list, tuple LEAF, all
    dtype widths incl. odd-length uint8/bool (1-byte aligned offsets), 0-numel,
    0-dim, non-contiguous (transposed) tensors, non-tensor leaves, None."""
    base = pat((6, 10), torch.float32, 40)
    return {
        "a": {
            "bf16": pat((7, 9), torch.bfloat16, 41),
            "u8_odd": pat((13,), torch.uint8, 42),
            "i64": torch.arange(11, dtype=torch.int64) * 3 - 7,
            "bool_odd": pat((5,), torch.bool, 43),
            "f16": pat((3, 3, 3), torch.float16, 44),
        },
        "list": [pat((4,), torch.int32, 45), {"x": pat((2, 2), torch.float64, 46)}, 17, "str",
                 [pat((1,), torch.int8, 47)]],
        "tuple_leaf": (1, 2, 3),
        "empty": torch.empty(0, dtype=torch.float32),
        "scalar": torch.tensor(2.5),
        "transposed": base.t(),            # shape (10,6), non-contiguous
        "strided": base[::2, 1::3],        # non-contiguous slice
        "i16": pat((9,), torch.int16, 48),
        "none": None,
        "flag": True,
    }


def fixture_llama_tiny():
    """Llama-3-8B key set with every dim shrunk (bf16), 2 layers."""
    h, inter, vocab, kv = 64, 224, 501, 16
    sd = OrderedDict()
    sd["model.embed_tokens.weight"] = pat((vocab, h), torch.bfloat16, 50)
    s = 51
    for l in range(2):
        p = f"model.layers.{l}."
        for name, shp in [("self_attn.q_proj.weight", (h, h)), ("self_attn.k_proj.weight", (kv, h)),
                          ("self_attn.v_proj.weight", (kv, h)), ("self_attn.o_proj.weight", (h, h)),
                          ("mlp.gate_proj.weight", (inter, h)), ("mlp.up_proj.weight", (inter, h)),
                          ("mlp.down_proj.weight", (h, inter)), ("input_layernorm.weight", (h,)),
                          ("post_attention_layernorm.weight", (h,))]:
            sd[p + name] = pat(shp, torch.bfloat16, s)
            s += 1
    sd["model.norm.weight"] = pat((h,), torch.bfloat16, s)
    sd["lm_head.weight"] = pat((vocab, h), torch.bfloat16, s + 1)
    return {"model": sd}


def fixture_dcp():
    """State dict for the DCP/FSDP engine golden: nested dicts of tensors (odd
    sizes, several dtypes) + non-tensor leaves (pickled by DCP as BYTE_IO)."""
    return {
        "model": {"w": pat((33, 17), torch.float32, 60), "b": pat((17,), torch.bfloat16, 61),
                  "emb": pat((50, 8), torch.float16, 62)},
        "optim": {"m": pat((33, 17), torch.float32, 63), "count": pat((3,), torch.int64, 64)},
        "step": 42,
        "name": "golden",
    }


FIXTURES = {
    "simplenet_sgd": fixture_simplenet_sgd,
    "toymodel_adam": fixture_toymodel_adam,
    "adamw_stepped": fixture_adamw_stepped,
    "mixed": fixture_mixed,
    "llama_tiny": fixture_llama_tiny,
}

KNOWN_SIZES = {"simplenet_sgd": 9640, "toymodel_adam": 1632}
