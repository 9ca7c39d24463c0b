"""Synthetic state_dict builders with the exact tensor names/shapes of the
configurations BASELINE.json names (no downloads: checkpointing is value
agnostic, but values are seeded and non-trivial so corruption shows)."""

from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch


def llama3_8b_shapes() -> List[Tuple[str, Tuple[int, ...]]]:
    """HF LlamaForCausalLM (Llama-3-8B): 291 tensors, 8,030,261,248 params."""
    h, inter, vocab, layers, kv = 4096, 14336, 128256, 32, 1024
    out: List[Tuple[str, Tuple[int, ...]]] = [("model.embed_tokens.weight", (vocab, h))]
    for l in range(layers):
        p = f"model.layers.{l}."
        out += [
            (p + "self_attn.q_proj.weight", (h, h)),
            (p + "self_attn.k_proj.weight", (kv, h)),
            (p + "self_attn.v_proj.weight", (kv, h)),
            (p + "self_attn.o_proj.weight", (h, h)),
            (p + "mlp.gate_proj.weight", (inter, h)),
            (p + "mlp.up_proj.weight", (inter, h)),
            (p + "mlp.down_proj.weight", (h, inter)),
            (p + "input_layernorm.weight", (h,)),
            (p + "post_attention_layernorm.weight", (h,)),
        ]
    out += [("model.norm.weight", (h,)), ("lm_head.weight", (vocab, h))]
    return out


def gpt2_small_shapes() -> List[Tuple[str, Tuple[int, ...]]]:
    """nanoGPT defaults (12 layers, 768, vocab 50304, block 1024, bias=True);
    the tied lm_head appears as its own state_dict entry."""
    d, L, vocab, block = 768, 12, 50304, 1024
    out = [("transformer.wte.weight", (vocab, d)), ("transformer.wpe.weight", (block, d))]
    for l in range(L):
        p = f"transformer.h.{l}."
        out += [
            (p + "ln_1.weight", (d,)), (p + "ln_1.bias", (d,)),
            (p + "attn.c_attn.weight", (3 * d, d)), (p + "attn.c_attn.bias", (3 * d,)),
            (p + "attn.c_proj.weight", (d, d)), (p + "attn.c_proj.bias", (d,)),
            (p + "ln_2.weight", (d,)), (p + "ln_2.bias", (d,)),
            (p + "mlp.c_fc.weight", (4 * d, d)), (p + "mlp.c_fc.bias", (4 * d,)),
            (p + "mlp.c_proj.weight", (d, 4 * d)), (p + "mlp.c_proj.bias", (d,)),
        ]
    out += [("transformer.ln_f.weight", (d,)), ("transformer.ln_f.bias", (d,)),
            ("lm_head.weight", (vocab, d))]
    return out


def scale_shapes(shapes, scale: float):
    """Shrink dim 0 of every >=2-D tensor by `scale` (for CPU-sized parity cases)."""
    if scale >= 1.0:
        return list(shapes)
    out = []
    for name, shp in shapes:
        if len(shp) >= 2:
            shp = (max(1, int(shp[0] * scale)),) + tuple(shp[1:])
        out.append((name, shp))
    return out


def shard_shapes(shapes, world: int, rank: int):
    """FSDP-style dim-0 row shard: rank r gets rows [r*ceil(n/w), ...)."""
    out = []
    for name, shp in shapes:
        n = shp[0]
        per = (n + world - 1) // world
        lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
        out.append((name, (hi - lo,) + tuple(shp[1:])))
    return out


def fill_(t: torch.Tensor, seed: int) -> torch.Tensor:
    """Cheap, seeded, non-trivial bit pattern (device-side, no big randn); large tensors
    are filled piece by piece so the temporaries stay small."""
    n = t.numel()
    if n == 0:
        return t
    flat = t.view(-1)
    piece = 32 << 20
    for lo in range(0, n, piece):
        hi = min(n, lo + piece)
        if t.dtype in (torch.float32, torch.bfloat16, torch.float16, torch.float64):
            idx = torch.arange(lo, hi, device=t.device, dtype=torch.float32)
            flat[lo:hi].copy_((torch.sin(idx * 0.001 + seed) * 0.02).to(t.dtype))
        else:
            idx = torch.arange(lo, hi, device=t.device, dtype=torch.int64)
            flat[lo:hi].copy_(((idx * 2654435761 + seed) % 251).to(t.dtype))
    return t


def build_state_dict(shapes, dtype=torch.bfloat16, device="cpu", seed=1234, fill=True):
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, (name, shp) in enumerate(shapes):
        t = torch.empty(shp, dtype=dtype, device=device)
        if fill:
            fill_(t, seed + i)
        sd[name] = t
    return sd


def adamw_state(params: Dict[str, torch.Tensor], seed=99, step_device=None):
    """optimizer.state_dict() of AdamW after one step: per param a 4-byte fp32
    `step` scalar + fp32 exp_avg + exp_avg_sq, then non-tensor param_groups.
    The 4-byte scalars make every following offset only 4-B aligned — the
    reference layout has no padding (ckpt_saver.py:293-301)."""
    state = {}
    for i, (name, p) in enumerate(params.items()):
        dev = step_device if step_device is not None else p.device
        state[i] = {
            "step": torch.tensor(1.0, dtype=torch.float32, device=dev),
            "exp_avg": fill_(torch.empty(p.shape, dtype=torch.float32, device=p.device), seed + 2 * i),
            "exp_avg_sq": fill_(torch.empty(p.shape, dtype=torch.float32, device=p.device), seed + 2 * i + 1),
        }
    groups = [{"lr": 3e-4, "betas": (0.9, 0.95), "eps": 1e-8, "weight_decay": 0.1,
               "amsgrad": False, "params": list(range(len(params)))}]
    return {"state": state, "param_groups": groups}


def payload_bytes(sd) -> int:
    total = 0
    stack = [sd]
    while stack:
        v = stack.pop()
        if isinstance(v, dict):
            stack.extend(v.values())
        elif isinstance(v, list):
            stack.extend(v)
        elif torch.is_tensor(v):
            total += v.numel() * v.element_size()
    return total


class ShardedStateFactory:
    """BASELINE configs[2] without the model code: the state an FSDP full-shard job hands
    to the checkpointer — every parameter row-sharded over the ranks (DTensor, Shard(0)),
    its AdamW moments sharded the same way (fp32), a replicated `step` per parameter and
    the non-tensor param_groups.  FSDP gives out FRESH tensors on every state_dict() call;
    `build(variant)` reproduces that: the local shards of variant 0 and 1 are views of
    the same flat buffers at different byte offsets, so every device pointer changes from
    one save to the next while the memory footprint stays that of one state."""

    SHIFT = 512  # bytes between the two variants' views

    def __init__(self, shapes, world: int, rank: int, device, mesh=None, with_optimizer=True,
                 weight_dtype=torch.bfloat16, seed=4321):
        self.world, self.rank, self.device, self.mesh = world, rank, device, mesh
        self.with_optimizer = with_optimizer
        self.weight_dtype = weight_dtype
        self.full_shapes = list(shapes)
        self.local_shapes = shard_shapes(self.full_shapes, world, rank)
        w_elems = sum(int(torch.Size(s).numel()) for _, s in self.local_shapes)
        w_es = torch.empty(0, dtype=weight_dtype).element_size()
        self._wbuf = torch.empty(w_elems * w_es + self.SHIFT, dtype=torch.uint8, device=device)
        fill_(self._wbuf, seed + rank)
        self._mbuf = None
        if with_optimizer:
            self._mbuf = torch.empty(2 * w_elems * 4 + self.SHIFT, dtype=torch.uint8, device=device)
            fill_(self._mbuf, seed + 100 + rank)
        self.local_bytes = w_elems * w_es + (2 * w_elems * 4 if with_optimizer else 0)

    def _wrap(self, local, full_shape):
        if self.mesh is None:
            return local
        from torch.distributed.tensor import DTensor, Shard

        stride = [1] * len(full_shape)
        for d in range(len(full_shape) - 2, -1, -1):
            stride[d] = stride[d + 1] * full_shape[d + 1]
        return DTensor.from_local(local, self.mesh, [Shard(0)], run_check=False,
                                  shape=torch.Size(full_shape), stride=tuple(stride))

    def build(self, variant: int):
        """{"model": {...}, "optim": {"state": {...}, "param_groups": [...]}} of this rank."""
        shift = self.SHIFT * (variant & 1)
        w_es = torch.empty(0, dtype=self.weight_dtype).element_size()
        model, opt_state = OrderedDict(), OrderedDict()
        wo, mo = shift, shift
        for (name, full), (_, loc) in zip(self.full_shapes, self.local_shapes):
            n = int(torch.Size(loc).numel())
            w = self._wbuf[wo:wo + n * w_es].view(self.weight_dtype).view(loc)
            wo += n * w_es
            model[name] = self._wrap(w, full)
            if self.with_optimizer:
                m1 = self._mbuf[mo:mo + n * 4].view(torch.float32).view(loc)
                m2 = self._mbuf[mo + n * 4:mo + 2 * n * 4].view(torch.float32).view(loc)
                mo += 2 * n * 4
                opt_state[name] = {"step": torch.tensor(float(1 + variant), device=self.device),
                                   "exp_avg": self._wrap(m1, full),
                                   "exp_avg_sq": self._wrap(m2, full)}
        sd = {"model": model}
        if self.with_optimizer:
            sd["optim"] = {"state": opt_state,
                           "param_groups": [{"lr": 3e-4, "betas": (0.9, 0.95), "eps": 1e-8,
                                             "weight_decay": 0.1,
                                             "params": [n for n, _ in self.full_shapes]}]}
        return sd

    @staticmethod
    def local_tensors(sd):
        """fqn -> the local (plain) tensor behind every sharded leaf of build()'s result."""
        out = OrderedDict()

        def local(t):
            return t.to_local() if hasattr(t, "to_local") else t

        for k, v in sd["model"].items():
            out[f"model.{k}"] = local(v)
        for k, st in sd.get("optim", {}).get("state", {}).items():
            out[f"optim.state.{k}.exp_avg"] = local(st["exp_avg"])
            out[f"optim.state.{k}.exp_avg_sq"] = local(st["exp_avg_sq"])
        return out
