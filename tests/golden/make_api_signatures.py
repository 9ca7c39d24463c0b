"""Dump the public call signatures of the reference's Flash Checkpoint path
 into api_signatures.json, by parsing the reference SOURCES
(no import: deepspeed / megatron are not installed).  Run here, where
/root/reference exists; the JSON is what tests/test_api_signatures.py checks our
classes against."""
import ast
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
FC = "dlrover/trainer/torch/flash_checkpoint/"
FILES = {
    FC + "checkpointer.py": "dlrover_b200.flash_checkpoint.api",
    FC + "ddp.py": "dlrover_b200.flash_checkpoint.api",
    FC + "fsdp.py": "dlrover_b200.flash_checkpoint.fsdp",
    FC + "deepspeed.py": "dlrover_b200.flash_checkpoint.deepspeed",
    FC + "megatron.py": "dlrover_b200.flash_checkpoint.megatron",
    FC + "megatron_dist_ckpt.py": "dlrover_b200.flash_checkpoint.megatron_dist_ckpt",
    FC + "engine.py": "dlrover_b200.flash_checkpoint.engine",
    FC + "full_ckpt_engine.py": "dlrover_b200.flash_checkpoint.engine",
    FC + "deepspeed_engine.py": "dlrover_b200.flash_checkpoint.engine",
    FC + "megatron_engine.py": "dlrover_b200.flash_checkpoint.engine",
    FC + "fsdp_engine.py": "dlrover_b200.flash_checkpoint.fsdp_engine",
    FC + "replica.py": "dlrover_b200.flash_checkpoint.replica",
    FC + "hf_trainer.py": "dlrover_b200.flash_checkpoint.hf_trainer",
    "dlrover/python/elastic_agent/torch/ckpt_saver.py": "dlrover_b200.ckpt_saver",
    "dlrover/python/common/storage.py": "dlrover_b200.common.storage",
    "dlrover/python/common/multi_process.py": "dlrover_b200.common.multi_process",
}


def params(fn: ast.FunctionDef):
    a = fn.args
    pos = a.posonlyargs + a.args
    defaults = [None] * (len(pos) - len(a.defaults)) + list(a.defaults)
    out = [[p.arg, None if d is None else ast.unparse(d)] for p, d in zip(pos, defaults)]
    if a.vararg:
        out.append(["*" + a.vararg.arg, None])
    for p, d in zip(a.kwonlyargs, a.kw_defaults):
        out.append([p.arg, None if d is None else ast.unparse(d)])
    if a.kwarg:
        out.append(["**" + a.kwarg.arg, None])
    return out


def public(name):
    return name == "__init__" or not name.startswith("_")


result = {}
for rel, ours in FILES.items():
    tree = ast.parse(open(os.path.join(REF, rel)).read())
    entry = result.setdefault(rel, {"ours": ours, "api": {}})
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and public(node.name):
            entry["api"][node.name] = params(node)
        elif isinstance(node, ast.ClassDef) and not node.name.startswith("_"):
            entry["api"][node.name] = None  # the class itself must exist
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and public(sub.name):
                    entry["api"][f"{node.name}.{sub.name}"] = params(sub)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_signatures.json")
json.dump(result, open(out, "w"), indent=1, sort_keys=True)
print(out, sum(len(v["api"]) for v in result.values()), "entries")
