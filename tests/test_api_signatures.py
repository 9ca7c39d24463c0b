"""Drop-in boundary: every public class / function / method of the
reference's Flash Checkpoint modules exists here under the same name and accepts the
same arguments in the same order (ours may add trailing, defaulted ones).  The
golden was extracted from the reference sources by tests/golden/make_api_signatures.py."""

import ast
import importlib
import inspect
import json
import os

import pytest

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "api_signatures.json")))

# Reference helpers that are implementation details of ITS design, not part of what a
# user / framework adapter calls; each with the reason it has no counterpart here.
NOT_CARRIED_OVER = {
    "engine.py:ReadyTensor":
        "singleton int32 flag tensor of the reference's readiness all-reduce; here the flag "
        "lives on the engine's private control stream (engine.check_all_rank_ready)",
    "engine.py:ReadyTensor.__init__": "see ReadyTensor",
    "fsdp_engine.py:SlicedBufferedReader":
        "file-window helper of the reference's FileReader; ours reads each item's "
        "(offset, length) window into a BytesIO and needs no stream wrapper",
    "fsdp_engine.py:SlicedBufferedReader.__init__": "see SlicedBufferedReader",
    "fsdp_engine.py:SlicedBufferedReader.seek": "see SlicedBufferedReader",
    "fsdp_engine.py:SlicedBufferedReader.tell": "see SlicedBufferedReader",
}

# Same argument, deliberately different default (value -> why).
DEFAULT_OVERRIDES = {
    ("multi_process.py:LocalSocketComm.__init__", "persist"):
        (None, "None = the class's own default (_persistent): lock, queue and dict clients keep "
               "one connection; the reference hard-codes False and overrides it in SharedLock"),
}


def _ours(module, qualname):
    obj = importlib.import_module(module)
    for part in qualname.split("."):
        obj = inspect.getattr_static(obj, part) if inspect.isclass(obj) else getattr(obj, part)
    if isinstance(obj, (staticmethod, classmethod)):
        obj = obj.__func__
    return obj


def _cases():
    for rel, entry in sorted(GOLDEN.items()):
        for name, params in sorted(entry["api"].items()):
            yield rel, entry["ours"], name, params


@pytest.mark.parametrize("rel,module,name,params", list(_cases()),
                         ids=[f"{os.path.basename(r)}:{n}" for r, _, n, _ in _cases()])
def test_signature_matches_reference(rel, module, name, params):
    key = f"{os.path.basename(rel)}:{name}"
    if key in NOT_CARRIED_OVER:
        pytest.skip(NOT_CARRIED_OVER[key])
    try:
        obj = _ours(module, name)
    except AttributeError:
        pytest.fail(f"{module} has no {name} (reference: {rel})")
    if params is None:
        assert inspect.isclass(obj), f"{name} should be a class"
        return
    if isinstance(obj, property):
        return
    sig = inspect.signature(obj)
    mine = list(sig.parameters.values())
    accepts_var_kw = any(p.kind is p.VAR_KEYWORD for p in mine)
    accepts_var_pos = any(p.kind is p.VAR_POSITIONAL for p in mine)
    named = [p for p in mine if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]
    want = [(n, d) for n, d in params if not n.startswith("*")]
    for i, (arg, default_src) in enumerate(want):
        if i >= len(named):
            assert accepts_var_kw or accepts_var_pos, f"{key}: missing argument {arg}"
            continue
        p = named[i]
        assert p.name == arg, f"{key}: argument {i} is {p.name!r}, reference has {arg!r}"
        if default_src is None:
            assert p.default is inspect.Parameter.empty, f"{key}: {arg} must stay required"
        else:
            assert p.default is not inspect.Parameter.empty, f"{key}: {arg} lost its default"
            try:
                literal = ast.literal_eval(default_src)
            except (ValueError, SyntaxError):
                continue  # symbolic default (a constant of the module): presence is enough
            if (key, arg) in DEFAULT_OVERRIDES:
                literal = DEFAULT_OVERRIDES[(key, arg)][0]
            assert p.default == literal, f"{key}: default of {arg} is {p.default!r}, not {literal!r}"
    for p in named[len(want):]:
        assert p.default is not inspect.Parameter.empty, \
            f"{key}: extra argument {p.name} must have a default"


def test_module_layout_mirrors_the_reference():
    """Every module of the reference's flash_checkpoint package has a same-named module
    here that exposes the same public classes and functions (so that switching is a
    change of the package prefix only)."""
    prefix = "dlrover/trainer/torch/flash_checkpoint/"
    for rel, entry in GOLDEN.items():
        if not rel.startswith(prefix):
            continue
        name = os.path.basename(rel)[:-3]
        mod = importlib.import_module(f"dlrover_b200.flash_checkpoint.{name}")
        for qual in entry["api"]:
            top = qual.split(".")[0]
            if f"{os.path.basename(rel)}:{top}" in NOT_CARRIED_OVER:
                continue
            assert hasattr(mod, top), f"dlrover_b200.flash_checkpoint.{name} lacks {top}"
