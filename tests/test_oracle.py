"""Pins the oracle (oracle/shm_layout.py + oracle/pack_oracle.c) to the golden
vectors produced by the reference itself (tests/golden/make_golden.py) and to
the known-answer values in the reference's tests."""

import ctypes
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import fixtures  # noqa: E402
from oracle import shm_layout as oracle  # noqa: E402
from tests.util import golden, meta_to_json, strip_config, tree_equal  # noqa: E402


@pytest.mark.parametrize("name", list(fixtures.FIXTURES))
def test_oracle_image_and_meta_match_reference(name):
    info, image = golden(name)
    sd = {"model_states": fixtures.FIXTURES[name]()}
    meta, buf = oracle.serialize(sd)
    assert buf.size == info["size"]
    assert hashlib.sha256(buf.tobytes()).hexdigest() == info["sha256"]
    assert np.array_equal(buf, image)
    got = meta_to_json(meta, oracle.OracleTensorMeta)
    assert got == strip_config(info["meta"])


def test_known_answer_sizes():
    # checkpoint_egine_test.py:251-252 ; checkpoint_backup_test.py:91,99
    for name, size in fixtures.KNOWN_SIZES.items():
        _, total = oracle.plan_layout({"model_states": fixtures.FIXTURES[name]()})
        assert total == size
    # test_ckpt_saver.py:117-124: 10x10 fp32 -> numel 100, element_size 4, offset 0
    meta, total = oracle.plan_layout({"x": torch.zeros(10, 10)})
    m = meta["x"]
    assert (m.numel, m.element_size, m.offset, m.shape) == (100, 4, 0, (10, 10)) and total == 400


@pytest.mark.parametrize("name", list(fixtures.FIXTURES))
def test_oracle_read_back(name):
    sd = {"model_states": fixtures.FIXTURES[name]()}
    meta, buf = oracle.serialize(sd)
    back = oracle.read_image(meta, buf)
    # non-contiguous inputs come back contiguous but equal
    assert tree_equal(back, sd)


def test_dcp_item_accounting():
    # fsdp_ckpt_test.py:188-209: a 2x4 fp32 item is 32 B, next offset 32
    assert oracle.dcp_item_offsets([32, 32, 5]) == [(0, 32), (32, 32), (64, 5)]


def _c_oracle():
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "oracle", "_ref",
                        "libpack_oracle.so")
    if not os.path.exists(path):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.dirname(os.path.dirname(path))])
    return ctypes.CDLL(path)


@pytest.mark.parametrize("threads", [1, 4])
def test_c_oracle_agrees_with_numpy_oracle(threads):
    lib = _c_oracle()
    sd = {"model_states": fixtures.fixture_mixed()}
    meta, want = oracle.serialize(sd)
    metas = oracle.flatten_tensor_metas(meta)
    leaves = []
    oracle.traverse(sd, lambda v: leaves.append(v) if torch.is_tensor(v) else None)
    srcs = [oracle.tensor_bytes(t).copy() for t in leaves]
    n = len(srcs)
    dst = np.zeros(want.size, dtype=np.uint8)
    ptrs = (ctypes.c_void_p * n)(*[s.ctypes.data if s.size else None for s in srcs])
    offs = (ctypes.c_uint64 * n)(*[m.offset for m in metas])
    lens = (ctypes.c_uint64 * n)(*[s.size for s in srcs])
    rc = lib.oracle_pack(ctypes.c_void_p(dst.ctypes.data), ctypes.c_uint64(n), ptrs, offs, lens,
                         ctypes.c_int(threads))
    assert rc == 0 and np.array_equal(dst, want)


@pytest.mark.parametrize("name", list(fixtures.FIXTURES))
def test_ref_port_matches_golden(name):
    """The bench's reference arm (oracle/ref_port.py) leaves the same bytes in
    its segment as the reference did."""
    from oracle.ref_port import RefPortSaver

    info, image = golden(name)
    saver = RefPortSaver(f"fc_refport_{os.getpid()}_{name}")
    try:
        saver.save({"model_states": fixtures.FIXTURES[name]()})
        got = np.frombuffer(saver.segment.buf, dtype=np.uint8)
        assert np.array_equal(got, image)
        del got
    finally:
        saver.close()


def test_product_never_touches_the_oracle_or_the_reference():
    """The oracle is the checker, never the thing shipped: nothing under dlrover_b200/
    imports it, names it, or reads /root/reference at run time (compat.py may NAME the
    reference's module paths — as strings to alias — but never opens the checkout)."""
    import ast
    import os

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dlrover_b200")
    offenders = []
    for dirpath, _, files in os.walk(root):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith((".cu", ".cuh", ".h")):
                if "oracle" in open(path).read():
                    offenders.append(path)
                continue
            if not f.endswith(".py"):
                continue
            src = open(path).read()
            if "/root/reference" in src:
                offenders.append(path + " (reads the reference checkout)")
            for node in ast.walk(ast.parse(src)):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                if any(n == "oracle" or n.startswith("oracle.") for n in names):
                    offenders.append(path)
    assert not offenders, offenders


_REAL_REF_CHILD = r"""
import json, os, sys
sys.path.insert(0, sys.argv[1])
sys.path.insert(0, os.path.join(sys.argv[1], "tests", "golden"))
os.environ["TORCHELASTIC_RUN_ID"] = "fcrealref%d" % os.getpid()
import hashlib
import torch
import fixtures
from oracle import ref_real
from oracle.ref_port import RefPortSaver
out = {"file": ref_real.load().__file__}
for i, (name, build) in enumerate(fixtures.FIXTURES.items()):
    sd = {"model_states": build()}
    real = ref_real.RealRefSaver(160 + i)
    port = RefPortSaver("fc_realref_%d_%d" % (os.getpid(), i))
    try:
        real.save(sd); real.save(sd)            # second save: the reference's cached-meta branch
        port.save(sd)
        a, b = bytes(real.buf), bytes(port.segment.buf)
        back = real.views()["model_states"]
        first = next((t for t in torch.utils._pytree.tree_leaves(back) if torch.is_tensor(t)), None)
        out[name] = {"sha256": hashlib.sha256(a).hexdigest(), "bytes": len(a), "same_as_port": a == b,
                     "views": first is not None}
        del back, first
    finally:
        port.close(); real.close()
left = [f for f in os.listdir("/dev/shm") if os.environ["TORCHELASTIC_RUN_ID"] in f]
out["left_in_dev_shm"] = left
print("FCREAL " + json.dumps(out))
"""


def test_compiled_reference_matches_goldens_and_port(tmp_path):
    """oracle/_ref/pyref (the reference's own byte code, oracle/build_ref.py) is what the
    bench's reference arm times when it is there: it leaves the golden images in its segment,
    the same bytes as the port.  In a process of its own — other tests alias `dlrover.*`."""
    import subprocess
    import sys

    from oracle import build_ref, ref_real

    if not ref_real.available():
        if build_ref.build(verbose=False) == 0:
            pytest.skip("no /root/reference here and oracle/_ref/pyref not built")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", _REAL_REF_CHILD, root], capture_output=True, text=True,
                       timeout=300, cwd=str(tmp_path))
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("FCREAL ")]
    assert p.returncode == 0 and lines, p.stderr[-3000:]
    out = json.loads(lines[-1][len("FCREAL "):])
    assert os.sep + os.path.join("oracle", "_ref", "pyref") + os.sep in out["file"]
    assert out["left_in_dev_shm"] == []
    for name in fixtures.FIXTURES:
        info, image = golden(name)
        rec = out[name]
        assert rec["bytes"] == image.size and rec["same_as_port"], name
        assert rec["sha256"] == hashlib.sha256(image.tobytes()).hexdigest(), name
