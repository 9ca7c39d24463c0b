"""The C-ABI library loads without a GPU and exports every symbol that
include/flashckpt.h declares (no compute calls here)."""

import ctypes
import os
import re

import numpy as np
import pytest

from dlrover_b200 import _native as native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "flashckpt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(native.library_path())
    names = header_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in flashckpt.h but not exported"
    assert sorted(native.EXPORTED_SYMBOLS) == names


def test_version_and_strerror():
    lib = native.load_library()
    assert lib.fc_version() == 200
    assert lib.fc_strerror(0) == b"ok"
    assert b"flight" in lib.fc_strerror(native.FC_EBUSY)


def test_errors_are_codes_not_crashes():
    lib = native.load_library()
    assert lib.fc_arena_reserve(None, 10) == native.FC_EINVAL
    assert lib.fc_plan_destroy(None) == native.FC_OK
    assert lib.fc_host_pack(None, 0, None, None, None, 1) == native.FC_EINVAL
    assert b"fc_host_pack" in lib.fc_last_error()


def test_host_pack_needs_no_gpu():
    rng = np.random.default_rng(3)
    srcs = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in (0, 1, 4097, 9 << 20, 33)]
    offs, o = [], 7
    for s in srcs:
        offs.append(o)
        o += s.size
    for threads in (1, 5):
        dst = np.zeros(o, dtype=np.uint8)
        native.host_pack(dst.ctypes.data, [s.ctypes.data if s.size else 0 for s in srcs], offs,
                         [s.size for s in srcs], threads)
        for s, f in zip(srcs, offs):
            assert np.array_equal(dst[f:f + s.size], s)
        assert not dst[:7].any()


def test_missing_gpu_is_loud():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(native.NativeError):
        native.Context(0)


def _header_prototypes():
    text = open(os.path.join(ROOT, "include", "flashckpt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_][\w\s\*]*?)\b(fc_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;",
                                      text, flags=re.S):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (" ".join(ret.split()), params)
    return protos


def _c_kind(decl: str) -> str:
    if "*" in decl:
        return "ptr"
    for c, kind in (("uint64_t", "u64"), ("uint32_t", "u32"), ("int", "int"), ("float", "float")):
        if re.search(rf"\b{c}\b", decl):
            return kind
    return "?"


def _ctypes_kind(t) -> str:
    if t is None:
        return "void"
    if t is ctypes.c_void_p or t is ctypes.c_char_p or isinstance(t, type(ctypes.POINTER(ctypes.c_int))):
        return "ptr"
    return {ctypes.c_uint64: "u64", ctypes.c_uint32: "u32", ctypes.c_int: "int",
            ctypes.c_float: "float"}.get(t, "?")


def test_ctypes_signatures_agree_with_the_header():
    """Every prototype of include/flashckpt.h against the ctypes declaration that
    calls it: argument count and kind (pointer / u64 / u32 / int), return kind."""
    protos = _header_prototypes()
    assert sorted(protos) == sorted(native._SIGNATURES)
    for name, (ret, params) in protos.items():
        restype, argtypes = native._SIGNATURES[name]
        assert len(argtypes) == len(params), f"{name}: {len(argtypes)} ctypes args, header has {params}"
        for i, (decl, t) in enumerate(zip(params, argtypes)):
            assert _c_kind(decl) == _ctypes_kind(t), f"{name} arg {i}: header '{decl}' vs ctypes {t}"
        assert _c_kind(ret) == _ctypes_kind(restype), f"{name}: return '{ret}' vs ctypes {restype}"


def test_host_unpack_is_the_inverse_of_host_pack():
    rng = np.random.default_rng(5)
    sizes = (0, 1, 4097, 9 << 20, 33, 5 << 20)
    srcs = [rng.integers(0, 256, size=s, dtype=np.uint8) for s in sizes]
    offs, o = [], 3
    for s in srcs:
        offs.append(o)
        o += s.size + 1          # gaps between the ranges
    seg = np.zeros(o, dtype=np.uint8)
    native.host_pack(seg.ctypes.data, [s.ctypes.data if s.size else 0 for s in srcs], offs,
                     [s.size for s in srcs], 4)
    for threads in (1, 3, 8):
        outs = [np.zeros(s, dtype=np.uint8) for s in sizes]
        native.host_unpack(seg.ctypes.data, [t.ctypes.data if t.size else 0 for t in outs], offs,
                           list(sizes), threads)
        for a, b in zip(outs, srcs):
            assert np.array_equal(a, b)
    lib = native.load_library()
    assert lib.fc_host_unpack(None, 0, None, None, None, 1) == native.FC_EINVAL
    assert b"fc_host_unpack" in lib.fc_last_error()
