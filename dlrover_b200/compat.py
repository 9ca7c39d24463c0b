"""Drop-in aliasing: make `import dlrover.trainer.torch.flash_checkpoint...` (and
the common/agent modules that path uses) resolve to this package, so training
scripts written against the reference run unchanged:

    import dlrover_b200.compat; dlrover_b200.compat.install()
    from dlrover.trainer.torch.flash_checkpoint.ddp import DdpCheckpointer, StorageType

Only the Flash Checkpoint module paths are provided (the rest of DLRover is out
of scope).  If the real `dlrover` distribution is importable, nothing is
touched unless force=True.
"""

from __future__ import annotations

import importlib
import importlib.util
import sys
import types

_FC = "dlrover.trainer.torch.flash_checkpoint"

# reference module path -> module of this package that carries the same names
ALIASES = {
    "dlrover.python.common.constants": "dlrover_b200.common.constants",
    "dlrover.python.common.env_utils": "dlrover_b200.common.env_utils",
    "dlrover.python.common.log": "dlrover_b200.common.log",
    "dlrover.python.common.multi_process": "dlrover_b200.common.multi_process",
    "dlrover.python.common.serialize": "dlrover_b200.common.serialize",
    "dlrover.python.common.singleton": "dlrover_b200.common.singleton",
    "dlrover.python.common.storage": "dlrover_b200.common.storage",
    "dlrover.python.elastic_agent.torch.ckpt_saver": "dlrover_b200.ckpt_saver",
    f"{_FC}.checkpointer": "dlrover_b200.flash_checkpoint.api",
    f"{_FC}.ddp": "dlrover_b200.flash_checkpoint.api",
    f"{_FC}.engine": "dlrover_b200.flash_checkpoint.engine",
    f"{_FC}.full_ckpt_engine": "dlrover_b200.flash_checkpoint.engine",
    f"{_FC}.deepspeed_engine": "dlrover_b200.flash_checkpoint.engine",
    f"{_FC}.megatron_engine": "dlrover_b200.flash_checkpoint.engine",
    f"{_FC}.fsdp_engine": "dlrover_b200.flash_checkpoint.fsdp_engine",
    f"{_FC}.fsdp": "dlrover_b200.flash_checkpoint.fsdp",
    f"{_FC}.deepspeed": "dlrover_b200.flash_checkpoint.deepspeed",
    f"{_FC}.megatron": "dlrover_b200.flash_checkpoint.megatron",
    f"{_FC}.megatron_dist_ckpt": "dlrover_b200.flash_checkpoint.megatron_dist_ckpt",
    f"{_FC}.replica": "dlrover_b200.flash_checkpoint.replica",
}


class _LazyAlias(types.ModuleType):
    """Module object that imports its target on first attribute access (the
    FSDP/DeepSpeed/Megatron modules pull in heavy or optional dependencies)."""

    def __init__(self, name, target):
        super().__init__(name)
        self.__dict__["_target_name"] = target
        self.__dict__["_target"] = None

    def _load(self):
        if self.__dict__["_target"] is None:
            self.__dict__["_target"] = importlib.import_module(self.__dict__["_target_name"])
        return self.__dict__["_target"]

    def __getattr__(self, item):
        if item.startswith("__") and self.__dict__["_target"] is None:
            # introspection (inspect.getmodule walks sys.modules) must not
            # trigger the import
            raise AttributeError(item)
        return getattr(self._load(), item)

    def __setattr__(self, key, value):  # tests monkeypatch e.g. megatron.get_args
        setattr(self._load(), key, value)

    def __dir__(self):
        return dir(self._load())


def _real_dlrover_present() -> bool:
    mod = sys.modules.get("dlrover")
    if mod is not None:
        return not getattr(mod, "__dlrover_b200_alias__", False)
    try:
        return importlib.util.find_spec("dlrover") is not None
    except (ImportError, ValueError):
        return False


def install(force: bool = False) -> bool:
    """Register the aliases.  Returns False (and does nothing) when the real
    dlrover is importable and force is not set."""
    if _real_dlrover_present() and not force:
        return False
    for alias, target in ALIASES.items():
        parts = alias.split(".")
        for i in range(1, len(parts)):
            pkg = ".".join(parts[:i])
            if pkg not in sys.modules or force and not getattr(
                    sys.modules[pkg], "__dlrover_b200_alias__", False):
                m = types.ModuleType(pkg)
                m.__path__ = []  # a package
                m.__dlrover_b200_alias__ = True
                sys.modules[pkg] = m
        mod = _LazyAlias(alias, target)
        mod.__dict__["__dlrover_b200_alias__"] = True
        sys.modules[alias] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    return True


# classes whose pickles cross the trainer <-> agent boundary, and the reference
# module each one is known under on the wire
_WIRE_CLASSES = {
    "dlrover.python.common.multi_process": (
        "dlrover_b200.common.multi_process",
        ["SocketRequest", "SocketResponse", "LockAcquireResponse", "LockedResponse",
         "QueueGetResponse", "QueueSizeResponse", "QueueEmptyResponse", "DictMessage"]),
    "dlrover.python.common.serialize": ("dlrover_b200.common.serialize", ["ClassMeta"]),
    "dlrover.python.common.storage": (
        "dlrover_b200.common.storage",
        ["PosixDiskStorage", "PosixStorageWithDeletion", "KeepStepIntervalStrategy",
         "KeepLatestStepStrategy"]),
    "dlrover.python.elastic_agent.torch.ckpt_saver": (
        "dlrover_b200.ckpt_saver",
        ["TensorMeta", "CheckpointConfig", "CheckpointEvent", "CheckpointEventType",
         "DdpCheckpointSaver", "MegatronCheckpointSaver", "DeepSpeedCheckpointSaver",
         "FsdpDcpSaver", "CommonDirCheckpointSaver", "TempDirCheckpointSaver"]),
    "dlrover.trainer.torch.flash_checkpoint.fsdp_engine": (
        "dlrover_b200.flash_checkpoint.fsdp_engine", ["_StorageInfo", "_StoragePrefix"]),
}


def enable_wire_compat() -> None:
    """Make this package's trainer half interoperable with the REFERENCE's agent
    half (a job launched with the reference's `dlrover-run`), and vice versa.

    Control messages, the meta tree and the saver/storage descriptors are
    pickles; pickle records `cls.__module__`.  After this call the listed
    classes claim the reference's module paths, and those paths resolve (in
    THIS process) to this package's modules, so what we send is what the
    reference would have sent, and what the reference sends unpickles into our
    classes.  Parent packages of an installed reference distribution are left
    in place — only the aliased modules are overridden.
    Also triggered by DLROVER_B200_WIRE_COMPAT=1 at `import dlrover_b200`.
    """
    for ref_mod, (our_mod, names) in _WIRE_CLASSES.items():
        ours = importlib.import_module(our_mod)
        parts = ref_mod.split(".")
        # parents: keep real ones when importable, else empty stub packages
        for i in range(1, len(parts)):
            pkg = ".".join(parts[:i])
            if pkg in sys.modules:
                continue
            try:
                importlib.import_module(pkg)
            except Exception:
                m = types.ModuleType(pkg)
                m.__path__ = []
                m.__dlrover_b200_alias__ = True
                sys.modules[pkg] = m
                if i > 1:
                    setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
        sys.modules[ref_mod] = ours
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], ours)
        for name in names:
            cls = getattr(ours, name)
            cls.__module__ = ref_mod
    # shm_handler defines TensorMeta/CheckpointConfig; ckpt_saver re-exports them


def uninstall():
    for name in [n for n, m in sys.modules.items()
                 if getattr(m, "__dlrover_b200_alias__", False) or
                 (isinstance(m, _LazyAlias))]:
        del sys.modules[name]
