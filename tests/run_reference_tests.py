"""Run the REFERENCE's own hot-path unit tests against THIS implementation.

The reference checkout (/root/reference) supplies the test files and their
helpers (local master, fixtures); the flash-checkpoint modules they import —
dlrover.python.common.{multi_process,storage,...},
dlrover.python.elastic_agent.torch.ckpt_saver,
dlrover.trainer.torch.flash_checkpoint.* — are replaced by this package's
modules before collection (dlrover_b200.compat.ALIASES, real parent packages
kept).  Usage (build container only):

    python tests/run_reference_tests.py [pytest args...]
"""

import importlib
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

DEFAULT_FILES = [
    "dlrover/python/tests/test_multi_process.py",
    "dlrover/python/tests/test_storage.py",
    "dlrover/python/tests/test_ckpt_saver.py",
    "dlrover/trainer/tests/torch/checkpoint_egine_test.py",
    "dlrover/trainer/tests/torch/ddp_checkpointer_test.py",
    "dlrover/trainer/tests/torch/megatron_ckpt_test.py",
    "dlrover/trainer/tests/torch/checkpoint_backup_test.py",
    "dlrover/trainer/tests/torch/fsdp_ckpt_test.py",
]


class _StubFinder:
    """`import kubernetes[.anything]` -> MagicMock modules (the package is not in
    this image; the reference's test helpers import it at module level)."""

    PREFIXES = ("kubernetes",)

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.PREFIXES:
            import importlib.machinery

            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock()
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        return m

    def exec_module(self, module):
        pass


def install_overrides():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    sys.meta_path.insert(0, _StubFinder())
    os.environ.setdefault("ROLE_NAME", "dlrover-trainer")
    os.environ.setdefault("DLROVER_B200_ASYNC_DRAIN", "0")  # the tests assume blocking saves
    from dlrover_b200 import compat

    # constants/log/env_utils stay the reference's (the test helpers need the
    # full modules); the path under test is swapped
    keep_real = {"dlrover.python.common.constants", "dlrover.python.common.log",
                 "dlrover.python.common.env_utils", "dlrover.python.common.singleton",
                 "dlrover.python.common.serialize"}
    swapped = []
    for alias, target in compat.ALIASES.items():
        if alias in keep_real:
            continue
        parent = alias.rsplit(".", 1)[0]
        importlib.import_module(parent)  # the real package
        mod = importlib.import_module(target)
        sys.modules[alias] = mod
        setattr(sys.modules[parent], alias.rsplit(".", 1)[1], mod)
        swapped.append(alias)
    return swapped


if __name__ == "__main__":
    import pytest

    # children started by the tests (mp.spawn) get the overrides through
    # tests/_ref_site/sitecustomize.py
    os.environ["DLROVER_B200_REF_TESTS"] = "1"
    site = os.path.join(ROOT, "tests", "_ref_site")
    os.environ["PYTHONPATH"] = os.pathsep.join(
        [site, ROOT, REF] + [p for p in os.environ.get("PYTHONPATH", "").split(os.pathsep) if p])
    sys._fc_ref_overrides = True
    swapped = install_overrides()
    print(f"[run_reference_tests] {len(swapped)} reference modules replaced by dlrover_b200")
    args = sys.argv[1:]
    files = [a for a in args if a.endswith(".py")]
    opts = [a for a in args if not a.endswith(".py")]
    if not files:
        files = DEFAULT_FILES
    os.chdir(REF)
    sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", "--rootdir", REF] + opts + files))
