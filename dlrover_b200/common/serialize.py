"""ClassMeta: how the trainer names a class for the agent to import.

Reference: dlrover/python/common/serialize.py:48-52.  The agent does
importlib.import_module(module_path).<class_name>(**kwargs)
(ckpt_saver.py:447-449, :490-493), so user-defined storages/savers plug in by
name.
"""

import importlib
from dataclasses import dataclass, field
from typing import Any, Dict


@dataclass
class ClassMeta:
    module_path: str = ""
    class_name: str = ""
    kwargs: Dict[str, Any] = field(default_factory=dict)

    def resolve(self):
        return getattr(importlib.import_module(self.module_path), self.class_name)

    def instantiate(self):
        return self.resolve()(**self.kwargs)
