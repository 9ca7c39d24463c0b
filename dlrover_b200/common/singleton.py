"""Per-class lazily created singleton (dlrover/python/common/singleton.py:31-47)."""

import threading


class Singleton:
    _singleton_guard = threading.Lock()

    @classmethod
    def singleton_instance(cls, *args, **kwargs):
        inst = cls.__dict__.get("_singleton_obj")
        if inst is None:
            with Singleton._singleton_guard:
                inst = cls.__dict__.get("_singleton_obj")
                if inst is None:
                    inst = cls(*args, **kwargs)
                    cls._singleton_obj = inst
        return inst
