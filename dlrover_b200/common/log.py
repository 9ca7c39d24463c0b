"""Logger for the checkpoint path (level from DLROVER_LOG_LEVEL, like
dlrover/python/common/log.py:90-127, without the rotating-file handlers)."""

import logging
import os
import sys

_LEVELS = {"DEBUG": logging.DEBUG, "INFO": logging.INFO, "WARNING": logging.WARNING,
           "ERROR": logging.ERROR, "CRITICAL": logging.CRITICAL}


def _build() -> logging.Logger:
    lg = logging.getLogger("dlrover_b200")
    if not lg.handlers:
        h = logging.StreamHandler(sys.stderr)
        h.setFormatter(logging.Formatter(
            "[%(asctime)s] [%(levelname)s] [%(filename)s:%(lineno)d] %(message)s"))
        lg.addHandler(h)
        lg.propagate = False
    lg.setLevel(_LEVELS.get(os.getenv("DLROVER_LOG_LEVEL", "INFO").upper(), logging.INFO))
    return lg


default_logger = _build()
