"""Loaded at interpreter start-up of the child processes the reference's tests
spawn (mp.spawn), so that they see the same module overrides as the parent
(tests/run_reference_tests.py).  Active only with DLROVER_B200_REF_TESTS=1."""
import os
import sys

if os.getenv("DLROVER_B200_REF_TESTS") == "1" and not getattr(sys, "_fc_ref_overrides", False):
    sys._fc_ref_overrides = True
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    try:
        import run_reference_tests

        run_reference_tests.install_overrides()
    except Exception as e:  # never break interpreter start-up
        sys.stderr.write(f"[ref-tests sitecustomize] overrides not installed: {e}\n")
