"""Stand-alone mode (no dlrover-run agent): the saver daemon forked by local rank 0
must not leave the shared-memory segments behind when the trainer ends or is killed
(reference: fc/engine.py:118-137 forks the same daemon and leaks them)."""

import os
import subprocess
import sys
import textwrap
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import os, sys, torch
    sys.path.insert(0, {root!r})
    from dlrover_b200.flash_checkpoint.api import DdpCheckpointer, StorageType
    c = DdpCheckpointer(sys.argv[1])
    c.save_checkpoint(1, {{"w": torch.arange(1000.)}}, storage_type=StorageType.MEMORY)
    assert [f for f in os.listdir("/dev/shm") if os.environ["TORCHELASTIC_RUN_ID"] in f]
    print("saved", flush=True)
    if sys.argv[2] == "kill":
        os.kill(os.getpid(), 9)
""").format(root=ROOT)


@pytest.mark.parametrize("how", ["exit", "kill"])
def test_segments_do_not_outlive_a_standalone_trainer(tmp_path, how):
    run_id = f"standalone{os.getpid()}{how}"
    env = dict(os.environ, TORCHELASTIC_RUN_ID=run_id, DLROVER_LOG_LEVEL="WARNING")
    env.pop("DLROVER_B200_WIRE_COMPAT", None)
    out = subprocess.run([sys.executable, "-c", SCRIPT, str(tmp_path), how], env=env,
                         capture_output=True, text=True, timeout=120)
    assert "saved" in out.stdout, out.stderr[-2000:]
    deadline = time.time() + 20
    while time.time() < deadline:
        left = [f for f in os.listdir("/dev/shm") if run_id in f]
        if not left:
            break
        time.sleep(0.5)
    assert not left, left
