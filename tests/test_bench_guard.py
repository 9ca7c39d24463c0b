"""bench.py's deadline guard around the N>1 extra legs (cooperative, FSDP): whatever a leg
does — raise on one rank, hang on the others inside a collective — rank 0 prints the one
JSON line with the contract's numbers and every rank exits 0.  Ranks are plain processes
here; a hanging collective is a sleep."""

import json
import os
import subprocess
import sys
import textwrap
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, time
    sys.path.insert(0, {root!r})
    import bench
    rank, world, case = int(sys.argv[1]), 2, sys.argv[2]
    bench.AuxDeadline.GRACE = 1.0
    line = {{"metric": "checkpoint_GBps", "value": 54.0, "ddp_cooperative": None, "fsdp": None}}
    g = bench.AuxDeadline(rank, world, line)
    g.start()

    def coop():
        if case == "one_rank_raises" and rank == 1:
            raise RuntimeError("boom")
        if case == "one_rank_raises" and rank == 0:
            time.sleep(600)            # the collective the failed rank never joins
        if case == "all_raise":
            raise ValueError("off")
        if case == "hang":
            time.sleep(600)
        return {{"value": 300.0}}

    def fsdp():
        if case == "second_leg_fails_late" and rank == 1:
            raise RuntimeError("late")
        return {{"value": 269.0}}

    g.run("ddp_cooperative", coop)
    g.run("fsdp", fsdp)
    g.park_if_failed()
    if case == "second_leg_fails_late":
        time.sleep(600)                # rank 0 in the closing barrier, rank 1 parked
    g.finish()
""")


def _run(case, tmp_path, deadline="30"):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, BENCH_AUX_DEADLINE_S=deadline, TORCHELASTIC_RUN_ID=f"guard{os.getpid()}{case}",
               MASTER_PORT="0")
    t0 = time.time()
    procs = [subprocess.Popen([sys.executable, str(script), str(r), case], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=60) for p in procs]
    assert [p.returncode for p in procs] == [0, 0], outs
    lines = [ln for ln in outs[0][0].splitlines() if ln.startswith("{")]
    assert len(lines) == 1, outs[0]
    assert not [ln for ln in outs[1][0].splitlines() if ln.startswith("{")]
    return json.loads(lines[0]), time.time() - t0


def test_all_legs_fine(tmp_path):
    line, _ = _run("ok", tmp_path)
    assert line["value"] == 54.0
    assert line["ddp_cooperative"] == {"value": 300.0} and line["fsdp"] == {"value": 269.0}


def test_one_rank_raises_the_other_hangs(tmp_path):
    line, took = _run("one_rank_raises", tmp_path)
    assert line["value"] == 54.0 and took < 20
    assert "rank(s) 1 failed" in line["ddp_cooperative"]["error"]
    assert "error" in line["fsdp"]


def test_every_rank_raises(tmp_path):
    line, took = _run("all_raise", tmp_path)
    assert took < 20 and "ValueError: off" in line["ddp_cooperative"]["error"]
    assert "error" in line["fsdp"]


def test_hang_hits_the_deadline(tmp_path):
    line, took = _run("hang", tmp_path, deadline="2")
    assert took < 20 and "deadline" in line["ddp_cooperative"]["error"]
    assert line["value"] == 54.0


def test_late_failure_keeps_what_rank0_measured(tmp_path):
    line, took = _run("second_leg_fails_late", tmp_path)
    assert took < 20
    assert line["ddp_cooperative"] == {"value": 300.0}
    assert line["fsdp"] == {"value": 269.0}      # rank 0's own leg finished
