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
    import glob
    import shutil
    for d in glob.glob(f"/tmp/fc_bench_flags_{env['TORCHELASTIC_RUN_ID']}_*"):
        shutil.rmtree(d, ignore_errors=True)
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


REF_CHILD = textwrap.dedent("""
    import json, os, sys
    sys.path.insert(0, {root!r})
    os.environ["TORCHELASTIC_RUN_ID"] = "fcguardref%d" % os.getpid()
    import torch
    import bench
    if sys.argv[1] == "hidden":
        from oracle import ref_real
        ref_real.PYREF = "/nonexistent/pyref"
    sd = {{"model_states": {{"w": torch.arange(12, dtype=torch.float32).reshape(3, 4), "k": 5}}}}
    saver, kind, why = bench.reference_saver(170, sd)
    saver.save(sd)
    views = bench.reference_views(saver, kind)
    ok = torch.equal(views["model_states"]["w"], sd["model_states"]["w"])
    del views
    saver.close()
    print("FCG " + json.dumps({{"kind": kind, "why": why, "ok": ok, "how": bench.REFERENCE_HOW[kind]}}))
""")


@pytest.mark.parametrize("mode", ["built", "hidden"])
def test_reference_arm_uses_the_reference_itself_else_the_port(tmp_path, mode):
    from oracle import build_ref, ref_real

    if mode == "built" and not ref_real.available() and build_ref.build(verbose=False) == 0:
        pytest.skip("oracle/_ref/pyref not built and no /root/reference here")
    script = tmp_path / "child.py"
    script.write_text(REF_CHILD.format(root=ROOT))
    p = subprocess.run([sys.executable, str(script), mode], capture_output=True, text=True, timeout=300)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("FCG ")]
    assert p.returncode == 0 and lines, p.stderr[-2000:]
    rec = json.loads(lines[-1][4:])
    assert rec["ok"]
    if mode == "built":
        assert rec["kind"] == "reference" and rec["why"] is None
    else:
        assert rec["kind"] == "port" and "ImportError" in rec["why"]
