"""Control segment (common/ctl_segment.py) and the meta plane built on it
(shm_handler._MetaPlane): SURVEY §8 f.4 — the meta tree of a shard in shared memory,
guarded by a seqlock, instead of two pickled SharedDict.set round trips per save
(reference ckpt_saver.py:315-327, multi_process.py:579-672)."""

import pickle
import threading
import time

import pytest
import torch

from dlrover_b200.common import ctl_segment as cs
from dlrover_b200.shm_handler import (DLROVER_CKPT_CONFIG_KEY, CheckpointConfig, CoopContext,
                                      SharedMemoryHandler)


def test_publish_snapshot_and_generations(run_env):
    owner = cs.ControlSegment.create(3)
    try:
        assert cs.ControlSegment.attach(4) is None
        peer = cs.ControlSegment.attach(3)
        assert peer is not None and peer.snapshot() is None and not peer.has_meta()
        tree = {"model": {"w": ("meta", 1, 2)}, "lr": 0.1}
        conf = pickle.dumps({"step": 5})  # any small blob
        assert peer.publish(step=5, writing=True, payload_bytes=99, conf_blob=conf,
                            meta_blob=pickle.dumps(tree))
        step, writing, payload, conf_blob, gen, meta = owner.snapshot()
        assert (step, writing, payload, gen) == (5, True, 99, 1) and meta == tree
        assert pickle.loads(conf_blob) == {"step": 5}
        # header-only update: same generation, the cached tree object is reused
        assert peer.publish(step=5, writing=False, payload_bytes=99, conf_blob=conf)
        step, writing, _, _, gen2, meta2 = owner.snapshot()
        assert (step, writing, gen2) == (5, False, 1) and meta2 is meta
        # a blob that does not fit is refused, nothing changes
        assert not peer.publish(step=6, writing=True, payload_bytes=1, conf_blob=conf,
                                meta_blob=b"x" * (peer.meta_capacity + 1))
        assert owner.snapshot()[0] == 5
        # a writer that died mid-update (odd seq) is taken over by the next one
        peer._put(cs._OFF_SEQ, peer._u64(cs._OFF_SEQ) + 1)
        assert owner.snapshot(timeout=0.05) is None
        assert peer.publish(step=7, writing=False, payload_bytes=1, conf_blob=conf,
                            meta_blob=pickle.dumps({"new": 1}))
        assert owner.snapshot()[0] == 7 and owner.snapshot()[5] == {"new": 1}
        owner.clear()
        assert peer.snapshot() is None
        peer.close()
    finally:
        owner.unlink()
        owner.close()


def test_readers_never_see_a_torn_header(run_env):
    owner = cs.ControlSegment.create(0)
    reader = cs.ControlSegment.attach(0)
    stop = threading.Event()
    bad = []

    def read_loop():
        while not stop.is_set():
            snap = reader.snapshot()
            if snap is None:
                continue
            step, writing, payload, conf_blob, gen, meta = snap
            if pickle.loads(conf_blob) != step or payload != step * 3 or meta != {"s": gen}:
                bad.append(snap)

    t = threading.Thread(target=read_loop)
    t.start()
    try:
        for step in range(1, 400):
            blob = pickle.dumps({"s": step // 2 + 1}) if step % 2 == 0 or step == 1 else None
            owner.publish(step=step, writing=bool(step & 1), payload_bytes=step * 3,
                          conf_blob=pickle.dumps(step), meta_blob=blob)
    finally:
        stop.set()
        t.join()
        reader.close()
        owner.unlink()
        owner.close()
    assert not bad


def test_cooperative_handshake_and_slots(run_env):
    ctl = cs.ControlSegment.create(0)
    try:
        base = ctl.coop_seq()
        ctxs = [CoopContext(ctl, i, 4, base) for i in range(4)]
        total = 10 * (2 << 20) + 12345
        wins = [c.window(total) for c in ctxs]
        assert wins[0][0] == 0 and wins[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(wins, wins[1:]))
        assert all(w[0] % (2 << 20) == 0 for w in wins)
        assert CoopContext(ctl, 0, 2, base).window(100) == (0, 100)  # tiny: all in the first slice
        assert CoopContext(ctl, 1, 2, base).window(100) == (100, 100)
        with pytest.raises(TimeoutError):
            ctl.wait_coop_open(base + 1, timeout=0.05)
        threading.Timer(0.05, ctl.next_coop_seq).start()
        assert ctl.wait_coop_open(base + 1, timeout=5) is True
        seq = base + 1
        assert not ctl.wait_slots(4, seq, timeout=0.05)
        for i in range(4):
            ctl.slot_set(i, seq)
        assert ctl.wait_slots(4, seq, timeout=1)
        ctl.slot_set(2, seq + 1, ok=False)
        for i in (0, 1, 3):
            ctl.slot_set(i, seq + 1)
        assert not ctl.wait_slots(4, seq + 1, timeout=1)     # a rank reported a failure
        assert ctl.next_coop_seq(aborted=True) == seq + 1
        assert ctl.wait_coop_open(seq + 1, timeout=1) is False
    finally:
        ctl.unlink()
        ctl.close()


def _save(trainer, sd, step):
    full = dict(sd)
    full[DLROVER_CKPT_CONFIG_KEY] = CheckpointConfig(step=step, paths={"m": f"/x/{step}"})
    trainer.save_state_dict(full)


def test_steady_state_saves_never_touch_the_shared_dict(run_env):
    agent = SharedMemoryHandler(0, host=True)      # creates the control segment
    trainer = SharedMemoryHandler(0, host=False)
    try:
        sd = {"w": torch.arange(1000, dtype=torch.float32), "opt": {"lr": 0.1, "step": 1}}
        ctl = agent.metadata.ctl
        assert ctl is not None and trainer.metadata.ctl is not None
        for step in (1, 2, 3):
            sd["w"].add_(1)
            _save(trainer, sd, step)
        assert trainer.metadata.dict_sets == 0 and trainer.metadata.ctl_publishes == 6
        assert ctl._u64(cs._OFF_META_GEN) == 1           # tree pickled once, header flipped 6x
        back = agent.load_state_dict()
        conf = back[DLROVER_CKPT_CONFIG_KEY]
        assert conf.step == 3 and conf.paths == {"m": "/x/3"} and conf.writing_shm is False
        assert torch.equal(back["w"], sd["w"]) and back["opt"] == {"lr": 0.1, "step": 1}
        del back
        # a non-tensor leaf changes (lr schedule): the tree goes out again
        sd["opt"]["lr"] = 0.05
        _save(trainer, sd, 4)
        assert ctl._u64(cs._OFF_META_GEN) == 2
        assert agent.load_state_dict()["opt"]["lr"] == 0.05
        # a shape changes: new layout, new tree, segment re-created, agent follows
        sd["w"] = torch.ones(500)
        _save(trainer, sd, 5)
        back = agent.load_state_dict()
        assert back["w"].shape == (500,) and back[DLROVER_CKPT_CONFIG_KEY].step == 5
        del back
        assert trainer.metadata.dict_sets == 0
        # the agent forgets the checkpoint ("node replaced"): both sides see nothing
        agent.metadata.set({})
        assert trainer.load_state_dict() == {} and agent.no_checkpoint_state()
    finally:
        trainer.close()
        agent.unlink()
        agent.close()


def test_without_a_control_segment_the_shared_dict_is_used(run_env, monkeypatch):
    monkeypatch.setenv("DLROVER_B200_CTL", "0")
    agent = SharedMemoryHandler(0, host=True)
    trainer = SharedMemoryHandler(0, host=False)
    try:
        assert agent.metadata.ctl is None and trainer.metadata.ctl is None
        sd = {"w": torch.arange(10, dtype=torch.float32)}
        _save(trainer, sd, 1)
        _save(trainer, sd, 2)
        assert trainer.metadata.dict_sets == 4 and trainer.metadata.ctl_publishes == 0
        assert agent.load_state_dict()[DLROVER_CKPT_CONFIG_KEY].step == 2
    finally:
        trainer.close()
        agent.unlink()
        agent.close()


def _coop_pair(agent, sd, step, fail_follower=False, monkeypatch=None):
    """Leader + follower handlers of one process saving the same (CPU) state dict."""
    import threading

    ctl = agent.metadata.ctl
    base = ctl.coop_seq()
    leader, follower = SharedMemoryHandler(0, host=False), SharedMemoryHandler(0, host=False)
    results = {}

    def run(name, handler, index):
        full = dict(sd)
        full[DLROVER_CKPT_CONFIG_KEY] = CheckpointConfig(step=step, paths={})
        coop = CoopContext(ctl, index, 2, base, timeout=5.0)
        try:
            if name == "follower" and fail_follower:
                # its slice cannot be written
                def boom(*a, **k):
                    raise OSError("disk on fire")
                handler.write_ranges = boom
            handler.save_state_dict(full, blocking=True, coop=coop)
            results[name] = "ok"
        except BaseException as e:  # noqa: BLE001
            results[name] = e

    threads = [threading.Thread(target=run, args=("leader", leader, 0)),
               threading.Thread(target=run, args=("follower", follower, 1))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(30)
    return leader, follower, results


def test_cooperative_save_two_handlers_one_image(run_env):
    agent = SharedMemoryHandler(0, host=True)
    try:
        sd = {"a": torch.arange(3 << 20, dtype=torch.float32), "n": 3,
              "b": torch.arange(1 << 20, dtype=torch.int64)}
        leader, follower, results = _coop_pair(agent, sd, step=4)
        assert results == {"leader": "ok", "follower": "ok"}
        back = agent.load_state_dict()
        assert back[DLROVER_CKPT_CONFIG_KEY].step == 4
        assert torch.equal(back["a"], sd["a"]) and torch.equal(back["b"], sd["b"]) and back["n"] == 3
        # the two windows really were written by different handlers
        total = leader.shared_memory.size
        (a0, a1), (b0, b1) = (CoopContext(None, i, 2, 0).window(total) for i in (0, 1))
        assert a0 == 0 and a1 == b0 and b1 == total and 0 < a1 < total
        del back
        leader.close()
        follower.close()
    finally:
        agent.unlink()
        agent.close()


def test_cooperative_save_fails_as_a_whole_when_one_rank_fails(run_env):
    """A follower that cannot write its slice reports it; the leader does NOT publish
    writing_shm=False — the torn image stays marked as being written."""
    agent = SharedMemoryHandler(0, host=True)
    try:
        sd = {"a": torch.arange(3 << 20, dtype=torch.float32)}
        leader, follower, results = _coop_pair(agent, sd, step=1)
        assert results == {"leader": "ok", "follower": "ok"}
        leader.close()
        follower.close()
        sd["a"].add_(1)
        leader, follower, results = _coop_pair(agent, sd, step=2, fail_follower=True)
        assert isinstance(results["follower"], OSError)
        assert isinstance(results["leader"], RuntimeError)      # "a local rank failed ..."
        conf = agent.metadata.get()[DLROVER_CKPT_CONFIG_KEY]
        assert conf.step == 2 and conf.writing_shm is True
        assert agent.load_state_dict() == {}                    # nobody trusts the segment
        leader.close()
        follower.close()
    finally:
        agent.unlink()
        agent.close()


def test_oversized_meta_falls_back_to_the_shared_dict(run_env, monkeypatch):
    """A tree that does not fit the control segment goes through the SharedDict (and the
    control segment is cleared so that readers do not see an older tree)."""
    agent = SharedMemoryHandler(0, host=True)
    trainer = SharedMemoryHandler(0, host=False)
    try:
        _save(trainer, {"w": torch.zeros(8)}, 1)
        assert trainer.metadata.ctl_publishes == 2 and trainer.metadata.dict_sets == 0
        big = {"w": torch.zeros(8), "blob": "x" * (agent.metadata.ctl.meta_capacity + 10)}
        _save(trainer, big, 2)
        assert trainer.metadata.dict_sets == 2
        back = agent.load_state_dict()
        assert back[DLROVER_CKPT_CONFIG_KEY].step == 2 and len(back["blob"]) == len(big["blob"])
        del back
        _save(trainer, {"w": torch.ones(8)}, 3)          # fits again: back on the segment
        assert agent.load_state_dict()[DLROVER_CKPT_CONFIG_KEY].step == 3
        assert float(agent.load_state_dict()["w"][0]) == 1.0
    finally:
        trainer.close()
        agent.unlink()
        agent.close()


def test_layout_fingerprint_notices_every_kind_of_change():
    from dlrover_b200.shm_handler import plan_layout

    def sd():
        return {"m": {"a": torch.zeros(4, 3), "b": [torch.zeros(2), {"c": torch.zeros(5)}]},
                "opt": {"lr": 0.1, "betas": (0.9, 0.95), "groups": [{"ids": [0, 1]}]},
                DLROVER_CKPT_CONFIG_KEY: CheckpointConfig(step=1, paths={"m": "/p/1"})}

    base = plan_layout(sd())
    same = sd()
    same[DLROVER_CKPT_CONFIG_KEY] = CheckpointConfig(step=2, paths={"m": "/p/2"})  # per save
    assert plan_layout(same, base).unchanged
    for mutate in (lambda d: d["m"].__setitem__("a", torch.zeros(4, 4)),        # shape
                   lambda d: d["m"].__setitem__("a", torch.zeros(4, 3).half()),  # dtype
                   lambda d: d["m"].__setitem__("z", d["m"].pop("a")),           # key renamed
                   lambda d: d["m"]["b"].append(torch.zeros(1)),                 # list grew
                   lambda d: d["opt"].__setitem__("lr", 0.05),                   # scalar leaf
                   lambda d: d["opt"]["groups"][0].__setitem__("ids", [0, 2]),   # nested leaf
                   lambda d: d["opt"].__setitem__("betas", (0.9, 0.99)),         # tuple leaf
                   lambda d: d.__setitem__("extra", 1)):                         # new key
        changed = sd()
        mutate(changed)
        assert not plan_layout(changed, base).unchanged, mutate
    # mutable leaves are captured by value at plan time
    live = sd()
    lay = plan_layout(live)
    live["opt"]["groups"][0]["ids"].append(99)
    assert lay.meta["opt"]["groups"][0]["ids"] == [0, 1]


def test_a_new_agent_starts_with_empty_meta(run_env):
    first = SharedMemoryHandler(0, host=True)
    trainer = SharedMemoryHandler(0, host=False)
    try:
        _save(trainer, {"w": torch.zeros(8)}, 5)
        assert first.metadata.get()[DLROVER_CKPT_CONFIG_KEY].step == 5
        first.close()                       # the agent process goes away, the segments stay
        second = SharedMemoryHandler(0, host=True)
        try:
            assert second.metadata.get() == {} and second.no_checkpoint_state()
            assert second.load_state_dict() == {}
        finally:
            second.unlink()
            second.close()
    finally:
        trainer.close()
