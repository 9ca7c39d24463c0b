"""A restarted trainer (bench leg `restore_fresh_process`): a NEW process that finds the
previous trainer's checkpoint in shared memory and restores it into live CUDA tensors.
  FC_FRESH_MODE=ours             DdpCheckpointer.load_checkpoint_into (staged bounce slots:
                                 no cudaHostRegister of the 16 GB segment before the copy)
  FC_FRESH_MODE=reference_style  load_checkpoint() (CPU views on the pageable segment) +
                                 per-tensor copy_ (ckpt_saver.py:144-161 + load_state_dict)
Prints one JSON line: seconds since process start at each stage."""
import json
import os
import sys
import time

T0 = time.time()
try:
    import psutil

    T0 = psutil.Process().create_time()
except Exception:
    pass

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ROLE_NAME"] = "dlrover-trainer"  # the agent (here: the parent's daemon) hosts the saver
os.environ.setdefault("DLROVER_LOG_LEVEL", "WARNING")
spec = json.loads(os.environ["FC_FRESH_SPEC"])
mode = os.getenv("FC_FRESH_MODE", "ours")

import torch  # noqa: E402

t_import = time.time()
from dlrover_b200 import shapes  # noqa: E402
from dlrover_b200.flash_checkpoint.api import DdpCheckpointer  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
sd = shapes.build_state_dict(shapes.scale_shapes(shapes.llama3_8b_shapes(), spec["scale"]),
                             torch.bfloat16, dev, fill=False)
for t in sd.values():
    t.zero_()
torch.cuda.synchronize()
t_model = time.time()
ckpt = DdpCheckpointer(spec["ckpt_dir"])
t_engine = time.time()
if mode == "ours":
    step = ckpt.load_checkpoint_into(sd)
    stats = dict(ckpt.engine._shm_handler.last_restore_stats)
else:
    loaded = ckpt.load_checkpoint()
    step = 1 if loaded else 0
    with torch.no_grad():
        for k, t in sd.items():
            t.copy_(loaded[k])
    stats = {}
    del loaded
torch.cuda.synchronize()
t_done = time.time()
ok = step > 0
for k, want in spec["probe"].items():
    got = int(sd[k].view(-1).view(torch.int16)[:4096].to(torch.int64).sum().item())
    ok = ok and got == want
S = shapes.payload_bytes(sd)
print(json.dumps({
    "mode": mode, "bit_exact_probe": bool(ok), "step": int(step),
    "restore_call_s": t_done - t_engine, "restore_GBps": S / (t_done - t_engine) / 1e9,
    "since_process_start_s": {"import_torch": t_import - T0, "cuda_init_and_alloc": t_model - T0,
                              "engine_ready": t_engine - T0, "restored": t_done - T0},
    "device_times": stats}), flush=True)
ckpt.engine.close()
os._exit(0)
