"""(Skipped in images without `accelerate`, this one included.)
FlashCkptTrainer on a tiny GPT-2 (CPU): checkpoints appear through the
agent, reload with from_pretrained, optimizer state is the Trainer's."""

import os
import time

import pytest
import torch

transformers = pytest.importorskip("transformers")
pytest.importorskip("accelerate")  # transformers.Trainer cannot be built without it

from dlrover_b200.ckpt_saver import AsyncCheckpointSaver  # noqa: E402


class _Data(torch.utils.data.Dataset):
    def __len__(self):
        return 16

    def __getitem__(self, i):
        ids = torch.arange(8) + i
        return {"input_ids": ids, "labels": ids}


def test_flash_ckpt_trainer_saves_through_agent(run_env, tmp_path):
    from transformers import GPT2Config, GPT2LMHeadModel, TrainingArguments

    from dlrover_b200.flash_checkpoint.hf_trainer import FlashCkptTrainer

    AsyncCheckpointSaver.start_async_saving_ckpt()
    cfg = GPT2Config(n_layer=1, n_head=2, n_embd=16, vocab_size=64, n_positions=16)
    model = GPT2LMHeadModel(cfg)
    args = TrainingArguments(output_dir=str(tmp_path), max_steps=4, save_steps=2,
                             per_device_train_batch_size=4, report_to=[], use_cpu=True,
                             save_total_limit=1, logging_steps=100, disable_tqdm=True)
    trainer = FlashCkptTrainer(model=model, args=args, train_dataset=_Data())
    before = torch.save
    trainer.train()
    assert torch.save is before
    trainer.wait_latest_checkpoint(timeout=120)
    deadline = time.time() + 60
    while trainer._get_last_checkpoint_step() != 4 and time.time() < deadline:
        time.sleep(0.5)
    assert trainer._get_last_checkpoint_step() == 4
    last = trainer.get_last_checkpoint()
    assert last == str(tmp_path / "checkpoint-4")
    files = set(os.listdir(last))
    assert {"model.safetensors", "optimizer.pt", "scheduler.pt", "rng_state.pth",
            "trainer_state.json", "config.json"} <= files
    reloaded = GPT2LMHeadModel.from_pretrained(last)
    for k, v in model.state_dict().items():
        assert torch.equal(reloaded.state_dict()[k], v.cpu()), k
    opt = torch.load(os.path.join(last, "optimizer.pt"), weights_only=False)
    assert set(opt) == {"state", "param_groups"}
    trainer.flash_checkpointer.async_save_engine.close()
