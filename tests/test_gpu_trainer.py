"""Trainer shell: a few epochs on synthetic molecules -- the loss goes down, checkpoints are written in the
reference's layout and load back into a fresh model, eval-mode validation leaves the BatchNorm buffers alone."""
import copy
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trainer_runs_checkpoints_and_resumes(tmp_path, monkeypatch):
    from molclr_b200 import GINet
    from molclr_b200.trainer import DEFAULT_CONFIG, MolCLR, SyntheticMoleculeDatasetWrapper
    monkeypatch.chdir(tmp_path)
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg.update(batch_size=64, epochs=3, warm_up=1, save_every_n_epochs=2, log_every_n_steps=4)
    cfg["model"].update(num_layer=3, emb_dim=64, feat_dim=64)
    cfg["dataset"].update(data_path="synthetic:700", valid_size=0.1)
    torch.manual_seed(0)
    trainer = MolCLR(SyntheticMoleculeDatasetWrapper(cfg["batch_size"], **cfg["dataset"]), cfg, log_root=str(tmp_path / "ckpt"))
    model, history = trainer.train()
    assert len(history) == 3 and all(h == h for h in history) and history[-1] < history[0]      # finite and decreasing
    ckpt = os.path.join(trainer.log_dir, "checkpoints")
    assert sorted(os.listdir(ckpt)) == ["model.pth", "model_1.pth"]
    sd = torch.load(os.path.join(ckpt, "model.pth"))
    fresh = GINet(**cfg["model"])
    fresh.load_state_dict(sd)                                   # strict: same keys / shapes as the reference layout
    assert int(sd["batch_norms.0.num_batches_tracked"]) > 0
    # validation runs in eval mode: running statistics and counters untouched
    before = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "tracked" in k}
    trainer._validate(model, trainer.dataset.get_data_loaders()[1])
    for k, v in model.state_dict().items():
        if k in before:
            assert torch.equal(v, before[k]), k


def test_trainer_on_the_device_resident_packed_store(tmp_path, monkeypatch):
    """Same loop fed by PackedMoleculeDatasetWrapper: batches are built and augmented on the GPU (no host batch, no H2D)."""
    from molclr_b200.trainer import DEFAULT_CONFIG, MolCLR, PackedMoleculeDatasetWrapper
    monkeypatch.chdir(tmp_path)
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg.update(batch_size=64, epochs=3, warm_up=1, save_every_n_epochs=5, log_every_n_steps=4)
    cfg["model"].update(num_layer=3, emb_dim=64, feat_dim=64)
    cfg["dataset"].update(data_path="synthetic:700", valid_size=0.1)
    torch.manual_seed(0)
    data = PackedMoleculeDatasetWrapper(cfg["batch_size"], **cfg["dataset"])
    train, valid = data.get_data_loaders()
    assert len(train) == 630 // 64 and len(valid) == 70 // 64
    xis, xjs = next(iter(train))
    assert xis.x.is_cuda and xis.num_graphs == 64 and xis.x.shape == xjs.x.shape and not torch.equal(xis.x, xjs.x)
    model, history = MolCLR(data, cfg, log_root=str(tmp_path / "ckpt")).train()
    assert len(history) == 3 and all(h == h for h in history) and history[-1] < history[0]


@pytest.mark.parametrize("aug", ["subgraph", "mix"])
def test_trainer_with_subgraph_and_mix_augmentation(tmp_path, monkeypatch, aug):
    """molclr.py:186-191 selects dataset_subgraph / dataset_mix by config['aug']: here the views come from the device kernels."""
    from molclr_b200.trainer import DEFAULT_CONFIG, MolCLR, build_dataset
    monkeypatch.chdir(tmp_path)
    cfg = copy.deepcopy(DEFAULT_CONFIG)
    cfg.update(batch_size=64, epochs=2, warm_up=1, save_every_n_epochs=5, log_every_n_steps=4, aug=aug)
    cfg["model"].update(num_layer=3, emb_dim=64, feat_dim=64)
    cfg["dataset"].update(data_path="synthetic:500", valid_size=0.1)
    torch.manual_seed(0)
    data = build_dataset(cfg)
    xis, xjs = next(iter(data.get_data_loaders()[0]))
    assert xis.x.is_cuda and int((xis.x[:, 0] == 118).sum()) > 0 and xis.num_graphs == 64
    model, history = MolCLR(data, cfg, log_root=str(tmp_path / "ckpt")).train()
    assert len(history) == 2 and all(h == h for h in history)
