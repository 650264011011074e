"""Trainer shell for MolCLR pre-training on the molclr_b200 kernels.

Mirrors what ``molclr.py`` does around the hot path -- same ``config.yaml`` keys, same optimiser / schedule / checkpoint
behaviour -- as thin PyTorch host code:

* ``MolCLR(dataset, config).train()``: Adam(init_lr, weight_decay) (molclr.py:84-87), cosine annealing after ``warm_up``
  epochs (88-91,146-147), validation in eval mode every ``eval_every_n_epochs`` with best-model checkpointing (131-140),
  periodic ``model_{epoch}.pth`` (142-143), ``load_model`` resume from ``./ckpt/<name>/checkpoints/model.pth`` (149-158),
  TensorBoard scalars when tensorboard is importable (116-118,139).
* ``SyntheticMoleculeDatasetWrapper(batch_size, num_workers, valid_size, data_path)``: stands in for
  ``dataset.MoleculeDatasetWrapper`` (RDKit is not a dependency): deterministic synthetic molecules with the reference's
  node-mask / bond-delete augmentation (dataset.py:112-145), ``drop_last`` batches of (xis, xjs).

    python -m molclr_b200.trainer [config.yaml] [--epochs N] [--steps-per-epoch K]
"""
import math
import os
import shutil
import sys
import time

import torch

from . import GCN, GINet, NTXentLoss, pretrain_loss
from .graph import poll_checks
from .synth import make_pair_batch

DEFAULT_CONFIG = {
    "batch_size": 512, "warm_up": 10, "epochs": 100, "load_model": "None", "eval_every_n_epochs": 1, "save_every_n_epochs": 5,
    "log_every_n_steps": 50, "fp16_precision": False, "init_lr": 0.0005, "weight_decay": "1e-5", "gpu": "cuda:0", "model_type": "gin",
    "model": {"num_layer": 5, "emb_dim": 300, "feat_dim": 512, "drop_ratio": 0, "pool": "mean"},
    "aug": "node", "dataset": {"num_workers": 12, "valid_size": 0.05, "data_path": "synthetic:20000"},
    "loss": {"temperature": 0.1, "use_cosine_similarity": True},
}


class _Loader:
    """Iterable of (xis, xjs) pinned host batches; batch k of epoch e is a pure function of (seed, e, k)."""

    def __init__(self, batch_size, num_batches, seed, reshuffle):
        self.batch_size, self.num_batches, self.seed, self.reshuffle, self.epoch = batch_size, num_batches, seed, reshuffle, 0

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        base = self.seed + (self.epoch * 1_000_003 if self.reshuffle else 0)
        self.epoch += 1
        for k in range(self.num_batches):
            xis, xjs = make_pair_batch(self.batch_size, seed=base + k)
            yield xis.pin_memory(), xjs.pin_memory()


class SyntheticMoleculeDatasetWrapper:
    """Same constructor as dataset.MoleculeDatasetWrapper (dataset.py:153-159).  ``data_path = "synthetic:<count>"`` sets the
    number of molecules; the train / validation split follows ``valid_size`` and both loaders drop the last ragged batch
    (dataset.py:176-184)."""

    def __init__(self, batch_size, num_workers, valid_size, data_path):
        self.batch_size, self.num_workers, self.valid_size = batch_size, num_workers, valid_size
        self.num_molecules = int(str(data_path).split(":", 1)[1]) if str(data_path).startswith("synthetic:") else 20000

    def get_data_loaders(self):
        n_valid = int(math.floor(self.valid_size * self.num_molecules))
        n_train = self.num_molecules - n_valid
        return (_Loader(self.batch_size, n_train // self.batch_size, seed=1, reshuffle=True),
                _Loader(self.batch_size, max(n_valid // self.batch_size, 1), seed=900_000_007, reshuffle=False))


class _DeviceLoader:
    """Iterable of (xis, xjs) built ON THE GPU from a packed store: ids = a fresh permutation of the subset every epoch
    (SubsetRandomSampler, dataset.py:166-177), views from ``augment_pair`` (one kernel per batch, no host batch at all)."""

    def __init__(self, store, ids, batch_size, seed, reshuffle, aug="node"):
        self.store, self.ids, self.batch_size, self.seed, self.reshuffle, self.epoch = store, ids, batch_size, seed, reshuffle, 0
        self.aug = aug

    def __len__(self):
        return len(self.ids) // self.batch_size                 # drop_last=True (dataset.py:179-184)

    def __iter__(self):
        import numpy as np
        from .dataset import augment_pair
        e = self.epoch if self.reshuffle else 0
        self.epoch += 1
        order = np.random.default_rng(self.seed + e).permutation(self.ids) if self.reshuffle else self.ids
        for k in range(len(self)):
            yield augment_pair(self.store, order[k * self.batch_size:(k + 1) * self.batch_size], seed=(self.seed + e) * 1_000_003 + k,
                               aug=self.aug)


class PackedMoleculeDatasetWrapper:
    """``MoleculeDatasetWrapper`` (dataset.py:153-184) over a packed molecule store resident in HBM.  ``data_path`` is a
    ``molclr-packed v1`` .npz file (``PackedMolecules.save``) or ``"synthetic:<count>"``; the train / validation split by
    ``valid_size`` over a seeded shuffle of the indices follows dataset.py:166-175."""

    def __init__(self, batch_size, num_workers, valid_size, data_path, device="cuda:0", seed=0, aug="node"):
        import numpy as np
        if aug not in ("node", "subgraph", "mix"):
            raise ValueError("Not defined molecule augmentation!")                    # molclr.py:186-191
        self.aug = aug
        from .dataset import PackedMolecules
        from .synth import random_molecule
        self.batch_size, self.num_workers, self.valid_size = batch_size, num_workers, valid_size
        if str(data_path).startswith("synthetic:"):
            rng = np.random.default_rng(seed)
            store = PackedMolecules.from_graphs(random_molecule(rng) for _ in range(int(str(data_path).split(":", 1)[1])))
        else:
            store = PackedMolecules.load(data_path)
        self.store = store.to(device)
        idx = np.random.default_rng(seed + 1).permutation(len(store))
        split = int(math.floor(valid_size * len(store)))
        self.valid_idx, self.train_idx = idx[:split], idx[split:]

    def get_data_loaders(self):
        return (_DeviceLoader(self.store, self.train_idx, self.batch_size, seed=1, reshuffle=True, aug=self.aug),
                _DeviceLoader(self.store, self.valid_idx, self.batch_size, seed=900_000_007, reshuffle=False, aug=self.aug))


class MolCLR:
    def __init__(self, dataset, config, log_root="ckpt"):
        self.config, self.dataset = config, dataset
        if not torch.cuda.is_available() or config["gpu"] == "cpu":
            raise RuntimeError("molclr_b200.trainer needs a CUDA device (the kernels have no CPU path)")
        self.device = torch.device(config["gpu"])
        torch.cuda.set_device(self.device)
        self.log_dir = os.path.join(log_root, time.strftime("%b%d_%H-%M-%S"))
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(log_dir=self.log_dir)
        except Exception:                                   # tensorboard is optional here
            self.writer = None
        self.nt_xent_criterion = NTXentLoss(self.device, config["batch_size"], **config["loss"])

    # the hot path: molclr.py:55-67
    def _step(self, model, xis, xjs, n_iter=None):
        # two encoder passes (view i, then view j), F.normalize, NT-Xent: molclr_b200.pretrain_loss runs the two passes as one
        # autograd node when the model offers forward_pair (same values)
        return pretrain_loss(model, self.nt_xent_criterion, xis, xjs)

    def _scalar(self, tag, value, step):
        if self.writer is not None:
            self.writer.add_scalar(tag, value, global_step=step)

    def _build_model(self):
        kind = self.config["model_type"]
        if kind not in ("gin", "gcn"):
            raise ValueError("Undefined GNN model.")
        model = (GINet if kind == "gin" else GCN)(**self.config["model"]).to(self.device)
        return self._load_pre_trained_weights(model)

    def train(self, max_steps_per_epoch=None):
        cfg = self.config
        train_loader, valid_loader = self.dataset.get_data_loaders()
        model = self._build_model()
        weight_decay = cfg["weight_decay"]
        weight_decay = float(eval(weight_decay)) if isinstance(weight_decay, str) else float(weight_decay)   # molclr.py:86
        optimizer = torch.optim.Adam(model.parameters(), cfg["init_lr"], weight_decay=weight_decay, fused=True)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=max(cfg["epochs"] - cfg["warm_up"], 1), eta_min=0)
        ckpt_dir = os.path.join(self.log_dir, "checkpoints")
        os.makedirs(ckpt_dir, exist_ok=True)
        if os.path.exists("./config.yaml"):
            shutil.copy("./config.yaml", os.path.join(ckpt_dir, "config.yaml"))
        n_iter, valid_n_iter, best_valid, history = 0, 0, float("inf"), []
        for epoch in range(cfg["epochs"]):
            for bn, (xis, xjs) in enumerate(train_loader):
                if max_steps_per_epoch is not None and bn >= max_steps_per_epoch:
                    break
                optimizer.zero_grad(set_to_none=True)
                loss = self._step(model, xis.to(self.device, non_blocking=True), xjs.to(self.device, non_blocking=True), n_iter)
                if n_iter % cfg["log_every_n_steps"] == 0:
                    self._scalar("train_loss", loss.item(), n_iter)
                    self._scalar("cosine_lr_decay", scheduler.get_last_lr()[0], n_iter)
                    print(epoch, bn, loss.item())
                loss.backward()
                optimizer.step()
                n_iter += 1
            poll_checks(block=True)                        # deferred batch validation (graph.py): raise here at the latest
            model.check_fp16_range()                       # ... and the operand-range check of precision "fp16x3"
            if epoch % cfg["eval_every_n_epochs"] == 0:
                valid_loss = self._validate(model, valid_loader)
                print(epoch, valid_loss, "(validation)")
                history.append(valid_loss)
                if valid_loss < best_valid:
                    best_valid = valid_loss
                    torch.save(model.state_dict(), os.path.join(ckpt_dir, "model.pth"))
                self._scalar("validation_loss", valid_loss, valid_n_iter)
                valid_n_iter += 1
            if (epoch + 1) % cfg["save_every_n_epochs"] == 0:
                torch.save(model.state_dict(), os.path.join(ckpt_dir, f"model_{epoch}.pth"))
            if epoch >= cfg["warm_up"]:
                scheduler.step()
        return model, history

    def _load_pre_trained_weights(self, model):
        path = os.path.join("./ckpt", str(self.config["load_model"]), "checkpoints", "model.pth")
        if os.path.exists(path):
            model.load_state_dict(torch.load(path, map_location=self.device))
            print("Loaded pre-trained model with success.")
        else:
            print("Pre-trained weights not found. Training from scratch.")
        return model

    def _validate(self, model, valid_loader):
        model.eval()
        total, count = 0.0, 0
        with torch.no_grad():
            for xis, xjs in valid_loader:
                total += self._step(model, xis.to(self.device), xjs.to(self.device)).item()
                count += 1
        model.train()
        return total / max(count, 1)


def build_dataset(config):
    """The dataset wrapper a config selects.  Nothing is substituted silently: ``synthetic:<count>`` and ``molclr-packed v1``
    .npz stores are served; a SMILES file (the reference's ``data/pubchem-10m-clean.txt``) needs RDKit, which is not a dependency
    here -- convert it offline to a packed store (``PackedMolecules.from_graphs(...).save``)."""
    if config["aug"] not in ("node", "subgraph", "mix"):
        raise ValueError("Not defined molecule augmentation!")                        # molclr.py:186-191
    if config.get("fp16_precision"):
        raise ValueError("molclr_b200.trainer: fp16_precision (apex AMP, molclr.py:93-98) is not supported: the kernels choose their "
                         "own operand precision (model.precision = 'fp16x3' | 'tf32x3' | 'tf32')")
    data_path = str(config["dataset"]["data_path"])
    if data_path.startswith("synthetic:") and config["aug"] == "node" and not config["dataset"].get("packed"):
        return SyntheticMoleculeDatasetWrapper(config["batch_size"], **{k: v for k, v in config["dataset"].items() if k != "packed"})
    if data_path.startswith("synthetic:") or data_path.endswith(".npz"):
        kw = {k: v for k, v in config["dataset"].items() if k != "packed"}
        return PackedMoleculeDatasetWrapper(config["batch_size"], device=config["gpu"], aug=config["aug"], **kw)
    raise ValueError(f"molclr_b200.trainer: data_path {data_path!r} is neither 'synthetic:<count>' nor a molclr-packed v1 .npz store; "
                     "SMILES text needs RDKit (not available here): convert it offline with PackedMolecules.from_graphs(...).save(path)")


def main(argv=None):
    import argparse
    import copy
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default=None)
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--steps-per-epoch", type=int, default=None)
    args = ap.parse_args(argv)
    config = copy.deepcopy(DEFAULT_CONFIG)
    if args.config:
        import yaml
        config.update(yaml.load(open(args.config), Loader=yaml.FullLoader))
    if args.epochs is not None:
        config["epochs"] = args.epochs
    dataset = build_dataset(config)
    MolCLR(dataset, config).train(max_steps_per_epoch=args.steps_per_epoch)


if __name__ == "__main__":
    main(sys.argv[1:])
