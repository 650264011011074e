"""Golden vectors for the ENCODER path, produced by running the reference's own, unmodified model classes.

Run in the dev container only (needs /root/reference):  python tests/golden/make_encoder_golden.py

`models/ginet_molclr.py`, `models/gcn_molclr.py`, `models/ginet_finetune.py` and `models/gcn_finetune.py` are imported from /root/reference as they
are; their third-party base (torch-geometric 1.6.3 / torch-scatter 2.0.6, not vendored, not installable offline) is
provided by the restatement in `pyg163_stub.py` (MessagePassing.propagate = index_select + scatter_add_, add_self_loops,
global_*_pool).  Every line of the reference's GINEConv / GCNConv / GINet / GCN / fine-tune GINet -- self-loop attributes,
edge embeddings, message, MLP, BatchNorm, ReLU/dropout placement, pooling, heads -- is therefore executed, not restated.
The loss is `MolCLR._step` (molclr.py:55-67) with the reference's own `utils/nt_xent.py:NTXentLoss`, or the fine-tune
criteria of finetune.py:70-77 (CrossEntropyLoss / MSELoss).

Weights are a deterministic function of (parameter name, shape, seed) -- tests/util.py:golden_weights -- so fixtures hold
only inputs and results.  Large gradients are stored as their norm plus a strided sample.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import pyg163_stub                      # noqa: E402
pyg163_stub.install()
sys.path.insert(0, REF)
from models.ginet_molclr import GINet as RefGINet            # noqa: E402  (the reference classes themselves)
from models.gcn_molclr import GCN as RefGCN                  # noqa: E402
from models.ginet_finetune import GINet as RefGINetFinetune  # noqa: E402
from models.gcn_finetune import GCN as RefGCNFinetune        # noqa: E402
from models.ginet_finetune_mp import GINet as RefGINetMotif   # noqa: E402
from utils.nt_xent import NTXentLoss                         # noqa: E402

from molclr_b200.synth import make_pair_batch, make_plain_batch   # noqa: E402
from tests.util import golden_weights                             # noqa: E402

SAMPLE = 4096     # gradients with more elements are stored as (norm, strided sample)


def batch_arrays(b, tag):
    return {f"{tag}_x": b.x.numpy(), f"{tag}_edge_index": b.edge_index.numpy(), f"{tag}_edge_attr": b.edge_attr.numpy(),
            f"{tag}_batch": b.batch.numpy()}


def grad_arrays(model):
    out = {}
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach()
        out[f"gradnorm.{k}"] = np.float64(g.double().norm().item())
        flat = g.reshape(-1)
        if flat.numel() <= SAMPLE:
            out[f"grad.{k}"] = g.numpy()
        else:
            step = flat.numel() // SAMPLE
            out[f"gradsample.{k}"] = flat[::step][:SAMPLE].numpy()
    return out


def layer_taps(model, rows=48):
    """Forward hooks (they do not alter the reference): first `rows` rows of every conv output and BatchNorm output."""
    taps, hooks = {}, []

    def tap(name):
        def hook(_module, _inputs, output):
            taps.setdefault(name, output.detach()[:rows].numpy().copy())      # returns None: the output is not replaced
        return hook
    for l, (conv, bn) in enumerate(zip(model.gnns, model.batch_norms)):
        hooks.append(conv.register_forward_hook(tap(f"conv{l}")))
        hooks.append(bn.register_forward_hook(tap(f"bn{l}")))
    return taps, hooks


def pretrain_case(name, cls, bs, seed, wseed):
    """MolCLR._step (molclr.py:55-67) in train mode, backward, then an eval-mode forward of view i."""
    model = cls(5, 300, 512, 0, "mean")
    model.load_state_dict(golden_weights(model.state_dict(), wseed))
    model.train()
    bi, bj = make_pair_batch(bs, seed=seed)
    taps, hooks = layer_taps(model)
    ris, zis = model(bi)
    for h in hooks:
        h.remove()
    rjs, zjs = model(bj)
    zis_n, zjs_n = F.normalize(zis, dim=1), F.normalize(zjs, dim=1)
    loss = NTXentLoss("cpu", bs, 0.1, True)(zis_n, zjs_n)
    loss.backward()
    out = {"batch_size": np.int64(bs), "weight_seed": np.int64(wseed), "loss": loss.detach().numpy(), "h_i": ris.detach().numpy(),
           "out_i": zis.detach().numpy(), "h_j": rjs.detach().numpy(), "out_j": zjs.detach().numpy()}
    out.update({f"tap_i.{k}": v for k, v in taps.items()})
    out.update(batch_arrays(bi, "i")); out.update(batch_arrays(bj, "j")); out.update(grad_arrays(model))
    for l in (0, 4):
        out[f"running_mean.{l}"] = model.batch_norms[l].running_mean.numpy().copy()
        out[f"running_var.{l}"] = model.batch_norms[l].running_var.numpy().copy()
    out["num_batches_tracked"] = np.int64(model.batch_norms[0].num_batches_tracked.item())
    model.eval()
    with torch.no_grad():
        he, oe = model(bi)
    out["h_i_eval"], out["out_i_eval"] = he.numpy(), oe.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, float(loss), sum(v.nbytes for v in out.values()) // 1024, "KB raw")


def small_case(name, cls, pool, graphs, seed, wseed, layers=3, emb=64, feat=64):
    """All three pooling modes on a small model: every gradient stored in full."""
    model = cls(layers, emb, feat, 0, pool)
    model.load_state_dict(golden_weights(model.state_dict(), wseed))
    model.train()
    b = make_plain_batch(graphs, seed=seed)
    taps, hooks = layer_taps(model, rows=96)
    h, o = model(b)
    for hk in hooks:
        hk.remove()
    loss = o.square().sum() + 0.5 * h.sum()
    loss.backward()
    out = {"weight_seed": np.int64(wseed), "layers": np.int64(layers), "emb": np.int64(emb), "feat": np.int64(feat), "pool": np.str_(pool),
           "loss": loss.detach().numpy(), "h": h.detach().numpy(), "out": o.detach().numpy()}
    out.update({f"tap.{k}": v for k, v in taps.items()})
    out.update(batch_arrays(b, "b")); out.update(grad_arrays(model))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, float(loss))


def finetune_case(name, task, graphs, seed, wseed, gcn=False):
    """models/ginet_finetune.py GINet (or models/gcn_finetune.py GCN) + the criteria of finetune.py:70-77 (classification:
    CrossEntropyLoss on data.y.flatten(); regression: MSELoss), drop_ratio 0 (dropout is not reproducible bit for bit)."""
    model = RefGCNFinetune(task, 5, 300, 256, 0, "mean") if gcn else RefGINetFinetune(task, 5, 300, 512, 0, "mean")
    model.load_state_dict(golden_weights(model.state_dict(), wseed))
    model.train()
    b = make_plain_batch(graphs, seed=seed, mean_atoms=46.0 if task == "classification" else 26.0, std_atoms=18.0 if task == "classification" else 13.0)
    g = torch.Generator().manual_seed(seed)
    if task == "classification":
        y = (torch.rand(graphs, 1, generator=g) < 0.77).long()
        crit = torch.nn.CrossEntropyLoss()
    else:
        y = -3.05 + 2.1 * torch.randn(graphs, 1, generator=g)
        crit = torch.nn.MSELoss()
    h, pred = model(b)
    loss = crit(pred, y.flatten()) if task == "classification" else crit(pred, y)
    loss.backward()
    out = {"weight_seed": np.int64(wseed), "task": np.str_(task), "y": y.numpy(), "loss": loss.detach().numpy(), "h": h.detach().numpy(),
           "pred": pred.detach().numpy()}
    out.update(batch_arrays(b, "b")); out.update(grad_arrays(model))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, float(loss))


def motif_case(name, task, graphs, num_motifs, seed, wseed):
    """models/ginet_finetune_mp.py GINet (motif embedding + GlobalAttention over [cliques of the molecule; the molecule]) with
    ``mol_idx`` / ``clique_idx`` built as finetune.py:202-210 builds them, and the criteria of finetune.py:70-77."""
    model = RefGINetMotif(num_motifs, task, 5, 300, 512, 0, "mean")
    model.load_state_dict(golden_weights(model.state_dict(), wseed))
    model.train()
    b = make_plain_batch(graphs, seed=seed, mean_atoms=30.0, std_atoms=10.0)
    g = torch.Generator().manual_seed(seed)
    mol_idx, clique_idx = [], []
    for i in range(graphs):                                    # every molecule owns 1..4 motifs (finetune.py:204-207)
        for c in torch.randperm(num_motifs, generator=g)[:int(torch.randint(1, 5, (1,), generator=g))].tolist():
            mol_idx.append(i)
            clique_idx.append(c)
    mol_idx.extend([i for i in range(max(mol_idx) + 1)])       # finetune.py:208
    mol_idx, clique_idx = torch.tensor(mol_idx), torch.tensor(clique_idx)
    if task == "classification":
        y = (torch.rand(graphs, 1, generator=g) < 0.77).long()
        crit = torch.nn.CrossEntropyLoss()
    else:
        y = -3.05 + 2.1 * torch.randn(graphs, 1, generator=g)
        crit = torch.nn.MSELoss()
    h, pred = model(b, mol_idx, clique_idx)
    loss = crit(pred, y.flatten()) if task == "classification" else crit(pred, y)
    loss.backward()
    out = {"weight_seed": np.int64(wseed), "task": np.str_(task), "num_motifs": np.int64(num_motifs), "y": y.numpy(), "loss": loss.detach().numpy(),
           "h": h.detach().numpy(), "pred": pred.detach().numpy(), "mol_idx": mol_idx.numpy(), "clique_idx": clique_idx.numpy()}
    out.update(batch_arrays(b, "b")); out.update(grad_arrays(model))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, float(loss))


if __name__ == "__main__":
    torch.manual_seed(0)
    pretrain_case("enc_gin_pretrain", RefGINet, 24, seed=11, wseed=1)
    pretrain_case("enc_gcn_pretrain", RefGCN, 24, seed=12, wseed=2)
    for pool in ("mean", "add", "max"):
        small_case(f"enc_gin_small_{pool}", RefGINet, pool, 14, seed=20, wseed=3)
    small_case("enc_gcn_small_max", RefGCN, "max", 14, seed=21, wseed=4)
    small_case("enc_gcn_small_mean", RefGCN, "mean", 14, seed=22, wseed=5)
    finetune_case("enc_finetune_cls", "classification", 12, seed=30, wseed=6)
    finetune_case("enc_finetune_reg", "regression", 12, seed=31, wseed=7)
    finetune_case("enc_gcn_finetune_cls", "classification", 12, seed=32, wseed=8, gcn=True)
    finetune_case("enc_gcn_finetune_reg", "regression", 12, seed=33, wseed=9, gcn=True)
    motif_case("enc_motif_cls", "classification", 12, 20, seed=40, wseed=10)
    motif_case("enc_motif_reg", "regression", 12, 20, seed=41, wseed=11)
