"""Golden vectors for BASELINE config 3's realistic-weights case: the reference's own, unmodified `GCN` class
(models/gcn_molclr.py) loaded -- strict -- with the checkpoint the reference ships
(ckpt/pretrained_gcn/checkpoints/model.pth) and run through `MolCLR._step` (molclr.py:55-67) with the reference's own
`NTXentLoss`, train mode + backward, then an eval-mode forward.

Run in the dev container only (needs /root/reference):  python tests/golden/make_gcn_ckpt_golden.py

The fixture carries the checkpoint's tensors (4 MB: the GPU box has no /root/reference), the two input batches, the loss, the
projections, running statistics after the step, and every parameter gradient (norm + strided sample for the large ones).
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
from models.gcn_molclr import GCN as RefGCN                  # noqa: E402  (the reference class itself)
from utils.nt_xent import NTXentLoss                         # noqa: E402

from molclr_b200.synth import make_pair_batch                # noqa: E402
from make_encoder_golden import batch_arrays, grad_arrays    # noqa: E402

if __name__ == "__main__":
    bs = 64
    sd = torch.load(os.path.join(REF, "ckpt", "pretrained_gcn", "checkpoints", "model.pth"), map_location="cpu")
    model = RefGCN(5, 300, 512, 0, "mean")
    model.load_state_dict(sd, strict=True)
    model.train()
    bi, bj = make_pair_batch(bs, seed=77)
    ris, zis = model(bi)
    rjs, zjs = model(bj)
    loss = NTXentLoss("cpu", bs, 0.1, True)(F.normalize(zis, dim=1), F.normalize(zjs, dim=1))
    loss.backward()
    out = {"batch_size": np.int64(bs), "loss": loss.detach().numpy(), "h_i": ris.detach().numpy(), "out_i": zis.detach().numpy(),
           "h_j": rjs.detach().numpy(), "out_j": zjs.detach().numpy()}
    out.update({f"state.{k}": v.numpy() for k, v in sd.items()})
    out.update(batch_arrays(bi, "i")); out.update(batch_arrays(bj, "j")); out.update(grad_arrays(model))
    for l in range(5):
        out[f"running_mean.{l}"] = model.batch_norms[l].running_mean.numpy().copy()
        out[f"running_var.{l}"] = model.batch_norms[l].running_var.numpy().copy()
    model.eval()
    with torch.no_grad():
        he, oe = model(bi)
    out["h_i_eval"], out["out_i_eval"] = he.numpy(), oe.numpy()
    np.savez_compressed(os.path.join(HERE, "enc_gcn_ckpt_pretrain.npz"), **out)
    print("enc_gcn_ckpt_pretrain", float(loss), sum(v.nbytes for v in out.values()) // 1024, "KB raw")
