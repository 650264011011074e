"""Generates the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Run in the dev container only (needs /root/reference):  python tests/golden/make_golden.py

* ntxent_*.npz           -- inputs, loss and input-gradients of the reference's own
                            ``utils/nt_xent.py:NTXentLoss`` (imported, not restated).
* gcn_ckpt_manifest.json -- key/shape/dtype listing of the only checkpoint the reference
                            ships (``ckpt/pretrained_gcn/checkpoints/model.pth``); pins the
                            state_dict layout of the GCN drop-in.
"""
import json
import zlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ntxent_cases():
    sys.path.insert(0, REF)
    from utils.nt_xent import NTXentLoss          # the reference class itself
    cases = [  # name, N, C, tau, cosine, pre-normalised inputs (molclr.py:63-64 does normalise)
        ("ntxent_n8_c16_cos", 8, 16, 0.1, True, True),
        ("ntxent_n64_c256_cos", 64, 256, 0.1, True, True),
        ("ntxent_n48_c40_cos_raw", 48, 40, 0.5, True, False),
        ("ntxent_n32_c64_dot", 32, 64, 0.1, False, True),
        ("ntxent_n128_c256_cos", 128, 256, 0.1, True, True),
    ]
    for name, n, c, tau, cos, norm in cases:
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % (2 ** 31))
        zis = torch.randn(n, c, generator=g)
        zjs = 0.5 * zis + torch.randn(n, c, generator=g)      # correlated views
        if norm:
            zis = torch.nn.functional.normalize(zis, dim=1)
            zjs = torch.nn.functional.normalize(zjs, dim=1)
        zis.requires_grad_(True); zjs.requires_grad_(True)
        crit = NTXentLoss("cpu", n, tau, cos)
        loss = crit(zis, zjs)
        loss.backward()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), zis=zis.detach().numpy(), zjs=zjs.detach().numpy(),
                            loss=loss.detach().numpy(), dzis=zis.grad.numpy(), dzjs=zjs.grad.numpy(),
                            temperature=np.float64(tau), cosine=np.bool_(cos), batch_size=np.int64(n))
        print(name, float(loss.detach()))


def gcn_manifest():
    sd = torch.load(os.path.join(REF, "ckpt/pretrained_gcn/checkpoints/model.pth"), map_location="cpu")
    man = {k: {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")} for k, v in sd.items()}
    extra = {"num_batches_tracked.0": int(sd["batch_norms.0.num_batches_tracked"])}
    with open(os.path.join(HERE, "gcn_ckpt_manifest.json"), "w") as f:
        json.dump({"entries": man, "facts": extra}, f, indent=1, sort_keys=True)
    print("gcn manifest:", len(man), "entries")


if __name__ == "__main__":
    ntxent_cases()
    gcn_manifest()
