"""TEST INFRASTRUCTURE (it lives under tests/ because it runs the oracle; pytest does not collect it).
Per-parameter parity report (GPU): the full MolCLR pre-training step (molclr.py:55-67) on the CUDA path against the CPU
oracle in fp32 AND in fp64, so that every measured error is printed next to its floor (oracle fp32 vs oracle fp64).

    python tests/parity_report.py [--batches 128,512] [--precision fp16x3|tf32x3|tf32] [--out profiles/parity_r2.json]

Also checks (a) run-twice bit-reproducibility of loss and gradients and (b) that an optimizer step on the parameters is
picked up by the next forward.  The JSON it writes is what the test tolerances are set from (tests/test_gpu_config_sizes.py).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))          # repo root
import torch

from tests.util import rel_err, max_rel, sync_oracle_from
from molclr_b200 import GCN, GINet, NTXentLoss, pretrain_loss
from molclr_b200.synth import make_pair_batch
from oracle import gnn as ognn
from oracle.nt_xent import ntxent_closed_form
from oracle.step import pretrain_loss as oracle_pretrain_loss

DEV = "cuda:0"


class ClosedForm(torch.nn.Module):
    """The loss of nt_xent.py:47-65 through the chunked closed form (same value; O(2N chunk) memory), in the input dtype's
    fp64 promotion -- usable at 4096 pairs where the reference's [2N,2N,C] broadcast needs 68.7 GB."""

    def __init__(self, tau, cos):
        super().__init__()
        self.tau, self.cos = tau, cos

    def forward(self, zis, zjs):
        return ntxent_closed_form(zis, zjs, self.tau, self.cos)


def build(model_name, seed, precision):
    torch.manual_seed(seed)
    cls, ocls = (GINet, ognn.GINet) if model_name == "gin" else (GCN, ognn.GCN)
    m = cls(5, 300, 512, 0, "mean").to(DEV)
    m.precision = precision
    with torch.no_grad():
        for bn in m.batch_norms:
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    o32 = sync_oracle_from(m, ocls(5, 300, 512, 0, "mean"))
    o64 = ocls(5, 300, 512, 0, "mean").double()
    o64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in o32.state_dict().items()})
    return m, o32, o64


def one_case(model_name, bs, precision, tau=0.1):
    m, o32, o64 = build(model_name, 3, precision)
    bi, bj = make_pair_batch(bs, seed=21)
    t0 = time.time()
    loss = pretrain_loss(m, NTXentLoss(DEV, bs, tau, True), bi.to(DEV), bj.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    l32 = oracle_pretrain_loss(o32, ClosedForm(tau, True), bi, bj)
    l32.backward()
    l64 = oracle_pretrain_loss(o64, ClosedForm(tau, True), bi, bj)
    l64.backward()
    rows = {}
    for (k, p), (_, q), (_, r) in zip(m.named_parameters(), o32.named_parameters(), o64.named_parameters()):
        rows[k] = {"cuda_vs_fp64": rel_err(p.grad, r.grad), "cuda_vs_fp32": rel_err(p.grad, q.grad), "fp32_vs_fp64": rel_err(q.grad, r.grad),
                   "norm": float(r.grad.norm())}
    res = {"model": model_name, "pairs": bs, "precision": precision, "nodes": int(bi.x.size(0)), "seconds": time.time() - t0,
           "loss": {"cuda": float(loss), "fp32": float(l32), "fp64": float(l64),
                    "cuda_vs_fp64": abs(float(loss) - float(l64)) / abs(float(l64)), "fp32_vs_fp64": abs(float(l32) - float(l64)) / abs(float(l64))},
           "grads": rows}
    worst = max(((v["cuda_vs_fp64"], k) for k, v in rows.items() if not k.endswith("mlp.2.bias") and v["norm"] > 1e-12), default=(0, ""))
    floor = max(((v["fp32_vs_fp64"], k) for k, v in rows.items() if not k.endswith("mlp.2.bias") and v["norm"] > 1e-12), default=(0, ""))
    res["worst"], res["worst_floor"] = {"key": worst[1], "err": worst[0]}, {"key": floor[1], "err": floor[0]}
    return res


def reproducibility(bs, precision):
    """Same inputs, same weights, two runs: are loss and every gradient bit-identical?"""
    m, _, _ = build("gin", 3, precision)
    bi, bj = make_pair_batch(bs, seed=21)
    bi, bj = bi.to(DEV), bj.to(DEV)
    runs = []
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        for bn in m.batch_norms:
            bn.reset_running_stats()
        from molclr_b200 import Batch
        fresh = lambda b: Batch(b.x, b.edge_index, b.edge_attr, b.batch, b.num_graphs)
        loss = pretrain_loss(m, NTXentLoss(DEV, bs, 0.1, True), fresh(bi), fresh(bj))
        loss.backward()
        runs.append((loss.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters()}))
    diff = [k for k in runs[0][1] if not torch.equal(runs[0][1][k], runs[1][1][k])]
    return {"pairs": bs, "loss_bit_identical": bool(torch.equal(runs[0][0], runs[1][0])), "grads_not_bit_identical": diff}


def weight_update_pickup(precision):
    """An in-place optimizer step (fused Adam does not bump tensor versions) must change the next forward."""
    m, _, _ = build("gin", 3, precision)
    bi, _ = make_pair_batch(16, seed=5)
    bi = bi.to(DEV)
    opt = torch.optim.Adam(m.parameters(), 1e-2, fused=True)
    m.eval()
    with torch.no_grad():
        out0 = m(bi)[1].clone()
    m.train()
    h, out = m(bi)
    out.square().sum().backward()
    opt.step()
    w_after = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.eval()
    with torch.no_grad():
        out1 = m(bi)[1].clone()
    # oracle with the updated weights
    o = ognn.GINet(5, 300, 512, 0, "mean")
    o.load_state_dict({k: v.cpu() for k, v in w_after.items()})
    o.eval()
    with torch.no_grad():
        want = o(bi.to("cpu"))[1]
    return {"output_changed": float((out1 - out0).abs().max()), "matches_oracle_with_updated_weights": max_rel(out1, want)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="128,512")
    ap.add_argument("--precision", default="fp16x3")
    ap.add_argument("--models", default="gin")
    ap.add_argument("--out", default="gpurun_out/parity.json")
    a = ap.parse_args()
    report = {"device": torch.cuda.get_device_name(0), "cases": []}
    for mn in a.models.split(","):
        for bs in [int(b) for b in a.batches.split(",")]:
            r = one_case(mn, bs, a.precision)
            report["cases"].append(r)
            print(f"== {mn} {bs} pairs ({a.precision}): loss cuda {r['loss']['cuda']:.6f} fp64 {r['loss']['fp64']:.6f} rel {r['loss']['cuda_vs_fp64']:.2e} "
                  f"(floor {r['loss']['fp32_vs_fp64']:.2e}); worst grad {r['worst']['key']} {r['worst']['err']:.2e}; worst floor {r['worst_floor']['key']} {r['worst_floor']['err']:.2e}",
                  flush=True)
            for k, v in r["grads"].items():
                print(f"   {k:40s} cuda-vs-fp64 {v['cuda_vs_fp64']:.2e}  cuda-vs-fp32 {v['cuda_vs_fp32']:.2e}  floor(fp32-vs-fp64) {v['fp32_vs_fp64']:.2e}  |g| {v['norm']:.2e}")
    report["reproducibility"] = reproducibility(256, a.precision)
    print("reproducibility:", report["reproducibility"], flush=True)
    report["weight_update"] = weight_update_pickup(a.precision)
    print("weight update pickup:", report["weight_update"], flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)
