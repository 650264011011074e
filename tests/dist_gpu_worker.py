"""torchrun worker of tests/test_dist_gpu.py: one rank of a data-parallel MolCLR step on real GPUs over NCCL (no oracle here)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path, B, global_neg, overlap = sys.argv[1], int(sys.argv[2]), sys.argv[3] == "1", sys.argv[4] == "1"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from molclr_b200 import GINet
    from molclr_b200.dist import DataParallelStep
    from molclr_b200.synth import make_pair_batch
    torch.manual_seed(1234 + rank)                        # DIFFERENT initial weights per rank: the constructor must broadcast rank 0's
    model = GINet(5, 300, 512, 0, "mean").to(dev)
    with torch.no_grad():
        for bn in model.batch_norms:
            bn.weight.uniform_(0.5, 1.5); bn.bias.uniform_(-0.3, 0.3)
    stepper = DataParallelStep(model, B, 0.1, True, global_negatives=global_neg, overlap=overlap)
    state0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    bi, bj = make_pair_batch(B, seed=900 + rank)
    loss = stepper.loss(bi.to(dev), bj.to(dev))
    loss.backward()
    stepper.allreduce_gradients()
    total = stepper.global_loss(loss)
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()}
    # every rank must hold the same reduced gradients
    flat = torch.cat([g.reshape(-1) for g in grads.values()]).to(dev)
    others = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(others, flat)
    same = all(torch.equal(o, flat) for o in others)
    if rank == 0:
        torch.save({"loss": float(total), "share": float(loss), "grads": grads, "state0": state0, "same_on_all_ranks": same, "world": world}, out_path)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
