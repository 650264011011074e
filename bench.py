#!/usr/bin/env python
"""Benchmark of the MolCLR pre-training hot path (BASELINE.json metric: molecules/s; GINE aggregation GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = the loop body of molclr.py:109-127 on one batch of B synthetic molecule pairs per GPU:
zero_grad -> CSR plan of both views -> 2 encoder forwards -> normalize -> NT-Xent -> backward -> Adam.
Prints ONE JSON line (see the task contract): `value` with inputs resident in HBM, `e2e` through the public
API with pinned HOST batches (H2D copies + loss read-back inside the timed region), `roofline` for the GINE
aggregation kernel timed in situ, and `cpu_baseline` (the oracle port on the host cores, bounded sample).
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "GIN-5 emb300 feat512 MolCLR pretrain fwd+bwd+Adam, batch 4096 pairs/GPU of synthetic ~25-atom molecules " \
           "(atom-mask/bond-delete augmented), NT-Xent tau=0.1 cosine"
METRIC = "MolCLR GIN pretrain molecules/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="molecule pairs per GPU per step")
    ap.add_argument("--cpu-batch", type=int, default=512, help="pairs per step of the CPU baseline sample")
    ap.add_argument("--precision", default=None, choices=["fp16x3", "tf32x3", "tf32"], help="default: the package default (molclr_b200.GINet.precision)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the few-step measurements of BASELINE configs 1 (on the GPU), 3 and 4")
    ap.add_argument("--local-negatives", action="store_true", help="N>1: NT-Xent over the local batch only")
    ap.add_argument("--same-batches", action="store_true", help="N>1 diagnostic: every rank processes the SAME batches (no rank skew from batch sizes)")
    ap.add_argument("--dp-bucket", type=int, default=None, help="N>1: gradient all-reduce bucket size in floats (0 = no overlap: one all-reduce after the backward)")
    ap.add_argument("--no-pdl", action="store_true", help="launch every kernel in plain stream order (no programmatic dependent launch): A/B timing")
    ap.add_argument("--model", default="gin", choices=["gin", "gcn"], help="gin = BASELINE configs 1/2/5 (headline), gcn = config 3")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons |= {k for k, bit in names.items() if r & bit}
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------------- CPU / reference arm
def host_cores():
    """Host threads this process may use: torchrun exports OMP_NUM_THREADS=1 for its workers, which would silently turn the
    "all host cores" baseline into a single-threaded one -- so the count comes from the scheduler affinity, not the environment."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(batch, steps, warmup, threads=None):
    """The reference path on the host cores: oracle restatement of the PyG encoder + the reference's NT-Xent
    formulation (oracle/), full step incl. Adam (molclr.py:109-127).  Returns (molecules/s, threads, s/step)."""
    from oracle import gnn as ognn
    from oracle.nt_xent import NTXentRestated
    from oracle.step import train_step
    from molclr_b200.synth import make_pair_batch
    torch.set_num_threads(threads or host_cores())
    torch.manual_seed(0)
    model = ognn.GINet(5, 300, 512, 0, "mean")
    crit = NTXentRestated("cpu", batch, 0.1, True)
    opt = torch.optim.Adam(model.parameters(), 5e-4, weight_decay=1e-5)
    data = [make_pair_batch(batch, seed=100 + i) for i in range(2)]
    for i in range(warmup):
        train_step(model, crit, opt, *data[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        train_step(model, crit, opt, *data[i % 2])
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return batch / dt, torch.get_num_threads(), dt


def cpu_reference_split(batch):
    """Where the reference step spends its time on the host: one forward + backward of the two encoder passes
    alone (loss = sum of squares of the projections) and one forward + backward of the reference's NT-Xent formulation alone on
    random unit projections.  Returns seconds per step of each part."""
    from oracle import gnn as ognn
    from oracle.nt_xent import NTXentRestated
    from molclr_b200.synth import make_pair_batch
    torch.manual_seed(0)
    model = ognn.GINet(5, 300, 512, 0, "mean")
    bi, bj = make_pair_batch(batch, seed=100)
    t0 = time.perf_counter()
    (model(bi)[1].square().sum() + model(bj)[1].square().sum()).backward()
    t_enc = time.perf_counter() - t0
    crit = NTXentRestated("cpu", batch, 0.1, True)
    z = torch.nn.functional.normalize(torch.randn(2 * batch, 256), dim=1).requires_grad_(True)
    t0 = time.perf_counter()
    crit(z[:batch], z[batch:]).backward()
    t_ntx = time.perf_counter() - t0
    return t_enc, t_ntx


def workload_config(args, world):
    """The `config` object both arms print: what the workload IS (BASELINE.json configs[1]; configs[4] for N > 1), not how an arm ran it."""
    return {"workload": WORKLOAD if args.model == "gin" else WORKLOAD.replace("GIN-5", "GCN-5 (un-normalised GCNConv as the reference computes it)"),
            "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "parallelism": f"dp{world}" + ("" if world == 1 else ("-localneg" if args.local_negatives else "-globalneg")),
            "l2": "no flush: per-step working set (~5 GB of activations) >> 126 MB L2"}


def run_reference(args):
    """The reference's own CPU implementation of the path (the oracle port: torch_geometric is not installable), all host cores,
    EXACTLY --steps timed steps after --warmup, each step a bounded sample of the workload: 512 of its 4096 pairs (the reference's
    default batch; its NT-Xent [2N,2N,C] broadcast needs 68.7 GB at 4096 pairs and ~1.5 s per step at 512)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    mols, threads, dt = cpu_reference_run(args.cpu_batch, steps, warmup)
    t_enc, t_ntx = cpu_reference_split(args.cpu_batch)
    sample = (f"{warmup} warm-up + {steps} full steps (zero_grad, 2 encoder passes, normalize, NT-Xent, backward, Adam) of {args.cpu_batch} pairs "
              f"out of the workload's {args.batch} per step, {threads} host threads; oracle port of the PyG path + the reference's NT-Xent formulation")
    line = {"impl": "reference", "metric": METRIC, "value": mols, "unit": "molecules/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": mols, "unit": "molecules/s", "cores": threads, "kind": "port", "sample": sample,
                             "split_s_per_step": {"encoder_fwd_bwd": t_enc, "ntxent_fwd_bwd": t_ntx, "full_step": dt}},
            "e2e": {"value": mols, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- our arm
def batch_bytes(b):
    return sum(t.numel() * t.element_size() for t in (b.x, b.edge_index, b.edge_attr, b.batch))


def run_ours(args):
    import torch.distributed as dist
    from molclr_b200 import Batch, GCN, GINet, NTXentLoss, _lib, ops, pretrain_loss
    from molclr_b200.synth import make_pair_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    if args.no_pdl:
        lib.molclr_set_pdl(0)
    pdl = bool(lib.molclr_set_pdl(-1))
    B = args.batch

    torch.manual_seed(0)
    model = (GINet if args.model == "gin" else GCN)(5, 300, 512, 0, "mean").to(dev)
    if args.precision is None:
        args.precision = model.precision
    model.precision = args.precision
    if world > 1:
        from molclr_b200.dist import DataParallelStep
        kw = {} if args.dp_bucket is None else ({"overlap": False} if args.dp_bucket == 0 else {"bucket_floats": args.dp_bucket})
        stepper = DataParallelStep(model, B, 0.1, True, global_negatives=not args.local_negatives, **kw)
    else:
        stepper = None
    crit = NTXentLoss(dev, B, 0.1, True)
    opt = torch.optim.Adam(model.parameters(), 5e-4, weight_decay=1e-5, fused=True)

    NB = 3   # distinct batches cycled through (per rank: seed + rank)
    host = [tuple(b.pin_memory() for b in make_pair_batch(B, seed=(0 if args.same_batches else 1000 * rank) + i)) for i in range(NB)]
    resident = [tuple(b.to(dev) for b in pair) for pair in host]
    torch.cuda.synchronize()

    def fresh(b):        # a new Batch object -> the CSR plan is rebuilt every step, as in real training
        return Batch(b.x, b.edge_index, b.edge_attr, b.batch, b.num_graphs)

    def step(bi, bj):
        opt.zero_grad(set_to_none=True)
        if stepper is not None:
            loss = stepper.loss(fresh(bi), fresh(bj))
            loss.backward()
            stepper.allreduce_gradients()
        else:
            loss = pretrain_loss(model, crit, fresh(bi), fresh(bj))
            loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------- value: inputs resident in HBM
    for i in range(args.warmup):
        step(*resident[i % NB])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = lib.molclr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(*resident[i % NB])
    e1.record()
    barrier()
    clocks = sampler.finish()
    launches = (lib.molclr_launch_count() - l0) // max(args.steps, 1)
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * B / (ms_step * 1e-3)
    last_loss = float((stepper.global_loss(loss) if stepper is not None else loss).item())

    # ---------------- e2e: pinned host batches, H2D + loss D2H inside the timed region.  As a real input pipeline would, the
    # copy of batch k+1 (side stream) overlaps the step of batch k; every timed step still performs one full H2D of its
    # inputs and one D2H read of its loss.
    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch(pair):
        main = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            dev_pair = tuple(b.to(dev, non_blocking=True) for b in pair)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        for b in dev_pair:
            for t in (b.x, b.edge_index, b.edge_attr, b.batch):
                t.record_stream(main)
        return dev_pair, ev

    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        """Every step: H2D of its inputs (side stream, overlapping the previous step) and a D2H read of its loss into pinned
        memory; the host consumes the loss of step k while step k+1 runs (a logger does not need to stall the GPU), the last one
        after the final synchronize."""
        nxt = prefetch(host[0])
        seen = []
        for i in range(n):
            (bi, bj), ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            nxt = prefetch(host[(i + 1) % NB])
            loss_host[i % 2].copy_(step(bi, bj).detach(), non_blocking=True)
            loss_ready[i % 2].record()
            if i > 0:
                loss_ready[(i - 1) % 2].synchronize()
                seen.append(float(loss_host[(i - 1) % 2]))
        torch.cuda.synchronize()
        seen.append(float(loss_host[(n - 1) % 2]))
        assert len(seen) == n and all(v == v for v in seen)

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    h2d = sum(batch_bytes(b) for b in host[0])

    # ---------------- roofline: GINE aggregation (layers >= 1: fused BN+ReLU gather) and the GIN MLP products, timed IN SITU:
    # CUDA event pairs recorded by the native whole-pass calls around those launches, on the stream they are launched on
    import ctypes as C
    roof, roof_gemm = None, None
    D = 300
    if args.model == "gin":
        lib.molclr_step_timing(1)
        n_timed = 6
        for i in range(n_timed):
            step(*resident[i % NB])
        torch.cuda.synchronize()

        def read(cat):
            ms, n = C.c_double(0.0), C.c_int(0)
            lib.molclr_step_timing_read(cat, C.byref(ms), C.byref(n))
            return ms.value, n.value
        (agg_ms, n_agg), (fwd1_ms, n_fwd1), (bwd_ms, n_bwd), (dw_ms, n_dw), (fwd2_ms, n_fwd2) = read(0), read(1), read(2), read(3), read(4)
        fwd_ms, n_fwd = fwd1_ms + fwd2_ms, n_fwd1 + n_fwd2
        lib.molclr_step_timing(0)
        # algorithmic bytes per launch (DESIGN.md / SURVEY 8d): read the source rows once, write the aggregate once, CSR
        # rowptr / col / packed attributes, BatchNorm coefficients and bond tables
        nodes = [int(b.x.size(0)) for i in range(n_timed) for b in resident[i % NB]]
        edges = [int(b.edge_index.size(1)) for i in range(n_timed) for b in resident[i % NB]]
        alg = lambda N, E: 4 * D * N + 4 * D * N + 4 * (N + 1) + 5 * E + 4 * D * (2 + 8)
        peak, peak_src = peaks()
        if n_agg:
            agg_bytes = sum(alg(N, E) for N, E in zip(nodes, edges)) / len(nodes)
            us = agg_ms * 1e3 / n_agg
            achieved = agg_bytes / (us * 1e-6) / 1e9
            roof = {"bound": "hbm", "kernel": "gine_aggregate_fwd_tile_kernel<1, 0> (BatchNorm+ReLU-fused gather, layers >= 1)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "us_per_launch": us,
                    "algorithmic_bytes_per_launch": agg_bytes, "launches_timed": n_agg}
        try:
            tf32_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]) / 2
            tf32_src = "measured sustained bf16 cuBLAS throughput / 2 (MEASURED_PEAKS.json; TF32 runs at half the bf16 rate)"
        except Exception:
            tf32_peak, tf32_src = 1100.0, "fallback: nominal dense TF32"
        flops_per_product = 2.0 * 300 * 600 * (sum(nodes) / len(nodes))          # one [N, 300] x [300, 600] (or transposed) product
        tfl = lambda ms, n: (flops_per_product * n / (ms * 1e-3) / 1e12) if n and ms > 0 else None
        fwd, bwd, dwt = tfl(fwd_ms, n_fwd), tfl(bwd_ms, n_bwd), tfl(dw_ms, n_dw)
        if fwd:
            roof_gemm = {"bound": "tensor", "kernel": "gemm_tf32_kernel, products of the GIN MLP (M or K = nodes)", "unit": "TFLOP/s", "peak": tf32_peak,
                         "peak_source": tf32_src, "achieved": fwd, "frac": fwd / tf32_peak,
                         "forward_compensated_useful_tflops": fwd, "forward_frac": fwd / tf32_peak, "forward_calls_timed": n_fwd,
                         "backward_single_pass_tflops": bwd, "backward_frac": bwd / tf32_peak if bwd else None, "backward_calls_timed": n_bwd,
                         "weight_gradient_tflops": dwt, "weight_gradient_frac": dwt / tf32_peak if dwt else None, "weight_gradient_calls_timed": n_dw,
                         # the weight-gradient products read 3 [N, 300]-sized operand blocks once and write 0.7 MB: they sit on the HBM roofline
                         "weight_gradient_hbm_gbs": (4.0 * 900 * (sum(nodes) / len(nodes)) * n_dw / (dw_ms * 1e-3) / 1e9) if n_dw and dw_ms > 0 else None,
                         "weight_gradient_hbm_frac": (4.0 * 900 * (sum(nodes) / len(nodes)) * n_dw / (dw_ms * 1e-3) / 1e9 / peak) if n_dw and dw_ms > 0 else None,
                         "us_per_call": {"forward": fwd_ms * 1e3 / max(n_fwd, 1), "forward_a_to_u": fwd1_ms * 1e3 / max(n_fwd1, 1),
                                         "forward_u_to_z": fwd2_ms * 1e3 / max(n_fwd2, 1), "backward": bwd_ms * 1e3 / max(n_bwd, 1), "weight_gradient": dw_ms * 1e3 / max(n_dw, 1)},
                         "note": "useful FLOPs 2MNK per product (the compensated forward issues three products per K slice: fp16x3 = three kind::f16 "
                                 "passes, tf32x3 = one TF32 pass + two bf16 correction passes); per-call times are event pairs around single "
                                 "launches and include their launch latency; see DESIGN.md section 6",
                         "bound_note": "the tensor peak is the nominal ceiling of these products; what the forward products actually sit on is the "
                                       "shared-memory data pipe (ncu --set full, profiles/r2_ncu_fwd2_fp16x3.md: LSU wavefronts 59 % + tensor-core "
                                       "operand wavefronts 29 % of peak, tensor pipe 48 %), the weight gradients on HBM"}
    ncu_traffic = os.path.join(ROOT, "profiles", "aggregate_traffic.json")
    if roof is not None and os.path.exists(ncu_traffic):
        try:      # DRAM bytes per launch of this kernel from an `ncu --set full` capture of the same command (cannot be measured outside a profiler)
            tj = json.load(open(ncu_traffic))
            roof["traffic"] = tj["dram_bytes_per_launch"]
            roof["traffic_source"] = tj.get("source", "profiles/aggregate_traffic.json")
        except Exception:
            pass

    # ---------------- whole-step roofline: useful FLOPs of one step / step time, against the TF32 and BF16 tensor peaks
    n_nodes = sum(int(b.x.size(0)) for b in resident[0])                      # both views
    D, H, Fd = 300, 600, 512
    mlp_flops = (3 * 4.0 * D * H if args.model == "gin" else 3 * 2.0 * D * D) * 5 * n_nodes      # fwd + dX + dW, 5 layers
    head_flops = 3 * 2.0 * (D * Fd + Fd * Fd + Fd * Fd // 2) * 2 * B
    rows_glob = 2 * B * (1 if (world == 1 or args.local_negatives) else world)
    ntx_flops = 4 * 2.0 * (2 * B) * rows_glob * (Fd // 2)                      # S forward, S recomputed, dZ (row and column terms)
    step_flops = mlp_flops + head_flops + ntx_flops
    try:
        bf16_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        bf16_src = "MEASURED_PEAKS.json bf16_tflops_sustained (TF32 = half)"
    except Exception:
        bf16_peak, bf16_src = 2250.0, "fallback: nominal dense bf16 (TF32 = half)"
    step_tfl = step_flops / (ms_step * 1e-3) / 1e12
    roof_step = {"useful_flops_per_step_per_gpu": step_flops, "achieved_tflops_per_gpu": step_tfl, "frac_of_tf32_peak": step_tfl / (bf16_peak / 2),
                 "frac_of_bf16_peak": step_tfl / bf16_peak, "peak_source": bf16_src,
                 "breakdown_gflop": {"encoder_gemms": mlp_flops / 1e9, "head": head_flops / 1e9, "ntxent": ntx_flops / 1e9}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": METRIC if args.model == "gin" else METRIC.replace("GIN", "GCN"), "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": workload_config(args, world),
            "run": {"precision": args.precision, "programmatic_dependent_launch": pdl, "ntxent_operands": "fp16 (kind::f16; same 11-bit significand as tf32), fp32 accumulation",
                    "loss": last_loss, "nodes_per_view": int(resident[0][0].x.size(0)), "edges_per_view": int(resident[0][0].edge_index.size(1))},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "molecules/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms},
            "roofline": roof, "roofline_gemm": roof_gemm, "roofline_step": roof_step}
    if world == 1 and not args.no_extra:
        line["extra"] = extra_configs(args, dev)
    if world == 1 and not args.no_cpu_baseline:
        mols, threads, dt = cpu_reference_run(args.cpu_batch, 3, 1)
        t_enc, t_ntx = cpu_reference_split(args.cpu_batch)
        line["cpu_baseline"] = {"value": mols, "unit": "molecules/s", "cores": threads, "kind": "port",
                                "sample": f"1 warm-up + 3 full steps of {args.cpu_batch} pairs ({dt:.2f} s/step) on {threads} host threads; oracle port of the "
                                          "PyG path + the reference's NT-Xent formulation (infeasible at 4096 pairs: 68.7 GB temporaries)",
                                "split_s_per_step": {"encoder_fwd_bwd": t_enc, "ntxent_fwd_bwd": t_ntx, "full_step": dt}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_configs(args, dev):
    """The other BASELINE.json configurations, measured in the same run with a few steps each (resident inputs, CUDA events):
    config 1 on the GPU (512 pairs: the like-for-like batch of the CPU arm), config 3 (GCN, 4096 pairs), config 4 (fine-tune
    GINet, 1024 graphs, BBBP- and ESOL-shaped, drop_ratio 0 and 0.3), the single-pass `tf32` precision mode of config 2, and
    configs 1 / 2 with plain stream-ordered launches (the A/B of programmatic dependent launch)."""
    from molclr_b200 import Batch, GCN, GINet, NTXentLoss, _lib, ginet_finetune, pretrain_loss
    from molclr_b200.synth import make_pair_batch, make_plain_batch
    lib = _lib.load()

    def timed(step, warm=5, n=10):
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            step()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    fresh = lambda b: Batch(b.x, b.edge_index, b.edge_attr, b.batch, b.num_graphs)

    def pretrain(cls, B, precision, pdl=None):
        if pdl is not None:
            before = lib.molclr_set_pdl(int(pdl))
            try:
                r = pretrain(cls, B, precision)
            finally:
                lib.molclr_set_pdl(before)
            r["programmatic_dependent_launch"] = bool(pdl)
            return r
        torch.manual_seed(0)
        model = cls(5, 300, 512, 0, "mean").to(dev)
        model.precision = precision
        crit = NTXentLoss(dev, B, 0.1, True)
        opt = torch.optim.Adam(model.parameters(), 5e-4, weight_decay=1e-5, fused=True)
        data = [tuple(b.to(dev) for b in make_pair_batch(B, seed=500 + i)) for i in range(2)]
        k = [0]

        def step():
            bi, bj = data[k[0] % 2]
            k[0] += 1
            opt.zero_grad(set_to_none=True)
            pretrain_loss(model, crit, fresh(bi), fresh(bj)).backward()
            opt.step()
        ms = timed(step)
        return {"molecules_per_s": B / (ms * 1e-3), "ms_per_step": ms, "batch": B, "precision": precision}

    def finetune(task, atoms, std, drop):
        torch.manual_seed(0)
        G = 1024
        model = ginet_finetune.GINet(task, 5, 300, 512, drop, "mean").to(dev)
        opt = torch.optim.Adam(model.parameters(), 5e-4, weight_decay=1e-6, fused=True)
        data = [make_plain_batch(G, seed=700 + i, mean_atoms=atoms, std_atoms=std).to(dev) for i in range(2)]
        if task == "classification":
            y, crit = (torch.rand(G, device=dev) < 0.77).long(), torch.nn.CrossEntropyLoss()
        else:
            y, crit = torch.randn(G, 1, device=dev) * 2.1 - 3.05, torch.nn.MSELoss()
        k = [0]

        def step():
            b = data[k[0] % 2]
            k[0] += 1
            opt.zero_grad(set_to_none=True)
            _h, pred = model(fresh(b))
            crit(pred, y).backward()
            opt.step()
        ms = timed(step)
        return {"graphs_per_s": G / (ms * 1e-3), "ms_per_step": ms, "graphs": G, "nodes": int(data[0].x.size(0)), "drop_ratio": drop}

    P = args.precision
    other = "tf32x3" if P == "fp16x3" else "fp16x3"       # the other form of the compensated forward products
    out = {"config1_gpu_512_pairs": pretrain(GINet, 512, P),
           "config3_gcn_4096_pairs": pretrain(GCN, 4096, P),
           "config2_precision_tf32": pretrain(GINet, 4096, "tf32"),
           f"config2_precision_{other}": pretrain(GINet, 4096, other),
           # A/B of programmatic dependent launch: the same steps with plain stream-ordered launches
           "config1_gpu_512_pairs_plain_launches": pretrain(GINet, 512, P, pdl=False),
           "config2_plain_launches": pretrain(GINet, 4096, P, pdl=False),
           "config4_finetune_bbbp_cls_drop0": finetune("classification", 46.0, 18.0, 0.0),
           "config4_finetune_bbbp_cls_drop0.3": finetune("classification", 46.0, 18.0, 0.3),
           "config4_finetune_esol_reg_drop0": finetune("regression", 26.0, 13.0, 0.0),
           "config4_finetune_esol_reg_drop0.3": finetune("regression", 26.0, 13.0, 0.3)}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
