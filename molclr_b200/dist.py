"""Data-parallel MolCLR pre-training over the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Molecules are independent
through the whole encoder, so every rank runs the encoder kernels on its own batch of B pairs with its own
BatchNorm statistics; the path has exactly two exchange steps:

1. **NT-Xent with global negatives** -- the (cosine-normalised) projections of all ranks are all-gathered
   (ONE collective, fp16 rows when the fp16 kernels apply) into the candidate matrix ``cols [W][2B][C]``
   (rank-major blocks ``[zjs_r ; zis_r]``: the loss does not depend on the candidate order, and a row's positive
   is found in its own rank's other block), each rank computes the log-sum-exp of ITS
   2B anchor rows against all 2*W*B candidates, the 2B row log-sum-exps are all-gathered (ONE collective), and the backward
   uses the symmetry of the loss (``molclr_ntxent_bwd``: the column-softmax term of a local row is
   recomputed from the same similarity tile with the other row's gathered log-sum-exp), so the gradient of
   the GLOBAL mean loss w.r.t. the local projections needs no reduce-scatter of a [2*W*B, C] gradient.
2. **Gradient all-reduce** -- one flat fp32 buffer (2.4 M elements for GIN-5/300/512) in buckets by backward-pass
   readiness, each bucket's all-reduce launched from a gradient hook so that it overlaps the rest of the backward.

Loss / gradient scale: the objective is the mean over all 2*W*B anchors.  With global negatives every rank
holds the exact partial derivative of that objective through its own projections, so parameter gradients
are SUMMED over ranks; with local negatives each rank's loss is the mean over its own 2B anchors and the
gradients are AVERAGED (the usual DDP convention).  ``loss()`` returns this rank's share; the global value
is ``global_loss(share)`` (one scalar all-reduce, for logging only).

The collective logic here is device-agnostic: ``kern`` supplies the five kernels it calls.  The product
default is the CUDA kernels (``molclr_b200.ops``); the CPU tests (gloo, world_size 2) inject a plain-torch
stand-in so that the gather ordering, row offsets and gradient scaling are checked without a GPU.
"""
import torch
import torch.distributed as dist


class CudaKernels:
    """The sm_100a kernels behind the C ABI (no fallback: raises without a CUDA device / the library)."""

    def __init__(self):
        from . import functional, ops
        self._ops = ops
        self.normalize = functional.normalize
        self.l2_normalize_bwd = ops.l2_normalize_bwd

    def rows_fwd(self, zjs, zis, eps, normalise, inv_temperature):
        """(y, operand, inv): the local rows [zjs; zis], optionally unit-normalised; `operand` is what gets all-gathered and fed to
        the tensor cores -- fp16 rows when the fp16 path applies (cosine similarity, C <= 256, moderate 1/tau), else tf32-rounded fp32."""
        ops = self._ops
        if normalise and ops.ntxent_h_supported(zjs.shape[1], inv_temperature):
            y, _, inv, y16 = ops.ntxent_rows_fwd(zjs.contiguous(), zis.contiguous(), eps, True, want_r=False, want16=True)
            return y, y16, inv
        y, y_r, inv = ops.ntxent_rows_fwd(zjs.contiguous(), zis.contiguous(), eps, normalise)
        return y, y_r, inv

    def ntxent_fwd(self, rep, cols, C, row_offset, inv_temperature, row_offset2, unit_rows):
        if cols.dtype == torch.float16:
            return self._ops.ntxent_fwd_h(rep, cols, C, row_offset, row_offset2, inv_temperature)
        return self._ops.ntxent_fwd(rep, cols, row_offset, inv_temperature, row_offset2, unit_rows=unit_rows)

    def ntxent_bwd(self, rep, cols, C, row_offset, inv_temperature, row_lse, col_lse, row_offset2, unit_rows):
        if cols.dtype == torch.float16:
            return self._ops.ntxent_bwd_h(rep, cols, C, row_offset, row_offset2, inv_temperature, row_lse, col_lse)
        return self._ops.ntxent_bwd(rep, cols, row_offset, inv_temperature, row_lse, col_lse, row_offset2, unit_rows=unit_rows)


class _GlobalNTXentFunction(torch.autograd.Function):
    """NT-Xent (nt_xent.py:47-65) of this rank's 2B anchors against the candidates of all ranks.

    Exchange steps: ONE all-gather of the local operand rows [zjs_r; zis_r] ([2B, C], fp16 when the fp16 kernels apply: half the
    bytes) into the candidate matrix [W][2B][C] -- rank-major, so this rank's rows are a VIEW of it -- and ONE all-gather of the
    2B row log-sum-exps.  The kernels find a row's positive at the same position of its rank's other block
    (row_offset = r 2B, row_offset2 = r 2B + B); the loss does not depend on how the candidates are ordered."""

    @staticmethod
    def forward(ctx, zis, zjs, temperature, use_cosine, group, kern):
        W, r = dist.get_world_size(group), dist.get_rank(group)
        B, C = zis.shape
        # rows [zjs; zis] (nt_xent.py:48), CosineSimilarity's normalisation (eps 1e-8, nt_xent.py:19,44), operand copy: one kernel
        local_n, local_op, inv = kern.rows_fwd(zjs, zis, 1e-8, use_cosine, 1.0 / temperature)
        cols = torch.empty(W * 2 * B, local_op.shape[1], dtype=local_op.dtype, device=local_op.device)
        dist.all_gather_into_tensor(cols, local_op, group=group)
        rep = cols[r * 2 * B:(r + 1) * 2 * B]
        share, lse, _pos = kern.ntxent_fwd(rep, cols, C, r * 2 * B, 1.0 / temperature, r * 2 * B + B, use_cosine)
        col_lse = torch.empty(W * 2 * B, dtype=lse.dtype, device=lse.device)
        dist.all_gather_into_tensor(col_lse, lse, group=group)
        ctx.save_for_backward(local_n, inv, cols, lse, col_lse)
        ctx.meta = (B, C, r, temperature, use_cosine, kern)
        return share[0]

    @staticmethod
    def backward(ctx, g_loss):
        local_n, inv, cols, lse, col_lse = ctx.saved_tensors
        B, C, r, temperature, use_cosine, kern = ctx.meta
        rep = cols[r * 2 * B:(r + 1) * 2 * B]
        g = kern.ntxent_bwd(rep, cols, C, r * 2 * B, 1.0 / temperature, lse, col_lse, r * 2 * B + B, use_cosine)
        if use_cosine:
            g = kern.l2_normalize_bwd(g, local_n, inv, 1e-8, gscale=g_loss.contiguous())
        else:
            g = g * g_loss
        return g[B:], g[:B], None, None, None, None


def global_ntxent(zis, zjs, temperature, use_cosine_similarity=True, group=None, kern=None):
    """This rank's share of the NT-Xent loss over the global batch; the sum over ranks is the loss of
    ``NTXentLoss(batch_size=W*B)`` applied to the rank-concatenated projections."""
    if zis.shape != zjs.shape or zis.dim() != 2:
        raise RuntimeError(f"global_ntxent: zis {tuple(zis.shape)} and zjs {tuple(zjs.shape)} must be equal 2-D shapes")
    return _GlobalNTXentFunction.apply(zis, zjs, float(temperature), bool(use_cosine_similarity), group, kern or CudaKernels())


def _bucket_of(name, num_layer):
    """Readiness order of the gradients in the backward pass: head first, then encoder layers from the last to the first, the
    node-embedding tables with layer 0."""
    parts = name.split(".")
    if parts[0] in ("gnns", "batch_norms") and len(parts) > 1 and parts[1].isdigit():
        return 1 + (num_layer - 1 - int(parts[1]))
    if parts[0].startswith("x_embedding"):
        return num_layer
    return 0


class _GradSink:
    """Receives finished slices of the flat gradient buffer from inside ``GINet``'s pair backward and all-reduces them.  The slices
    arrive from the END of the buffer towards its start (head, layer L-1, ..., layer 0, embeddings) and are contiguous, so they are
    coalesced: an all-reduce is launched once at least ``bucket_floats`` elements are pending (and for the remainder when the
    backward pass is done) -- few enough launches not to crowd the GEMMs off the SMs, early enough to hide under the backward."""

    def __init__(self, stepper):
        self.stepper = stepper
        self.reset()

    def reset(self):
        self.works, self.flat, self.views, self.covered, self.pending = [], None, None, 0, None

    def _launch(self, flat):
        lo, hi = self.pending
        self.works.append(dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.stepper.group, async_op=True))
        self.covered += hi - lo
        self.pending = None

    def slice_ready(self, flat, lo, hi, name):
        if not self.stepper.overlap:
            return
        if self.pending is None:
            self.pending = (lo, hi)
        elif hi == self.pending[0]:
            self.pending = (lo, self.pending[1])
        elif lo == self.pending[1]:
            self.pending = (self.pending[0], hi)
        else:                                                 # not adjacent: flush what is pending, start a new range
            self._launch(flat)
            self.pending = (lo, hi)
        if self.pending[1] - self.pending[0] >= self.stepper.bucket_floats:
            self._launch(flat)

    def backward_done(self, flat, views):
        self.flat, self.views = flat, views
        if self.pending is not None:
            self._launch(flat)

    def finish(self, scale):
        if self.covered == 0:                                 # overlap off: one all-reduce of the whole buffer now
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.stepper.group)
        elif self.covered != self.flat.numel():
            raise RuntimeError(f"DataParallelStep: the backward pass announced {self.covered} of {self.flat.numel()} gradient elements")
        for w in self.works:
            w.wait()
        if scale != 1.0:
            self.flat.mul_(scale)
        for p, v in zip(self.stepper.model._params(), self.views):
            p.grad = v
        self.reset()


class DataParallelStep:
    """The loop body of molclr.py:109-127 for one rank of a data-parallel job.

        stepper = DataParallelStep(model, batch_size, temperature, use_cosine_similarity)
        optimizer.zero_grad(); loss = stepper.loss(xis, xjs); loss.backward()
        stepper.allreduce_gradients(); optimizer.step()

    Construction broadcasts rank 0's parameters and buffers so that all replicas start identical.

    Gradient exchange: the parameters are grouped into buckets by the order in which the backward pass finishes them (head,
    layer L-1, ..., layer 0 + embeddings); each bucket is one contiguous slice of a flat fp32 buffer.  With ``overlap`` (default)
    a post-accumulate hook counts the finished gradients of a bucket and, when the last one lands, copies them into the slice
    (one fused copy) and launches that slice's all-reduce asynchronously -- it runs on NCCL's stream underneath the rest of the
    backward pass.  ``allreduce_gradients()`` then only waits (and handles buckets that did not complete).
    """

    def __init__(self, model, batch_size, temperature, use_cosine_similarity, global_negatives=True, group=None, kern=None,
                 local_criterion=None, overlap=True, bucket_floats=1 << 20):
        if not dist.is_initialized():
            raise RuntimeError("DataParallelStep: torch.distributed is not initialised")
        self.model, self.batch_size, self.temperature = model, batch_size, float(temperature)
        self.use_cosine, self.global_negatives, self.group = bool(use_cosine_similarity), bool(global_negatives), group
        self.kern = kern or CudaKernels()
        self.world = dist.get_world_size(group)
        self._criterion = local_criterion
        if not self.global_negatives and self._criterion is None:
            from .nt_xent import NTXentLoss
            self._criterion = NTXentLoss(None, batch_size, temperature, use_cosine_similarity)
        with torch.no_grad():
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t, src=0, group=group)
        # ---- buckets over a flat buffer
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        L = getattr(model, "num_layer", 0)
        order = sorted(range(len(named)), key=lambda i: (_bucket_of(named[i][0], L), i))
        self._params = [named[i][1] for i in order]
        bucket_ids = [_bucket_of(named[i][0], L) for i in order]
        total = sum(p.numel() for p in self._params)
        pair = hasattr(model, "forward_pair")            # (pair models bring their own flat buffer: see _GradSink)
        self._flat = torch.zeros(0 if pair else total, dtype=self._params[0].dtype, device=self._params[0].device)
        self._views, self._buckets, off = [], [], 0          # bucket: dict(lo, hi, idx=[param indices])
        for i, (p, b) in enumerate(zip(self._params, bucket_ids)):
            self._views.append(None if pair else self._flat[off:off + p.numel()].view_as(p))
            if not self._buckets or self._buckets[-1]["id"] != b:
                self._buckets.append({"id": b, "lo": off, "hi": off, "idx": []})
            self._buckets[-1]["idx"].append(i)
            off += p.numel()
            self._buckets[-1]["hi"] = off
        self._bucket_of_param = {}
        for bi, bk in enumerate(self._buckets):
            for i in bk["idx"]:
                self._bucket_of_param[i] = bi
        self._ready = [0] * len(self._buckets)
        self._work = [None] * len(self._buckets)
        self.overlap = bool(overlap)
        self.bucket_floats = int(bucket_floats)
        # Models with ``forward_pair`` (GINet) run both views in one autograd node and hand finished gradient SLICES (head, layer
        # L-1, ..., layer 0, embeddings) to a sink from inside the backward pass: the all-reduce of a slice overlaps the rest of
        # the second view's backward.  Other models: post-accumulate hooks per parameter bucket.
        self._pair = hasattr(model, "forward_pair")
        self._sink = _GradSink(self) if self._pair else None
        if self.overlap and not self._pair:
            for i, p in enumerate(self._params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    # gradients are summed (global negatives) or averaged (local negatives) over ranks -- see the module docstring
    @property
    def grad_scale(self):
        return 1.0 if self.global_negatives else 1.0 / self.world

    def _make_hook(self, i):
        def hook(_param):
            bi = self._bucket_of_param[i]
            self._ready[bi] += 1
            if self._ready[bi] == len(self._buckets[bi]["idx"]):
                self._flush(bi, async_op=True)
        return hook

    def _flush(self, bi, async_op):
        bk = self._buckets[bi]
        src, dst = [], []
        for i in bk["idx"]:
            p, v = self._params[i], self._views[i]
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad); dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for i in bk["idx"]:
            self._params[i].grad = self._views[i]
        self._work[bi] = dist.all_reduce(self._flat[bk["lo"]:bk["hi"]], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op) or True

    def loss(self, xis, xjs):
        self._ready = [0] * len(self._buckets)
        self._work = [None] * len(self._buckets)
        if self._pair:
            self._sink.reset()
            (_ris, zis), (_rjs, zjs) = self.model.forward_pair(xis, xjs, grad_sink=self._sink)      # molclr.py:57,60
        else:
            _ris, zis = self.model(xis)                     # molclr.py:57
            _rjs, zjs = self.model(xjs)                     # molclr.py:60
        zis = self.kern.normalize(zis, dim=1)               # molclr.py:63-64
        zjs = self.kern.normalize(zjs, dim=1)
        if zis.shape[0] != self.batch_size:
            raise RuntimeError(f"DataParallelStep: got {zis.shape[0]} graphs but batch_size={self.batch_size} "
                               "(the reference requires drop_last=True, dataset.py:180)")
        if self.global_negatives:
            return global_ntxent(zis, zjs, self.temperature, self.use_cosine, self.group, self.kern)
        return self._criterion(zis, zjs)

    def global_loss(self, share):
        """The logged loss value: mean over all anchors of all ranks (one scalar all-reduce)."""
        t = share.detach().clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t * self.grad_scale

    def allreduce_gradients(self):
        """Completes the gradient exchange: buckets whose all-reduce was launched from the backward hooks are waited for, the others
        are reduced now; afterwards ``p.grad`` of every parameter aliases its slice of the flat buffer."""
        if self._pair:
            if self._sink.flat is None:
                raise RuntimeError("DataParallelStep.allreduce_gradients: no gradients to reduce -- call loss(xis, xjs).backward() first")
            self._sink.finish(self.grad_scale)
            return
        for bi in range(len(self._buckets)):
            if self._work[bi] is None:
                self._flush(bi, async_op=False)
        for w in self._work:
            if w is not None and w is not True:
                w.wait()
        if self.grad_scale != 1.0:
            self._flat.mul_(self.grad_scale)
        self._ready = [0] * len(self._buckets)
        self._work = [None] * len(self._buckets)
