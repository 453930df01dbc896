"""Training path shared by the ResNet and FPN modules (BASELINE.json config 4: frozen/eval BatchNorm,
frozen stem + stage 1, hand-written dgrad / wgrad implicit-GEMM kernels, bucketed gradient all-reduce).

The reference trains through torch autograd over nn.Conv2d / nn.BatchNorm2d (resnet.py:97-119,
fpn.py:88-125); here each module owns ONE autograd.Function whose backward runs a pre-compiled plan
of ``tdet_op`` descriptors on the caller's stream:

    dgrad  = TDET_OP_CONV over the output gradient with the 180-degree-rotated, BN-scale-folded,
             transposed weights (tdet_pack_dgrad_weight); the ReLU backward of the layer it feeds is
             the conv's ``mask`` epilogue, residual / shortcut / external gradients are its
             ``residual`` / ``coarse`` (parity scatter) operands, so every gradient tensor is
             written exactly once.  A stride-2 3x3 conv's dgrad is four parity-class convs over the
             coarse gradient + TDET_OP_PARITY_MERGE (no zero insertion).
    wgrad  = TDET_OP_WGRAD (pixels are the contraction dimension) into fp32 accumulators that live
             in one flat bucket per stage -- the all-reduce unit.

Gradient tensors are plain bf16 (tcgen05 needs both MMA operands in one format, and the stage
outputs that cross module boundaries are bf16), parameter gradients fp32.
"""
import torch

from . import engine


class BucketAllReduce(object):
    """Sums (averages) flat gradient buckets across ranks, overlapped with the rest of backward.

    ``reduce(bucket, params)`` is called by a module's backward as soon as the kernels writing that bucket
    have been enqueued.  On a side stream that waits for exactly that point of the compute stream
    the bucket is copied out (the plan re-zeroes its accumulators in the next step) and the copy is
    all-reduced (``ReduceOp.AVG``: no separate division pass), so NCCL traffic over NVLink overlaps the
    remaining dgrad/wgrad kernels; the returned tensor is what the module hands to autograd.
    ``finish()`` makes the compute stream wait for every outstanding bucket: modules call it at the end of
    their backward unless ``defer=True``, in which case the training loop calls it once after ``backward()``
    and BEFORE anything reads ``p.grad`` (then even the neck's bucket overlaps the whole backbone backward).

    Deferred hand-over is only sound while autograd merely *adopts* the returned views (``p.grad is None``:
    AccumulateGrad stores the tensor without launching a kernel).  When a parameter of the bucket already
    holds a gradient (gradient accumulation, ``zero_grad(set_to_none=False)``) autograd would run
    ``p.grad += g`` on the compute stream, racing the side stream -- for such buckets the compute stream
    waits for the collective before the views are handed over (correct, that bucket just does not overlap).
    The world size is queried when first needed, so the object may be built before ``init_process_group``.
    With CPU tensors (gloo; host-logic tests) or a single rank the collective runs inline.
    """

    def __init__(self, group=None, average=True, defer=False, sm_reserve=0):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.average = average
        self.defer = defer
        # SMs the modules' persistent GEMM grids leave to the NCCL kernels (pair with NCCL_MAX_CTAS)
        self.sm_reserve = sm_reserve
        self.enabled = True   # False: copy the buckets out but skip the collective (bench A/B of its cost)
        self._streams = {}
        self._pending = []
        self._copied = {}  # bucket address -> event: its side-stream copy has completed
        self._timing = None  # list of (start, end) events around each collective while profiling
        self.buckets_reduced = 0
        self.bytes_reduced = 0
        self.buckets_not_overlapped = 0

    @property
    def world(self):
        return self.dist.get_world_size(self.group) if self.dist.is_initialized() else 1

    def attach(self, module, device):
        """Called by module.set_grad_sync: re-size the module's grids if SMs are reserved for NCCL."""
        if self.sm_reserve and self.world > 1 and device.type == "cuda":
            engine.set_sm_reserve(device, self.sm_reserve)
            module._plans = engine.PlanCache()

    def _comm_stream(self, device):
        s = self._streams.get(device)
        if s is None:
            s = torch.cuda.Stream(device=device)
            self._streams[device] = s
        return s

    def guard(self, bucket):
        """Before a plan overwrites `bucket`: wait until the previous step's copy of it is done."""
        ev = self._copied.pop(bucket.data_ptr(), None)
        if ev is not None:
            torch.cuda.current_stream(bucket.device).wait_event(ev)

    def profile(self, on=True):
        """Record CUDA events around every collective (side stream) until profile(False)."""
        self._timing = [] if on else None

    def collective_ms(self):
        """Sum of the device time of the collectives recorded since profile(True) (synchronises)."""
        if not self._timing:
            return 0.0
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self._timing)

    def reduce(self, bucket, params=()):
        self.buckets_reduced += 1
        self.bytes_reduced += bucket.numel() * bucket.element_size()
        world = self.world
        if not bucket.is_cuda:
            out = bucket.clone()
            if world > 1 and self.enabled:
                self.dist.all_reduce(out, group=self.group)
                if self.average:
                    out.div_(world)
            return out
        main = torch.cuda.current_stream(bucket.device)
        comm = self._comm_stream(bucket.device)
        # The copy comes from the MAIN stream's pool and is NOT record_stream()ed on the side stream: the object keeps
        # a reference until finish() has made the main stream wait for the collective, so it cannot be freed (and
        # handed to another main-stream allocation) while the side stream still works on it.  With record_stream a
        # freed copy only became reusable once an event had completed, the pool kept growing by a segment every few
        # steps, and each of those cudaMallocs can stall the host for 40-100 ms (measured: one 9.3 ms step in ten
        # taking 50-100 ms at random).
        out = torch.empty_like(bucket)
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            out.copy_(bucket)
            ev = torch.cuda.Event()
            ev.record(comm)
            if world > 1 and self.enabled:
                t0 = t1 = None
                if self._timing is not None:
                    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    t0.record(comm)
                if self.average and AVG_IN_NCCL:
                    self.dist.all_reduce(out, op=self.dist.ReduceOp.AVG, group=self.group)
                else:
                    self.dist.all_reduce(out, group=self.group)
                    if self.average:
                        out.div_(world)
                if t1 is not None:
                    t1.record(comm)
                    self._timing.append((t0, t1))
        self._copied[bucket.data_ptr()] = ev
        if any(p.grad is not None for p in params):
            # autograd will accumulate into existing gradients on the compute stream: order it after the collective
            main.wait_stream(comm)
            self.buckets_not_overlapped += 1
        else:
            self._pending.append((bucket.device, comm, out))
        return out

    def finish(self):
        for device, comm, _out in self._pending:
            torch.cuda.current_stream(device).wait_stream(comm)
        self._pending = []   # (drops the references that kept the copies alive while the side stream used them)

    def module_done(self):
        if not self.defer:
            self.finish()


# ncclAvg (1) or ncclSum + a division pass on the side stream (0)
AVG_IN_NCCL = __import__("os").environ.get("TDET_NCCL_AVG", "1") != "0"


class GradBucket(object):
    """Flat fp32 buffer holding the OIHW gradients of a list of parameters (one all-reduce unit)."""

    def __init__(self, params, device):
        self.params = list(params)
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 63) // 64 * 64  # keep every view 256-byte aligned (vector reductions)
        self.flat = torch.zeros(max(off, 64), dtype=torch.float32, device=device)

    def view(self, i):
        p = self.params[i]
        return self.flat[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape)

    def index_of(self, param):
        for i, p in enumerate(self.params):
            if p is param:
                return i
        raise KeyError("parameter not in bucket")


class BackwardBuilder(object):
    """Emits the backward ops of conv layers into one op list, with a liveness-based pool for the
    gradient tensors and one zero-initialised workspace for the k x k wgrad accumulators."""

    def __init__(self, device, cache, dtype=torch.bfloat16, meta=None):
        self.device = device
        self.cache = cache     # operand cache of the owning module (rotated weights live there)
        self.dtype = dtype     # storage format of the gradient tensors (= format of the saved activations)
        self.meta = meta       # engine.MetaArena: gradient tensors carry a per-tensor power-of-two exponent
        self.scaled = meta is not None
        self.ops = []
        self.free = {}
        self.buffers = []
        self.acc_jobs = []     # (conv name, cout, cin, kh, kw, destination view) of k x k convs
        self.acc_ws = None

    # ---- gradient tensors ---------------------------------------------------------------------
    def new_act(self, shape, meta="new"):
        numel = 1
        for s in shape:
            numel *= s
        bucket = self.free.get(numel)
        if bucket:
            buf = bucket.pop()
        else:
            buf = torch.empty(numel, dtype=torch.bfloat16, device=self.device)  # raw 16-bit storage
            self.buffers.append(buf)
        if meta == "new":
            meta = self.meta.new() if self.scaled else None
        return engine.Act(buf, shape, self.dtype, meta)

    def release(self, act):
        if act is not None:
            self.free.setdefault(act.buf.numel(), []).append(act.buf)

    # ---- ops ------------------------------------------------------------------------------------
    def dgrad_weight(self, name, module, scale, deps=()):
        """`scale` is the folded-BN scale tensor of the conv (refreshed in place before this entry: the
        cache keeps insertion order); `deps` = the BatchNorm parameters / buffers it derives from."""
        if module.groups > 1:
            # grouped conv (ResNeXt): the data-gradient operand of its dense block-diagonal form, which is again
            # block diagonal -- the dgrad conv runs as a grouped conv of the same cardinality
            return self.cache.get((name, "wd", self.dtype),
                                  lambda out: engine.pack_dgrad_weight(
                                      engine.dense_group_weight(module.weight, module.groups), scale, dtype=self.dtype,
                                      out=out),
                                  deps=(module.weight,) + tuple(deps))
        return self.cache.get((name, "wd", self.dtype),
                              lambda out: engine.pack_dgrad_weight(module.weight, scale, dtype=self.dtype, out=out),
                              deps=(module.weight,) + tuple(deps))

    def dgrad_consts(self, name, module, wd, deps=()):
        """{G, 0} with |dgrad(g)| <= G * max|g|: the bound the kernel picks the output exponent from."""
        if not self.scaled:
            return None
        return self.cache.get((name, "wdc", self.dtype), lambda out: engine.bound_consts(wd, None, None, out=out),
                              deps=(module.weight,) + tuple(deps))

    def conv_dgrad_op(self, name, module, wd, src, dx, k, pad, dil, deps=(), **kw):
        self.ops.append(engine.op_conv(src, wd, dx, k, k, 1, pad, dil, consts=self.dgrad_consts(name, module, wd, deps),
                                       scaled_out=self.scaled, groups=module.groups if GROUPED_DGRAD else 1, **kw))

    def dgrad(self, name, module, scale, g, in_shape, residual=None, coarse=None, mask=None, deps=()):
        """Gradient w.r.t. the input (shape `in_shape`, NHWC) of conv `module` given g = dL/d(BN(conv))."""
        k = module.kernel_size[0]
        stride, pad, dil = module.stride[0], module.padding[0], module.dilation[0]
        wd = self.dgrad_weight(name, module, scale, deps)
        if (stride == 2 and k == 3 and pad == 1 and dil == 1 and residual is None and coarse is None
                and S2_DGRAD_PARITY):
            return self._dgrad_s2_parity(name, module, wd, g, in_shape, mask, deps)
        dx = self.new_act(in_shape)
        src = g
        tmp = None
        if stride == 2:
            # adjoint of the output subsampling: zero-insertion upsample, then the stride-1 dgrad
            # (zero insertion keeps the exponent and the |max|: the metadata is shared)
            tmp = self.new_act((in_shape[0], in_shape[1] + 2 * pad - dil * (k - 1),
                                in_shape[2] + 2 * pad - dil * (k - 1), g.shape[3]), meta=g.meta)
            self.ops.append(engine.op_dilate2(g, tmp))
            src = tmp
        elif stride != 1:
            raise NotImplementedError("dgrad for conv stride %d" % stride)
        self.conv_dgrad_op(name, module, wd, src, dx, k, dil * (k - 1) - pad, dil, deps, residual=residual,
                           coarse=coarse, coarse_parity=coarse is not None, mask=mask)
        self.release(tmp)
        return dx

    def _dgrad_s2_parity(self, name, module, wd, g, in_shape, mask, deps):
        """3x3 / stride 2 / pad 1 dgrad without zero insertion: dx[2i+a][2j+b] only sees the taps of the rotated
        kernel whose parity matches (a, b), so it is four small stride-1 convs over the coarse gradient (1x1, 1x2,
        2x1, 2x2: a quarter of the MMAs of the 3x3 conv over the zero-inserted gradient) and one interleaving pass
        that also applies the ReLU-backward mask (TDET_OP_PARITY_MERGE)."""
        n, hc, wc, _ = g.shape
        cin = in_shape[3]
        parts = []
        for a in (0, 1):
            for b in (0, 1):
                rs = [1] if a == 0 else [0, 2]
                ss = [1] if b == 0 else [0, 2]
                kh, kw = len(rs), len(ss)
                pad = 1 if (a or b) else 0

                def make(out, rs=rs, ss=ss):
                    sub = wd[:, rs][:, :, ss]
                    if out is None:
                        return sub.contiguous()
                    out.copy_(sub)
                    return out

                w_ab = self.cache.get((name, "wd%d%d" % (a, b), self.dtype), make, deps=(module.weight,) + tuple(deps))
                consts = None
                if self.scaled:
                    consts = self.cache.get((name, "wdc%d%d" % (a, b), self.dtype),
                                            lambda out, w_ab=w_ab: engine.bound_consts(w_ab, None, None, out=out),
                                            deps=(module.weight,) + tuple(deps))
                part = self.new_act((n, hc + 2 * pad - kh + 1, wc + 2 * pad - kw + 1, cin))
                self.ops.append(engine.op_conv(g, w_ab, part, kh, kw, 1, pad, 1, consts=consts, scaled_out=self.scaled,
                                               groups=module.groups if GROUPED_DGRAD else 1))
                parts.append(part)
        dx = self.new_act(in_shape)
        self.ops.append(engine.op_parity_merge(parts, dx, hc, wc, mask=mask, scaled_out=self.scaled))
        for part in parts:
            self.release(part)
        return dx

    def wgrad(self, name, module, scale, x, g, dst):
        """dst: fp32 OIHW view inside a GradBucket."""
        k = module.kernel_size[0]
        cout, cin = module.out_channels, module.in_channels
        stride, pad, dil = module.stride[0], module.padding[0], module.dilation[0]
        if module.groups > 1 and k == 1:
            raise NotImplementedError("wgrad of a grouped 1x1 conv")
        if k == 1:
            self.ops.append(engine.op_wgrad(x, g, dst, 1, 1, stride, pad, dil, scale=scale))
        else:
            # (a grouped conv accumulates its DENSE weight gradient; the unpack keeps the block diagonal)
            self.acc_jobs.append((len(self.ops), name, cout, cin, k, stride, pad, dil, scale, x, g, dst, module.groups))
            self.ops.append(None)  # placeholder: resolved once the accumulator workspace exists
            self.ops.append(None)

    def finalize(self):
        """Allocates the accumulator workspace, resolves the k x k wgrad ops and prepends its reset."""
        total = sum(c[2] * c[3] * c[4] * c[4] for c in self.acc_jobs)
        head = []
        if total:
            self.acc_ws = torch.zeros(total, dtype=torch.float32, device=self.device)
            head.append(engine.op_zero(self.acc_ws))
            off = 0
            for (idx, name, cout, cin, k, stride, pad, dil, scale, x, g, dst, groups) in self.acc_jobs:
                n = cout * cin * k * k
                acc = self.acc_ws[off:off + n]
                off += n
                self.ops[idx] = engine.op_wgrad(x, g, acc, k, k, stride, pad, dil, scale=scale)
                self.ops[idx + 1] = engine.op_dw_unpack(acc, dst, cout, cin, k, k, groups=groups)
        return head + self.ops, len(head)


# dgrad of a grouped conv as a grouped conv (64-channel band kernel); 0 = dense over the block-diagonal operand
GROUPED_DGRAD = __import__("os").environ.get("TDET_GROUPED_DGRAD", "1") != "0"

# stride-2 3x3 dgrad as four parity-class convs (default) instead of a 3x3 conv over the zero-inserted gradient
S2_DGRAD_PARITY = __import__("os").environ.get("TDET_S2_DGRAD", "parity").lower() != "dilate"


def as_grad_nhwc(g, like):
    """Incoming autograd gradient -> dense NHWC bf16 (zero-copy when it already is)."""
    if g is None:
        return torch.zeros_like(like, dtype=torch.bfloat16, memory_format=torch.channels_last)
    if g.dtype == torch.bfloat16 and g.is_contiguous(memory_format=torch.channels_last):
        return g
    return g.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


class PlanFunction(torch.autograd.Function):
    """autograd bridge: forward/backward are the owning module's compiled plans.

    apply(module, n_inputs, *inputs_then_params) -> tuple of outputs."""

    @staticmethod
    def forward(ctx, module, n_inputs, *args):
        inputs, params = args[:n_inputs], args[n_inputs:]
        outs, state = module._train_forward(inputs, params)
        ctx.module = module
        ctx.state = state
        ctx.n_inputs = n_inputs
        ctx.n_params = len(params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        g_inputs, g_params = ctx.module._train_backward(ctx.state, gouts)
        ctx.state = None
        return (None, None) + tuple(g_inputs) + tuple(g_params)
