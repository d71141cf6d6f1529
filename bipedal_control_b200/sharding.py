"""Multi-GPU layout of a batch of MPC instances: contiguous shards, one all-gather of the solved policies per tick.

Instances are independent (own x0, reference, mode schedule, warm start: SURVEY.md section 8e), so the data path needs no
collective; the only exchange is the north star's all-gather of the feedback policies after each tick.
"""
from __future__ import annotations


def shard_range(batch_total: int, rank: int, world: int):
    """Contiguous block of instances owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(batch_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_policies(local: dict, batch_total: int, world: int):
    """All-gathers every policy tensor (first dim = local instances) into [batch_total, ...] tensors on each rank."""
    import torch
    import torch.distributed as dist
    out = {}
    for name, t in local.items():
        sizes = [shard_range(batch_total, r, world) for r in range(world)]
        if all(b - a == sizes[0][1] - sizes[0][0] for a, b in sizes):
            full = torch.empty((batch_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(full, t.contiguous())
        else:
            parts = [torch.empty((b - a,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for a, b in sizes]
            dist.all_gather(parts, t.contiguous())
            full = torch.cat(parts, dim=0)
        out[name] = full
    return out


class PolicyExchange:
    """The north star's "one all-gather of the solved feedback policies per MPC tick" for one shard (one process per GPU).

    Every policy buffer of the library is ONE contiguous slab [K | uff | x | u | times | events | n_nodes] (bmpc_device_view.slab), so the
    exchange is a single collective per tick.  It runs on a side stream and overlaps with the next tick: the library double-buffers its
    policies, so the slab being gathered is only overwritten two ticks later, and that tick first waits for the gather (`before_tick`).
    `window=True` gathers only the nodes consumers read before the next tick (evaluatePolicy is called for t in [t0, t0 + 1/50 s],
    bipedal_controllers/src/BipedalController.cpp:200): the first `window_nodes` nodes of every instance.

    `slab_provider()` returns the newest policy slab as a 1-D float64 tensor (GPU: an alias of library memory, ordered after the tick in flight
    on `compute_stream`); tests on CPU pass plain tensors and the gloo backend through the same code.
    """

    def __init__(self, mpc, dist, rank, world, window=False, window_nodes=4, slab_provider=None, compute_stream=None, impl="native", max_ctas=0, copy_engines=0):
        """impl "native": the library's own exchange (bmpc_exchange_*: ncclAllGather issued by libbmpc on its exchange stream, optional SM cap /
        copy-engine mode); "torch": torch.distributed all_gather_into_tensor on a side stream (own NCCL group when max_ctas > 0).  The consumed-window
        variant and the CPU tests always use the torch path; a native initialisation failure falls back to it and is reported by describe()."""
        import torch
        self.torch, self.dist, self.mpc, self.rank, self.world = torch, dist, mpc, rank, world
        self.window, self.window_nodes = window, window_nodes
        self.impl, self.max_ctas, self.copy_engines, self.note, self.group = impl, max_ctas, copy_engines, "", None
        if slab_provider is not None or window:
            self.impl = "torch"
        if self.impl == "native":
            try:
                mpc.exchangeInit(dist, rank, world, max_ctas=max_ctas, copy_engines=copy_engines)
            except Exception as e:   # noqa: BLE001
                self.impl, self.note = "torch", "native exchange unavailable: " + str(e).splitlines()[0][:120]
            ok = torch.tensor([1 if self.impl == "native" else 0], device=torch.device("cuda", torch.cuda.current_device()))
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)   # all ranks or none
            if int(ok.item()) == 0 and self.impl == "native":
                mpc.exchangeDestroy(); self.impl, self.note = "torch", "native exchange unavailable on another rank"
        if self.impl == "torch" and slab_provider is None and max_ctas > 0:
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(max_ctas); opts.config.min_ctas = 1
                self.group = dist.new_group(ranks=list(range(world)), backend="nccl", pg_options=opts)
            except Exception as e:   # noqa: BLE001
                self.note += " (no CTA cap: %s)" % str(e).splitlines()[0][:80]
        self.slab_provider = slab_provider or self._library_slab
        self.cuda = slab_provider is None
        self.out = [None, None]
        self.events = []
        self.ticks = 0
        self.bytes_per_tick = 0
        self._keep = []
        if self.cuda:
            self.device = torch.device("cuda", torch.cuda.current_device())
            self.compute_stream = compute_stream or torch.cuda.ExternalStream(mpc.stream(), device=self.device)
            self.side = torch.cuda.Stream(device=self.device)

    # -- newest policy slab of the library as a tensor (no copy)
    def _library_slab(self):
        v = self.mpc.getDeviceView(inflight=True)
        n = int(v.slab_bytes) // 8

        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(v.slab), False), "version": 3, "strides": None}
        t = self.torch.as_tensor(a, device=self.device)
        self._keep = self._keep[-8:] + [a]
        return t

    def _window_of(self, slab):
        """First `window_nodes` nodes of K, uff, x, u, times of every instance, packed into one send buffer."""
        m = self.mpc
        B, NS, nx, nu, k = m.batch, m.max_nodes, m.nx, m.nu, self.window_nodes
        sizes = [B * NS * nu * nx, B * NS * nu, B * NS * nx, B * NS * nu, B * NS]
        parts, o = [], 0
        for sz in sizes:
            parts.append(slab[o:o + sz].view(B, NS, -1)[:, :k].reshape(-1))
            o += sz
        return self.torch.cat(parts)

    def _gather(self, slab):
        src = self._window_of(slab) if self.window else slab
        i = self.ticks & 1
        if self.out[i] is None or self.out[i].numel() != self.world * src.numel():
            self.out[i] = self.torch.empty(self.world * src.numel(), dtype=src.dtype, device=src.device)
        self.dist.all_gather_into_tensor(self.out[i], src, group=self.group)
        self.bytes_per_tick = src.numel() * 8 * self.world
        return self.out[i]

    def after_tick(self):
        """Call right after bmpc_advance_async: enqueues the all-gather of the policy that tick produces."""
        torch = self.torch
        if self.impl == "native":
            self.mpc.exchangeStart()
            self.ticks += 1
            self.bytes_per_tick = self.mpc.getDeviceView(inflight=True).slab_bytes * self.world
            return None
        if not self.cuda:
            res = self._gather(self.slab_provider())
            self.ticks += 1
            return res
        slab = self.slab_provider()
        ready = torch.cuda.Event()
        ready.record(self.compute_stream)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            res = self._gather(slab)
            done = torch.cuda.Event()
            done.record(self.side)
        self.events.append(done)
        self.events = self.events[-4:]
        self.ticks += 1
        return res

    def before_tick(self):
        """Call before bmpc_advance_async: the tick about to start overwrites the slab that was gathered two ticks ago."""
        if self.impl == "native":
            return   # the library orders the tick after the gather that still reads its slab
        if self.cuda and len(self.events) >= 2:
            self.compute_stream.wait_event(self.events[-2])

    def join(self, stream=None):
        if self.impl == "native":
            self.mpc.exchangeWait()
        elif self.cuda:
            (stream or self.compute_stream).wait_stream(self.side)

    def shard(self, gathered, r):
        """View of rank r's part of a gathered buffer."""
        n = gathered.numel() // self.world
        return gathered[r * n:(r + 1) * n]

    def describe(self):
        return {"collectives_per_tick": 1, "bytes_received_per_rank_per_tick": int(self.bytes_per_tick), "what": "first %d nodes of every instance" % self.window_nodes if self.window else "whole policy slab [K | uff | x | u | times | events | n_nodes], node slots sized to the workload",
                "backend": ("libbmpc bmpc_exchange_* (ncclAllGather%s)" % ["", ", copy engines, zero SMs", ", symmetric windows"][self.mpc.exchangeView()[3]] if self.impl == "native"
                            else ("torch.distributed all_gather_into_tensor (NCCL)" if self.cuda else "gloo")),
                "max_ctas": self.max_ctas, "note": self.note}
