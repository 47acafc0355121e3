"""Pipeline definition and execution -- the nvidia.dali.pipeline surface for the hot path
(dali/python/nvidia/dali/pipeline.py: Pipeline :1202 build, :1515 run; pipeline_def).

Graph construction mirrors the reference: fn.* calls executed inside a pipeline definition create operator nodes in
the *current* pipeline and return DataNodes; build() hands the OpSpecs to the C++ Pipeline (dali_b200/host), which
instantiates the operators through the registry (DALI_REGISTER_OPERATOR) and runs them in order on one CUDA stream.
"""
import functools
import threading

import numpy as np

from . import backend, types

_tls = threading.local()


def _current():
    return getattr(_tls, "pipe", None)


class DataNode:
    """An edge of the graph (dali/python/nvidia/dali/data_node.py)."""

    def __init__(self, name, device, source=None):
        self.name, self.device, self.source = name, device, source

    def gpu(self):
        if self.device == "gpu":
            return self
        pipe = _current()
        if pipe is None:
            raise RuntimeError("DataNode.gpu() must be called inside a pipeline definition")
        return pipe._to_gpu(self)

    def __repr__(self):
        return f"DataNode(name={self.name!r}, device={self.device!r})"


class TensorListGPU:
    """Output batch living in device memory.  Uniform batches expose __cuda_array_interface__ (zero copy into torch)."""

    def __init__(self, dtype, layout, contiguous, shapes, ptrs, stream, owner=None):
        self._dtype, self._layout, self._contig, self._shapes, self._ptrs, self._stream = dtype, layout, contiguous, shapes, ptrs, stream
        self._owner = owner          # the pipeline that owns the device buffers stays alive as long as its outputs are referenced

    def __len__(self):
        return len(self._ptrs)

    def shape(self):
        return [tuple(int(v) for v in s) for s in self._shapes]

    @property
    def dtype(self):
        return types.DALIDataType(self._dtype)

    def layout(self):
        return self._layout

    def is_dense_tensor(self):
        return len(self._ptrs) > 0 and all(tuple(s) == tuple(self._shapes[0]) for s in self._shapes)

    def sample(self, i):
        return _CudaArray(self._ptrs[i], tuple(int(v) for v in self._shapes[i]), self._dtype)

    def __getitem__(self, i):
        return self.sample(i)

    def as_tensor(self):
        """The batch as one [N, ...] device tensor.  Samples are stored 256-byte aligned, so a zero-copy dense view
        exists only when the per-sample byte size is a multiple of 256; otherwise the samples are gathered."""
        if not self.is_dense_tensor():
            raise RuntimeError("The batch is not uniform; use .sample(i) / at(i)")
        shp = tuple(int(v) for v in self._shapes[0])
        nbytes = int(np.prod(shp)) * types.to_numpy_type(self._dtype).itemsize
        if self._contig and (len(self._ptrs) == 1 or self._ptrs[1] - self._ptrs[0] == nbytes):
            return _CudaArray(self._ptrs[0], (len(self._ptrs),) + shp, self._dtype)
        import torch
        return torch.stack([torch.as_tensor(self.sample(i), device="cuda") for i in range(len(self))])

    def as_cpu(self):
        import torch
        return [torch.as_tensor(self.sample(i), device="cuda").cpu().numpy() for i in range(len(self))]


class _CudaArray:
    def __init__(self, ptr, shape, dtype):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": types.to_numpy_type(dtype).str, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class TensorListCPU:
    def __init__(self, arrays, layout):
        self._arrays, self._layout = arrays, layout

    def __len__(self):
        return len(self._arrays)

    def at(self, i):
        return self._arrays[i]

    def __getitem__(self, i):
        return self._arrays[i]

    def layout(self):
        return self._layout


class _ExternalSourceGroup:
    def __init__(self, source, outputs, batch, cycle, layout, dtype, device, batch_info, no_copy=False):
        self.source, self.outputs, self.batch, self.cycle = source, outputs, batch, cycle
        self.no_copy = bool(no_copy)
        self.layout, self.dtype, self.device, self.batch_info = layout, dtype, device, batch_info
        self.iterator = None
        self.is_callable = callable(source) and not hasattr(source, "__iter__")


class Pipeline:
    def __init__(self, batch_size=-1, num_threads=-1, device_id=-1, seed=-1, exec_pipelined=True, prefetch_queue_depth=2,
                 exec_async=True, exec_dynamic=False, **_ignored):
        if batch_size is None or batch_size <= 0:
            raise ValueError("batch_size must be a positive integer")
        self.max_batch_size = int(batch_size)
        self.num_threads = max(1, int(num_threads) if num_threads and num_threads > 0 else 1)
        self.device_id = None if device_id is None or device_id < 0 else int(device_id)
        self.seed = seed
        self._nodes = []            # (schema, inst_name, spec)
        self._externals = []        # _ExternalSourceGroup
        self._ext_names = {}
        self._outputs = []
        self._counter = 0
        self._built = False
        self._backend = None
        self._slots = []            # one C++ pipeline (stream, buffers, operator plans) per prefetch slot
        self._depth = max(1, int(prefetch_queue_depth) if exec_pipelined else 1)
        self._sched = []            # slots with a batch in flight, oldest first
        self._next_slot = 0
        self._exhausted = False
        self._definition = None
        self._iteration = 0
        self._epoch_idx = 0
        self._sample_idx = 0
        self._keepalive = []
        self._cur_slot = 0
        self._pending = []
        self._readers = {}          # reader instance name -> dali_b200.readers.FileReader
        self._op_graph = {}         # operator instance -> (input names, output names, has a seed argument, seed left open)

    # ---- context management (with pipe: ...)
    def __enter__(self):
        self._prev = _current()
        _tls.pipe = self
        return self

    def __exit__(self, *exc):
        _tls.pipe = self._prev
        return False

    @property
    def batch_size(self):
        return self.max_batch_size

    def _new_name(self, base):
        self._counter += 1
        return f"__{base}_{self._counter}"

    def _to_gpu(self, node):
        # An external-source output that has not been consumed on the CPU is simply produced on the GPU
        # (the reference inserts a MakeContiguous copy node; the H2D copy happens in feed_input here).
        if isinstance(node.source, _ExternalSourceGroup) and not getattr(node, "_consumed_cpu", False):
            node.device = "gpu"
            node.source.device = "gpu"
            return node
        raise RuntimeError(".gpu() is supported on fn.external_source outputs only; use device='mixed' / 'gpu' operators")

    def set_outputs(self, *nodes):
        self._outputs = list(nodes)

    # ---- build: python graph -> C++ pipeline
    def build(self):
        if self._built:
            return self
        definition = self._definition
        if definition is None and not self._outputs and callable(getattr(self, "define_graph", None)):
            definition = self.define_graph              # legacy subclass style (pipeline.py: Pipeline.define_graph)
        if definition is not None and not self._outputs:
            with self:
                outs = definition()
            if isinstance(outs, DataNode):
                outs = (outs,)
            self._outputs = list(outs)
        if not self._outputs:
            raise RuntimeError("Pipeline has no outputs; call set_outputs() or use @pipeline_def")
        for o in self._outputs:
            if not isinstance(o, DataNode):
                raise TypeError(f"Pipeline outputs must be DataNodes, got {type(o).__name__}")
        # prefetch_queue_depth independent executor slots (the reference's queue depth, exec2.h:66-131): while the GPU works on
        # batch i, the host parses / stages / uploads batch i+1 into the other slot.
        self._assign_seeds()
        for schema, inst, spec in self._nodes:
            spec.add_arg("_state_key", f"{id(self)}:{inst}")       # operator instances of one node share their random state
        for _ in range(self._depth):
            be = backend.Pipeline(self.max_batch_size, self.num_threads, self.device_id)
            for g in self._externals:
                for o in g.outputs:
                    be.add_external_input(o.name, o.device, g.layout or "")
            for schema, inst, spec in self._nodes:
                be.add_operator(spec, inst)
            be.set_outputs([(o.name, o.device) for o in self._outputs])
            be.build()
            self._slots.append(be)
        self._backend = self._slots[0]
        self._keep = [[] for _ in self._slots]
        self._built = True
        return self

    def _assign_seeds(self):
        """Seeds of the operators the user did not seed, as the reference derives them: a table of 1024 values generated from the
        pipeline seed (pipeline.cc:303-308; the clock when there is none), handed out in the order Pipeline::AddOperator sees the
        operators that take a seed (pipeline.cc:823-831) -- the depth-first, inputs-first order from the outputs of
        pipeline.py:2423-2463 _collect_ops; unreachable operators are pruned and take no seed."""
        from . import readers
        producers = {}
        for schema, inst, spec in self._nodes:
            for name in self._op_graph.get(inst, ((), (), False, False))[1]:
                producers[name] = inst
        for g in self._externals:
            for o in g.outputs:
                producers[o.name] = g
        order, visited = [], set()

        def visit(name):
            prod = producers.get(name)
            if prod is None:
                return
            key = prod if isinstance(prod, str) else id(prod)
            if key in visited:
                return
            visited.add(key)
            if isinstance(prod, str):
                for n in self._op_graph[prod][0]:
                    visit(n)
            order.append(prod)
        for o in self._outputs:
            visit(o.name)
        specs = {inst: spec for _, inst, spec in self._nodes}
        table, k = None, 0
        for prod in order:
            if isinstance(prod, str):
                _, _, seeded, open_seed = self._op_graph[prod]
                if not (seeded and open_seed):
                    continue
            else:
                src = prod.source
                if not hasattr(src, "set_seed") or getattr(src, "seed", 0) >= 0:
                    continue
            if table is None:
                base = self.seed if self.seed is not None and self.seed >= 0 else readers._clock_seed()
                table = readers.seed_table(base)
            if isinstance(prod, str):
                specs[prod].add_arg("seed", int(table[k]))
            else:
                prod.source.set_seed(int(table[k]))
            k = (k + 1) % len(table)
        self._seed_order = [p if isinstance(p, str) else p.outputs[0].name for p in order]

    # ---- external source feeding (external_source.py:312-1150, _run_input_callbacks)
    def _next_batch(self, g):
        if g.is_callable:
            if g.batch:
                arg = types.BatchInfo(self._iteration, self._epoch_idx) if g.batch_info else self._iteration
                try:
                    return g.source(arg)
                except TypeError:
                    return g.source()
            res = []
            for i in range(self.max_batch_size):
                res.append(g.source(types.SampleInfo(self._sample_idx + i, i, self._iteration, self._epoch_idx)))
            return res
        if g.iterator is None:
            g.iterator = iter(g.source)
        try:
            return next(g.iterator)
        except StopIteration:
            if g.cycle in (True, "quiet", "raise"):
                g.iterator = iter(g.source)
                if g.cycle == "raise":
                    raise
                return next(g.iterator)
            raise

    def feed_input(self, name_or_node, data, layout=None):
        name = name_or_node.name if isinstance(name_or_node, DataNode) else self._ext_names.get(name_or_node, name_or_node)
        self._cur_slot = self._next_slot
        self._feed(name, data, layout or "")

    def _feed(self, name, data, layout, no_copy=False):
        if hasattr(data, "cpu") and hasattr(data, "numpy") and not isinstance(data, np.ndarray):     # torch tensor batch
            data = data.cpu().numpy()
        if isinstance(data, np.ndarray):
            samples = [data[i] for i in range(data.shape[0])]
        else:
            samples = [s.cpu().numpy() if hasattr(s, "cpu") and not isinstance(s, np.ndarray) else np.asarray(s) for s in data]
        if len(samples) == 0:
            raise RuntimeError("External source returned an empty batch")
        if len(samples) > self.max_batch_size:
            raise RuntimeError(f"External source batch ({len(samples)}) exceeds max batch size ({self.max_batch_size})")
        dt = samples[0].dtype
        nd = samples[0].ndim
        arrs = []
        for s in samples:
            if s.dtype != dt or s.ndim != nd:
                raise TypeError("All samples of an external source batch must have the same type and dimensionality")
            arrs.append(np.ascontiguousarray(s))
        self._pending.append(arrs)          # kept alive until the slot that consumes them is reused (the backend borrows the pointers)
        shapes = np.array([a.shape for a in arrs], np.int64).reshape(len(arrs), nd) if nd else np.zeros((len(arrs), 0), np.int64)
        self._slots[self._cur_slot].feed_input(name, [a.ctypes.data for a in arrs], shapes, nd, int(types.from_numpy_type(dt)), layout,
                                               no_copy)

    def _run_input_callbacks(self, slot):
        self._cur_slot = slot
        if callable(getattr(self, "iter_setup", None)):
            self.iter_setup()                           # legacy hook: the subclass calls feed_input() here, once per iteration
        for g in self._externals:
            if g.source is None:
                continue
            batch = self._next_batch(g)
            if len(g.outputs) == 1:
                batch = (batch,)
            for o, b in zip(g.outputs, batch):
                self._feed(o.name, b, g.layout or "", g.no_copy)
        # everything fed for this iteration (feed_input() calls and source callbacks) replaces what the slot held before
        self._keep[slot] = self._pending
        self._pending = []

    # ---- run
    def schedule_run(self):
        """Feeds the external sources and enqueues one iteration on the next free slot (asynchronous)."""
        if not self._built:
            self.build()
        if len(self._sched) >= self._depth:
            raise RuntimeError("All prefetch slots are in flight; call share_outputs() / release_outputs() first")
        slot = self._next_slot
        self._run_input_callbacks(slot)
        self._slots[slot].run()
        self._sched.append(slot)
        self._next_slot = (slot + 1) % self._depth
        self._iteration += 1
        self._sample_idx += self.max_batch_size

    def share_outputs(self):
        if not self._sched:
            raise StopIteration
        slot = self._sched.pop(0)
        self._slots[slot].wait()
        return self._collect_outputs(slot)

    def release_outputs(self):
        pass

    def outputs(self):
        return self.share_outputs()

    def run(self):
        """One iteration; returns its outputs (valid until the slot is reused, i.e. for `prefetch_queue_depth` - 1 further
        run() calls -- the reference invalidates them at the next run()).  The following batches are scheduled before waiting."""
        if not self._built:
            self.build()
        while len(self._sched) < self._depth and not self._exhausted:
            try:
                self.schedule_run()
            except StopIteration:
                self._exhausted = True
        if not self._sched:
            self._exhausted = False
            raise StopIteration
        return self.share_outputs()

    def _collect_outputs(self, slot=0):
        outs = []
        be = self._slots[slot]
        stream = be.stream()
        for i in range(be.num_outputs()):
            gpu, dt, lay, cont, shapes, ptrs = be.output(i)
            if gpu:
                outs.append(TensorListGPU(dt, lay, cont, shapes, ptrs, stream, owner=self))
            else:
                import ctypes as C
                npdt = types.to_numpy_type(dt)
                arrs = []
                for s, p in zip(shapes, ptrs):
                    n = int(np.prod(s)) if len(s) else 1
                    buf = (C.c_char * (n * npdt.itemsize)).from_address(p) if n else b""
                    arrs.append(np.frombuffer(buf, npdt, n).reshape(tuple(int(v) for v in s)).copy())
                outs.append(TensorListCPU(arrs, lay))
        return tuple(outs)

    def reset(self):
        self._epoch_idx += 1
        self._sample_idx = 0
        self._exhausted = False
        for g in self._externals:
            g.iterator = None

    def epoch_size(self, name=None):
        if name is not None:
            return self._readers[name].meta()["epoch_size_padded"]
        return {k: r.meta()["epoch_size_padded"] for k, r in self._readers.items()}

    def reader_meta(self, name=None):
        """dali/python/nvidia/dali/pipeline.py reader_meta: epoch_size, epoch_size_padded, number_of_shards, shard_id,
        pad_last_batch, stick_to_shard."""
        if name is not None:
            if name not in self._readers:
                raise KeyError(f"Reader '{name}' not found in the pipeline")
            return self._readers[name].meta()
        return {k: r.meta() for k, r in self._readers.items()}

    def executor_statistics(self):
        return {}


def pipeline_def(fn=None, **pipeline_kwargs):
    """@pipeline_def(batch_size=..., num_threads=..., device_id=...)"""
    def actual(func):
        @functools.wraps(func)
        def create(*args, **kwargs):
            ctor = dict(pipeline_kwargs)
            for k in ("batch_size", "num_threads", "device_id", "seed", "prefetch_queue_depth", "exec_async", "exec_pipelined",
                      "exec_dynamic", "py_num_workers", "py_start_method", "enable_conditionals"):
                if k in kwargs:
                    ctor[k] = kwargs.pop(k)
            pipe = Pipeline(**ctor)
            pipe._definition = lambda: func(*args, **kwargs)
            return pipe
        return create
    return actual(fn) if fn is not None else actual
