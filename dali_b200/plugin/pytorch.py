"""nvidia.dali.plugin.pytorch.DALIGenericIterator for dali_b200 pipelines
(dali/python/nvidia/dali/plugin/pytorch/__init__.py:43-283, plugin/base_iterator.py).

One dict of torch tensors per pipeline (= per GPU) and iteration.  Outputs are copied out of the pipeline's buffers on
the pipeline's device (feed_ndarray in the reference), so they stay valid while the next iteration runs.
"""
import numpy as np

from .. import types


class LastBatchPolicy:
    FILL, DROP, PARTIAL = 0, 1, 2


class DALIGenericIterator:
    def __init__(self, pipelines, output_map, size=-1, reader_name=None, auto_reset=False, fill_last_batch=None, dynamic_shape=False,
                 last_batch_padded=False, last_batch_policy=LastBatchPolicy.FILL, prepare_first_batch=True):
        import torch
        self._torch = torch
        self._pipes = pipelines if isinstance(pipelines, (list, tuple)) else [pipelines]
        self.output_map = list(output_map)
        if len(set(self.output_map)) != len(self.output_map):
            raise ValueError("output_map names should be distinct")
        if fill_last_batch is not None:      # deprecated spelling (plugin/base_iterator.py)
            last_batch_policy = LastBatchPolicy.FILL if fill_last_batch else LastBatchPolicy.PARTIAL
        if reader_name is not None and size != -1:
            raise ValueError("When reader_name is provided, size should not be set")
        self._size = size
        self._auto_reset = auto_reset
        self._counter = 0
        self._policy = last_batch_policy
        self._reader_name = reader_name
        for p in self._pipes:
            p.build()
        self.batch_size = self._pipes[0].max_batch_size
        self._last_batch_padded = bool(last_batch_padded)
        if reader_name is not None:
            self._init_from_reader(last_batch_padded)
        self._first = None
        self._pool = None
        if prepare_first_batch:
            try:
                self._first = self._fetch()
            except StopIteration:
                self._first = None

    def _init_from_reader(self, last_batch_padded):
        """plugin/base_iterator.py:306-371 (_extract_from_reader_and_validate): the epoch length comes from the reader's meta
        data; per-shard read-ahead counters keep the epochs aligned with the data when shards are uneven."""
        import math
        metas = [p.reader_meta(self._reader_name) for p in self._pipes]
        for k, what in (("epoch_size", "size value"), ("number_of_shards", "`num_shards` argument set"),
                        ("pad_last_batch", "`pad_last_batch` argument set"), ("stick_to_shard", "`stick_to_shard` argument set")):
            if any(m[k] != metas[0][k] for m in metas):
                raise AssertionError(f"Reader Operator should have the same {what} in all the pipelines.")
        n, shards = metas[0]["epoch_size"], metas[0]["number_of_shards"]
        self._size_no_pad, self._shards_num = n, shards
        self._last_batch_padded = metas[0]["pad_last_batch"]
        self._stick = metas[0]["stick_to_shard"]
        self._shards_id = np.array([m["shard_id"] for m in metas], np.int64)
        if self._policy == LastBatchPolicy.DROP:
            self._size = n // shards
        elif self._last_batch_padded:
            self._size = metas[0]["epoch_size_padded"] // shards
        else:
            self._size = int(math.ceil(math.ceil(n / shards) / self.batch_size)) * self.batch_size
        ids = np.arange(shards, dtype=np.int64)
        self._counter_per_gpu = np.zeros(shards, np.int64)          # where each shard starts inside this epoch (read-ahead)
        self._shard_sizes_per_gpu = (ids + 1) * n // shards - ids * n // shards
        self._shard_sizes_initial = self._shard_sizes_per_gpu.copy()

    def _run_all(self):
        """One iteration of every pipeline.  With several pipelines (one per GPU in one process, the layout of the reference's
        multi-GPU examples) the host work of an iteration -- feeding, header parsing, H2D submission, the wait for the oldest
        batch -- runs concurrently, one thread per pipeline (the native calls release the GIL)."""
        if len(self._pipes) == 1:
            return [self._pipes[0].run()]
        if self._pool is None:
            import concurrent.futures as cf
            self._pool = cf.ThreadPoolExecutor(max_workers=len(self._pipes), thread_name_prefix="dali_b200_iter")
        futs = [self._pool.submit(p.run) for p in self._pipes]
        return [f.result() for f in futs]          # a StopIteration / error of any pipeline propagates

    def _fetch(self):
        torch = self._torch
        res = []
        for p, outs in zip(self._pipes, self._run_all()):
            if len(outs) != len(self.output_map):
                raise RuntimeError(f"The pipeline has {len(outs)} outputs but output_map has {len(self.output_map)} names")
            d = {}
            dev = torch.device("cuda", p.device_id) if p.device_id is not None else torch.device("cpu")
            for name, o in zip(self.output_map, outs):
                if hasattr(o, "as_tensor"):
                    with torch.cuda.device(dev):
                        t = torch.as_tensor(o.as_tensor(), device=dev)
                        d[name] = t.clone()
                else:
                    d[name] = torch.from_numpy(np.stack([o.at(i) for i in range(len(o))]))
            if p.device_id is not None:
                # The copies above run on torch's current stream; the pipeline refills this slot on its own (non-blocking)
                # stream at the next run().  Like the reference's feed_ndarray (plugin/pytorch/__init__.py:222-226, which
                # copies on the torch stream and waits), block until the copies have left the slot before handing it back.
                torch.cuda.current_stream(dev).synchronize()
            res.append(d)
            p.release_outputs()
        return res

    def __iter__(self):
        return self

    def _end_iteration(self):
        if self._auto_reset:
            self.reset()
        raise StopIteration

    def _advance_and_check_drop_last(self, dry_run=False):
        """plugin/base_iterator.py:440-466."""
        counter, should_end = self._counter, False
        if self._reader_name is not None:
            counter += self.batch_size                              # per-GPU counter, as in the reference
            if self._policy == LastBatchPolicy.DROP:
                should_end = bool(np.any(self._counter_per_gpu + counter > self._shard_sizes_per_gpu))
        else:
            counter += self.batch_size * len(self._pipes)
            if self._policy == LastBatchPolicy.DROP:
                should_end = self._size > 0 and counter > self._size
        if not dry_run:
            self._counter = counter
        return should_end

    def _get(self):
        if self._first is not None:
            out, self._first = self._first, None
            return out
        return self._fetch()

    def __next__(self):
        if self._size > 0 and self._counter >= self._size:
            self._end_iteration()
        try:
            out = self._get()
        except StopIteration:
            if self._size < 0 and self._auto_reset:
                self.reset()
            raise
        if self._advance_and_check_drop_last():
            self._end_iteration()                                   # the incomplete last batch has been fetched and is dropped
        if self._reader_name is not None:
            if self._policy == LastBatchPolicy.PARTIAL:
                left = self.batch_size - (self._counter - self._shard_sizes_initial[self._shards_id])
                if np.any(left < self.batch_size):                  # the tail of this batch is padding / wrapped-around samples
                    out = [{k: v[:max(int(l), 0)] for k, v in d.items()} for d, l in zip(out, left)]
        elif self._policy == LastBatchPolicy.PARTIAL and self._size > 0 and self._counter > self._size:
            diff = len(self._pipes) * self.batch_size - (self._counter - self._size)
            ngrab = int(np.ceil(diff / self.batch_size))
            last = diff % self.batch_size or self.batch_size
            out = out[:ngrab]
            out[-1] = {k: v[:last] for k, v in out[-1].items()}
        return out

    next = __next__

    def reset(self):
        """plugin/base_iterator.py:489-566: with FILL and a reader that wraps into the next epoch (pad_last_batch=False) every
        GPU may have read ahead of its next shard; the counters start the next epoch from there and `size` is re-evaluated."""
        import math
        if self._policy == LastBatchPolicy.DROP:
            should_end = self._advance_and_check_drop_last(dry_run=True)
            already_ended = self._size > 0 and self._counter >= self._size
            if should_end and not already_ended:
                try:
                    self._get()                                     # the incomplete batch still sits in the pipeline: drop it
                except StopIteration:
                    pass
                self._advance_and_check_drop_last()
        if not (self._counter >= self._size or self._size < 0):
            import logging
            logging.warning("DALI iterator does not support resetting while epoch is not finished. Ignoring...")
            return
        fill_wrap = self._policy == LastBatchPolicy.FILL and not getattr(self, "_last_batch_padded", False)
        if fill_wrap:
            if self._reader_name is not None:
                self._counter -= int(self._counter_per_gpu.min())
                self._counter_per_gpu = self._counter_per_gpu + self._counter - self._shard_sizes_per_gpu
                self._counter = int(self._counter_per_gpu.min())
            elif self._size > 0:
                self._counter = self._counter % self._size
            else:
                self._counter = 0
        else:
            self._counter = 0
        if self._reader_name is not None:
            if not self._stick:
                self._shards_id = (self._shards_id + 1) % self._shards_num
            if fill_wrap:
                if not self._stick:
                    self._shard_sizes_per_gpu = np.roll(self._shard_sizes_per_gpu, 1)
                read_next = self._shard_sizes_per_gpu - self._counter_per_gpu
                self._size = int(math.ceil(int(read_next.max()) / self.batch_size)) * self.batch_size
                if self._size == 0:                                 # read so far ahead that the next epoch is already done
                    self._counter_per_gpu = np.zeros(self._shards_num, np.int64)
                    self._counter = 0
                    self._shard_sizes_per_gpu = np.roll(self._shard_sizes_per_gpu, 1)
                    self._size = int(math.ceil(int(self._shard_sizes_per_gpu.max()) / self.batch_size)) * self.batch_size
        for p in self._pipes:
            p.reset()

    def __len__(self):
        """plugin/base_iterator.py:600-616."""
        if self._size < 0:
            raise TypeError("size is unknown (-1)")
        import math
        if self._reader_name is not None:
            if self._policy != LastBatchPolicy.DROP:
                return int(math.ceil(self._size / self.batch_size))
            return self._size // self.batch_size
        if self._policy != LastBatchPolicy.DROP:
            return int(math.ceil(self._size / (len(self._pipes) * self.batch_size)))
        return self._size // (len(self._pipes) * self.batch_size)

    @property
    def size(self):
        return self._size


class DALIClassificationIterator(DALIGenericIterator):
    def __init__(self, pipelines, *a, **kw):
        super().__init__(pipelines, ["data", "label"], *a, **kw)
