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
        self._shard_size = None
        if reader_name is not None:
            self._init_from_reader(last_batch_padded)
        self._first = None
        if prepare_first_batch:
            try:
                self._first = self._fetch()
            except StopIteration:
                self._first = None

    def _init_from_reader(self, last_batch_padded):
        """plugin/base_iterator.py:_extract_from_reader_and_validate: the epoch length comes from the reader's meta data."""
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
        self._shard_ids = [m["shard_id"] for m in metas]
        if self._policy == LastBatchPolicy.DROP:
            self._size = n // shards
        elif self._last_batch_padded:
            self._size = metas[0]["epoch_size_padded"] // shards
        else:
            self._size = int(math.ceil(math.ceil(n / shards) / self.batch_size)) * self.batch_size
        self._epoch = 0

    def _shard_sizes_now(self):
        res = []
        for sid in self._shard_ids:
            v = sid if self._stick else (sid + self._epoch) % self._shards_num
            res.append(self._size_no_pad * (v + 1) // self._shards_num - self._size_no_pad * v // self._shards_num)
        return res

    def _fetch(self):
        torch = self._torch
        res = []
        for p in self._pipes:
            outs = p.run()          # keeps `prefetch_queue_depth` batches in flight and returns the oldest
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
            res.append(d)
            p.release_outputs()
        return res

    def __iter__(self):
        return self

    def __next__(self):
        if self._size > 0 and self._counter >= self._size:
            if self._auto_reset:
                self.reset()
            raise StopIteration
        if self._reader_name is not None and self._policy == LastBatchPolicy.DROP and \
                any(self._counter + self.batch_size > s for s in self._shard_sizes_now()):
            # the incomplete last batch is dropped: it is the next batch of the stream, so it is fetched and discarded
            if self._counter < max(self._shard_sizes_now()):
                try:
                    if self._first is not None:
                        self._first = None
                    else:
                        self._fetch()
                except StopIteration:
                    pass
            self._counter = self._size if self._size > 0 else self._counter
            if self._auto_reset:
                self.reset()
            raise StopIteration
        if self._first is not None:
            out, self._first = self._first, None
        else:
            try:
                out = self._fetch()
            except StopIteration:
                if self._auto_reset:
                    self.reset()
                raise
        if self._reader_name is not None:
            self._counter += self.batch_size                    # per-GPU counter, as in the reference
            if self._policy == LastBatchPolicy.PARTIAL:
                for d, ssz in zip(out, self._shard_sizes_now()):
                    left = self.batch_size - (self._counter - ssz)
                    if left < self.batch_size:                  # the tail of this batch is padding / wrapped-around samples
                        for k in list(d):
                            d[k] = d[k][:max(left, 0)]
        else:
            self._counter += self.batch_size * len(self._pipes)
        return out

    next = __next__

    def reset(self):
        self._counter = 0
        if self._reader_name is not None:
            self._epoch += 1
        for p in self._pipes:
            p.reset()

    def __len__(self):
        if self._size < 0:
            raise TypeError("size is unknown (-1)")
        return (self._size + self.batch_size * len(self._pipes) - 1) // (self.batch_size * len(self._pipes))

    @property
    def size(self):
        return self._size


class DALIClassificationIterator(DALIGenericIterator):
    def __init__(self, pipelines, *a, **kw):
        super().__init__(pipelines, ["data", "label"], *a, **kw)
