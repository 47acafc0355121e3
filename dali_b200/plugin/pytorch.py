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
        if reader_name is not None:
            raise NotImplementedError("reader_name: file readers are outside the hot path; feed data through fn.external_source")
        self._size = size
        self._auto_reset = auto_reset
        self._counter = 0
        for p in self._pipes:
            p.build()
        self.batch_size = self._pipes[0].max_batch_size
        self._first = None
        if prepare_first_batch:
            try:
                self._first = self._fetch()
            except StopIteration:
                self._first = None

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
        if self._first is not None:
            out, self._first = self._first, None
        else:
            try:
                out = self._fetch()
            except StopIteration:
                if self._auto_reset:
                    self.reset()
                raise
        self._counter += self.batch_size * len(self._pipes)
        return out

    next = __next__

    def reset(self):
        self._counter = 0
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
