"""fn.readers.file and the fn.random.* argument generators -- host-side input staging in front of the hot path
(SURVEY.md 8f rank 2).  They produce CPU batches exactly like an fn.external_source callback does, so everything behind
them (mixed decoder, GPU operators, prefetch slots) is unchanged.

Reference behaviour mirrored here:
  * file discovery      dali/operators/reader/loader/file_label_loader.{h,cc}, filesystem.cc, utils.h:28-36: `file_root` is
                        traversed one level deep, class directories in sorted order -> label = directory index, files sorted
                        inside a directory, only files with a known image / audio extension; `file_list` = text file of
                        "<relative path> <label>" lines; `files` (+ optional `labels`) = explicit lists.
  * sharding / epochs   dali/operators/reader/loader/loader.{h,cc}: a shard starts at N * shard_id / num_shards
                        (loader.cc:78-81) and is read sequentially; behind its end the loader continues with the NEXT shard
                        (the shards rotate from epoch to epoch) unless `stick_to_shard`, which wraps to the shard's own start;
                        `pad_last_batch` repeats the last sample until every shard has returned ceil(N / num_shards) samples
                        and the batch is full (loader.h:206-216, 262-283); without it batches simply run on into the next
                        epoch's samples.  `reader_meta()` has the reference's keys (pipeline.py reader_meta).
  * shuffling           `random_shuffle`: a buffer of `initial_fill` samples from which a random one is returned and replaced
                        by the next one read (loader.h:218-330); `shuffle_after_epoch`: the whole file list is re-shuffled
                        with the seed + epoch before every epoch and sharding is by stick_to_shard (file_label_loader.h).
                        The random ORDER is not the reference's (std::mt19937 streams cannot be reproduced from numpy); the
                        distribution and the epoch / shard structure are.
  * random.coin_flip / random.uniform   dali/operators/random/{coin_flip,uniform_distribution}_cpu.cc: one value (or `shape`)
                        per sample, int32 for coin_flip, float32 for uniform (`range` continuous, `values` discrete).
"""
import math
import os

import numpy as np

KNOWN_EXTENSIONS = (".jpg", ".jpeg", ".png", ".bmp", ".tif", ".tiff", ".pnm", ".ppm", ".pgm", ".pbm", ".jp2", ".webp",
                    ".flac", ".ogg", ".wav")


def start_index(shard_id, num_shards, size):
    """loader.cc:78-81."""
    return size * shard_id // num_shards


def discover_files(file_root=None, file_list=None, files=None, labels=None):
    """[(path, label)] in the reference's order."""
    if files is not None:
        if file_list is not None:
            raise ValueError("`files` and `file_list` are mutually exclusive")
        paths = [os.path.join(file_root, f) if file_root else f for f in files]
        if labels is None:
            labels = list(range(len(paths)))
        if len(labels) != len(paths):
            raise ValueError(f"Provided {len(labels)} labels for {len(paths)} files")
        return list(zip(paths, [int(v) for v in labels]))
    if labels is not None:
        raise ValueError("`labels` requires `files`")
    if file_list is not None:
        root = file_root or os.path.dirname(os.path.abspath(file_list))
        out = []
        with open(file_list) as f:
            for ln in f:
                ln = ln.strip()
                if not ln:
                    continue
                path, _, lab = ln.rpartition(" ")
                if not path:
                    raise ValueError(f"file_list line without a label: {ln!r}")
                out.append((os.path.join(root, path), int(lab)))
        return out
    if file_root is None:
        raise ValueError("One of `file_root`, `file_list` or `files` is required")
    out = []
    classes = sorted(d for d in os.listdir(file_root) if os.path.isdir(os.path.join(file_root, d)))
    for label, d in enumerate(classes):
        for f in sorted(os.listdir(os.path.join(file_root, d))):
            p = os.path.join(file_root, d, f)
            if os.path.isfile(p) and f.lower().endswith(KNOWN_EXTENSIONS):
                out.append((p, label))
    return out


class FileReader:
    """Stateful batch source with the reference loader's shard / epoch / padding / shuffling rules."""

    def __init__(self, batch_size, file_root=None, file_list=None, files=None, labels=None, random_shuffle=False,
                 shuffle_after_epoch=False, initial_fill=1024, shard_id=0, num_shards=1, stick_to_shard=False, pad_last_batch=False,
                 seed=-1, shuffle_after_epoch_seed=None):
        if not (0 <= shard_id < num_shards):
            raise ValueError("num_shards needs to be greater than shard_id")
        if random_shuffle and shuffle_after_epoch:
            raise ValueError("shuffle_after_epoch and random_shuffle cannot be both true")
        if shuffle_after_epoch and stick_to_shard:
            raise ValueError("shuffle_after_epoch and stick_to_shard cannot be both true")
        self.entries = discover_files(file_root, file_list, files, labels)
        if not self.entries:
            raise RuntimeError("No files found.")
        if num_shards > len(self.entries):
            raise RuntimeError(f"The number of input samples: {len(self.entries)}, needs to be at least equal to the requested "
                               f"number of shards: {num_shards}.")
        self.batch_size, self.shard_id, self.num_shards = batch_size, shard_id, num_shards
        self.random_shuffle, self.shuffle_after_epoch = random_shuffle, shuffle_after_epoch
        # file_label_loader.h:134-138: shuffle_after_epoch implies stick_to_shard (every epoch is a new global permutation, the
        # shard index stays) and reader_meta reports it; the permutation seed must be the SAME on every rank
        # (`shuffle_after_epoch_seed`, default kDaliDataloaderSeed = 524287, file_label_loader.h:61-63,236-237) -- never the
        # per-rank pipeline / operator seed, or the shards would stop partitioning the data set.
        self.stick_to_shard, self.pad_last_batch = bool(stick_to_shard or shuffle_after_epoch), pad_last_batch
        self.shuffle_after_epoch_seed = 524287 if shuffle_after_epoch_seed is None else int(shuffle_after_epoch_seed)
        self.initial_fill = max(1, int(initial_fill)) if random_shuffle else 1
        self.seed = 524287 if seed is None or seed < 0 else int(seed)
        self.rng = np.random.default_rng(self.seed)
        self.order = list(range(len(self.entries)))
        # read side: sequential position inside the (virtual) shard of the epoch being READ
        self.read_epoch = 0
        self.virtual_shard = shard_id
        self._reshuffle()
        self.cursor = start_index(self.virtual_shard, num_shards, len(self.entries))
        self.read_in_shard = 0
        # return side: the epoch whose samples are being RETURNED (the shuffle buffer lets the reads run ahead)
        self.cur_epoch = 0
        self.returned_in_epoch = 0
        self.buffer = []                # (epoch tag, sample index), in read order
        self.last = None

    # ---- reference: reader_meta keys
    def meta(self):
        n = len(self.entries)
        return {"epoch_size": n, "epoch_size_padded": int(math.ceil(n / self.num_shards)) * self.num_shards,
                "number_of_shards": self.num_shards, "shard_id": self.shard_id, "pad_last_batch": self.pad_last_batch,
                "stick_to_shard": self.stick_to_shard}

    def _reshuffle(self):
        if self.shuffle_after_epoch:
            seed = (self.shuffle_after_epoch_seed + ((self.read_epoch + 1) << 32)) & 0xFFFFFFFFFFFFFFFF      # Reset(): ++epoch first
            self.order = list(np.random.default_rng(seed).permutation(len(self.entries)))

    def _shard_bounds(self, shard):
        n = len(self.entries)
        return start_index(shard, self.num_shards, n), start_index(shard + 1, self.num_shards, n)

    def _read_one(self):
        """Sequential read with the shard switch of loader.h (IncreaseReadSampleCounter / MoveToNextShard / Reset)."""
        item = (self.read_epoch, self.order[self.cursor])
        self.cursor += 1
        self.read_in_shard += 1
        lo, hi = self._shard_bounds(self.virtual_shard)
        if self.read_in_shard >= hi - lo:                    # the shard has been read completely: next epoch
            self.read_in_shard = 0
            self.read_epoch += 1
            if not self.stick_to_shard:
                self.virtual_shard = (self.virtual_shard + 1) % self.num_shards
            self._reshuffle()
            self.cursor = self._shard_bounds(self.virtual_shard)[0]
        return item

    def _next_sample(self, first_in_batch):
        while len(self.buffer) < self.initial_fill:
            self.buffer.append(self._read_one())
        cand = [k for k, (t, _) in enumerate(self.buffer) if t == self.cur_epoch]
        if not cand:
            # the epoch's samples are exhausted.  pad_last_batch (loader.h ShouldPadBatch): repeat the last sample until every
            # shard has returned ceil(N / num_shards) samples AND the batch is complete
            target = int(math.ceil(len(self.entries) / self.num_shards))
            if self.pad_last_batch and (self.returned_in_epoch < target or not first_in_batch):
                self.returned_in_epoch += 1
                return self.last
            self.cur_epoch += 1
            self.returned_in_epoch = 0
            cand = [k for k, (t, _) in enumerate(self.buffer) if t == self.cur_epoch]
        k = cand[int(self.rng.integers(0, len(cand)))] if self.random_shuffle else cand[0]
        idx = self.buffer.pop(k)[1]
        self.returned_in_epoch += 1
        self.last = idx
        return idx

    def enable_pinned(self, num_buffers, device=None):
        """Read the files into a ring of page-locked arenas (one per batch in flight, `num_buffers` = prefetch depth + 1): the
        mixed decoder then copies the streams to the GPU by DMA straight from the reader's buffers, as the reference's readers
        feed its mixed operators from their own pinned buffers."""
        self._pin_ring = [None] * max(2, int(num_buffers))
        self._pin_next = 0
        self._pin_device = device            # the arenas are allocated on the read-ahead thread: with THIS GPU current, not device 0

    def _arena(self, nbytes):
        from . import capi
        k = self._pin_next
        self._pin_next = (k + 1) % len(self._pin_ring)
        if self._pin_ring[k] is None or self._pin_ring[k].size < nbytes:
            size = max(nbytes + nbytes // 4, 1 << 20)
            dev = getattr(self, "_pin_device", None)
            self._pin_ring[k] = capi.pinned_empty(size) if dev is None else capi.pinned_empty(size, dev)
        return self._pin_ring[k]

    def enable_prefetch(self, ahead=2):
        """Read `ahead` batches ahead of the consumer on a background thread (the reference's loader fills its queue on a thread of
        its own, loader.h PrefetchWorker): file I/O then overlaps the GPU work instead of sitting on the thread that schedules the
        pipeline.  The sequence of batches is unchanged -- the reader's state is only advanced by that one thread.  With page-locked
        arenas the ring must hold the batches in flight in the pipeline, the `ahead` queued ones and the one being read."""
        import queue
        import threading
        if getattr(self, "_q", None) is not None:
            return
        self._ahead = max(1, int(ahead))
        self._q = queue.Queue(maxsize=self._ahead)
        self._stop = False

        import weakref
        ref, q = weakref.ref(self), self._q

        def producer():
            # holds the reader only while a batch is being read: a reader dropped with its pipeline ends the thread (and frees its arenas)
            while True:
                r = ref()
                if r is None or r._stop:
                    return
                try:
                    item = r._produce()
                except BaseException as ex:          # surfaced on the consumer's thread
                    item = ex
                del r
                while True:
                    r = ref()
                    if r is None or r._stop:
                        return
                    del r
                    try:
                        q.put(item, timeout=0.2)
                        break
                    except queue.Full:
                        continue
                if isinstance(item, BaseException):
                    return
        self._thr = threading.Thread(target=producer, name="dali_b200_reader_prefetch", daemon=True)
        self._thr.start()

    def close(self):
        self._stop = True

    def __call__(self, _iteration=None):
        if getattr(self, "_q", None) is not None:
            if getattr(self, "_dead", None) is not None:        # the read-ahead thread stopped at an error: keep reporting it
                raise self._dead
            item = self._q.get()
            if isinstance(item, BaseException):
                self._dead = item
                raise item
            return item
        return self._produce()

    def _produce(self):
        paths, labels = [], []
        for i in range(self.batch_size):
            idx = self._next_sample(i == 0)
            path, lab = self.entries[idx]
            paths.append(path)
            labels.append(np.array([lab], np.int32))
        pool = self._read_pool()
        if getattr(self, "_pin_ring", None) is None:
            return list(pool.map(lambda p: np.fromfile(p, dtype=np.uint8), paths)), labels
        sizes = list(pool.map(os.path.getsize, paths))
        offs = np.concatenate([[0], np.cumsum([(s + 63) & ~63 for s in sizes])])
        arena = self._arena(int(offs[-1]))
        views = [arena[int(o):int(o) + s] for s, o in zip(sizes, offs[:-1])]

        def read(job):
            p, s, view = job
            with open(p, "rb") as f:
                got = f.readinto(memoryview(view))          # releases the GIL: the files of a batch are read concurrently
            if got != s:
                raise IOError(f"short read from {p}: {got} of {s} bytes")
        list(pool.map(read, zip(paths, sizes, views)))
        return views, labels

    def _read_pool(self):
        """File reads of a batch run on a small thread pool (the reference's loader reads ahead on its own thread; at 256 x 0.5 MB per
        batch a sequential Python loop would cap the pipeline at ~10 k images/s)."""
        if getattr(self, "_pool", None) is None:
            import concurrent.futures as cf
            try:
                ncpu = len(os.sched_getaffinity(0))
            except AttributeError:
                ncpu = os.cpu_count() or 4
            self._pool = cf.ThreadPoolExecutor(max_workers=max(2, min(16, ncpu)), thread_name_prefix="dali_b200_reader")
        return self._pool


class CoinFlip:
    def __init__(self, batch_size, probability=0.5, shape=None, seed=-1, dtype=np.int32):
        self.batch_size, self.p, self.shape, self.dtype = batch_size, float(probability), shape, dtype
        self.rng = np.random.default_rng(None if seed is None or seed < 0 else seed)

    def __call__(self, _iteration=None):
        shp = tuple(self.shape) if self.shape is not None else ()
        return [np.asarray(self.rng.random(shp) < self.p, self.dtype) for _ in range(self.batch_size)]


class Uniform:
    def __init__(self, batch_size, range=(-1.0, 1.0), values=None, shape=None, seed=-1, dtype=np.float32):
        self.batch_size, self.shape, self.dtype = batch_size, shape, dtype
        self.values = None if values is None else np.asarray(values, dtype)
        self.lo, self.hi = float(range[0]), float(range[1])
        if self.values is None and not self.lo < self.hi:
            raise ValueError(f"Invalid range. It shall be left-closed [a, b), where a < b. Got: [{self.lo}, {self.hi})")
        self.rng = np.random.default_rng(None if seed is None or seed < 0 else seed)

    def __call__(self, _iteration=None):
        shp = tuple(self.shape) if self.shape is not None else ()
        if self.values is not None:
            return [np.asarray(self.values[self.rng.integers(0, len(self.values), shp)], self.dtype) for _ in range(self.batch_size)]
        return [np.asarray(self.rng.uniform(self.lo, self.hi, shp), self.dtype) for _ in range(self.batch_size)]
