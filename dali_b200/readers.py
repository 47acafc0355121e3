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
  * shuffling           `random_shuffle`: one global shuffle with the fixed data-loader seed, then a buffer of `initial_fill`
                        samples from which a random one is returned and replaced by the next one read (loader.h:225-342);
                        `shuffle_after_epoch`: the whole file list is re-shuffled with the seed + epoch before every epoch and
                        sharding is by stick_to_shard (file_label_loader.h:198-241).  The ORDER is the reference's: the loader's
                        state machine is restated in dali_b200/host/host_random.cc on the same C++ standard library engines and
                        distributions the reference instantiates (dalihSampleOrder*).
  * random.coin_flip / random.uniform   dali/operators/random/{coin_flip,uniform_distribution}.h, rng_base{,_cpu}.h, random_dist.h:
                        Philox4x32-10 addressed by (seed, samples generated so far + 65537 * sample, 257 * element); int32 for
                        coin_flip, float32 for uniform (`range` continuous, `values` discrete) -- the reference's numbers, bit
                        for bit (dalihRandomCoinFlip / dalihRandomUniform; pinned by tests/test_random_ref_cpu.py).
  * seeds               operators without a `seed` get theirs from the pipeline's seed table in the order the reference's
                        Pipeline::AddOperator would see them (dali_b200/pipeline.py _assign_seeds; pipeline.cc:303-308,823-831).
"""
import ctypes as C
import math
import os
import time

import numpy as np


def _host():
    """libdali_b200_host.so entry points of host_random.cc."""
    from . import backend
    L = backend.lib()
    if not getattr(L, "_random_ready", False):
        L.dalihRandomLastError.restype = C.c_char_p
        L._random_ready = True
    return L


def _hcheck(rc):
    if rc != 0:
        raise RuntimeError(_host().dalihRandomLastError().decode("utf-8", "replace"))


def seed_table(seed, n=1024):
    """The per-operator seeds the reference's pipeline derives from its `seed` (pipeline.cc:303-308)."""
    out = (C.c_int64 * n)()
    _hcheck(_host().dalihSeedTable(C.c_int64(int(seed)), out, n))
    return list(out)


def _clock_seed():
    return time.time_ns() & 0x7FFFFFFFFFFFFFFF

KNOWN_EXTENSIONS = (".jpg", ".jpeg", ".png", ".bmp", ".tif", ".tiff", ".pnm", ".ppm", ".pgm", ".pbm", ".jp2", ".webp",
                    ".flac", ".ogg", ".wav")


def start_index(shard_id, num_shards, size):
    """loader.cc:78-81."""
    return size * shard_id // num_shards


KNOWN_EXTENSIONS_GLOB = tuple("*" + e for e in KNOWN_EXTENSIONS)          # utils.h:34-36 kKnownExtensionsGlob


def _glob_match(name, filters, case_sensitive):
    """fnmatch(3) of discover_files.cc:64-67,100-105 (FNM_CASEFOLD unless case_sensitive_filter)."""
    import fnmatch
    if not case_sensitive:
        name = name.lower()
    return any(fnmatch.fnmatchcase(name, f if case_sensitive else f.lower()) for f in filters)


def discover_files(file_root=None, file_list=None, files=None, labels=None, file_filters=None, dir_filters=None,
                   case_sensitive_filter=False):
    """[(path, label)] in the reference's order (discover_files.cc:124-157: sub-directories sorted -> label, files sorted inside a
    directory, `file_filters` / `dir_filters` globs; the filters are ignored with `file_list` / `files`, file_reader_op.cc:128-138)."""
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
    if isinstance(file_filters, str):
        file_filters = [file_filters]
    if isinstance(dir_filters, str):
        dir_filters = [dir_filters]
    file_filters = list(file_filters) if file_filters else list(KNOWN_EXTENSIONS_GLOB)
    classes = sorted(d for d in os.listdir(file_root) if os.path.isdir(os.path.join(file_root, d))
                     and (not dir_filters or _glob_match(d, dir_filters, case_sensitive_filter)))
    for label, d in enumerate(classes):
        for f in sorted(os.listdir(os.path.join(file_root, d))):
            p = os.path.join(file_root, d, f)
            if os.path.isfile(p) and _glob_match(f, file_filters, case_sensitive_filter):
                out.append((p, label))
    return out


class FileReader:
    """Stateful batch source with the reference loader's shard / epoch / padding / shuffling rules."""

    def __init__(self, batch_size, file_root=None, file_list=None, files=None, labels=None, random_shuffle=False,
                 shuffle_after_epoch=False, initial_fill=1024, shard_id=0, num_shards=1, stick_to_shard=False, pad_last_batch=False,
                 seed=-1, shuffle_after_epoch_seed=None, file_filters=None, dir_filters=None, case_sensitive_filter=False):
        if not (0 <= shard_id < num_shards):
            raise ValueError("num_shards needs to be greater than shard_id")
        if random_shuffle and shuffle_after_epoch:
            raise ValueError("shuffle_after_epoch and random_shuffle cannot be both true")
        if shuffle_after_epoch and stick_to_shard:
            raise ValueError("shuffle_after_epoch and stick_to_shard cannot be both true")
        self.entries = discover_files(file_root, file_list, files, labels, file_filters, dir_filters, case_sensitive_filter)
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
        # the operator seed: the user's, else the pipeline's seed table (Pipeline.build -> set_seed), else the data-loader constant
        self.seed = -1 if seed is None or seed < 0 else int(seed)
        self._order = None              # dalihSampleOrder handle, created at the first read (the seed may be assigned until then)
        self.last = None

    def set_seed(self, seed):
        if self._order is not None:
            raise RuntimeError("the reader has already started reading")
        self.seed = int(seed)

    def _order_handle(self):
        if self._order is None:
            h = C.c_void_p()
            seed = 524287 if self.seed < 0 else self.seed
            _hcheck(_host().dalihSampleOrderCreate(C.byref(h), C.c_int64(len(self.entries)), int(self.random_shuffle), int(self.initial_fill),
                                                   C.c_int64(seed), int(self.shard_id), int(self.num_shards), int(self.stick_to_shard),
                                                   int(bool(self.pad_last_batch)), int(self.shuffle_after_epoch),
                                                   C.c_int64(self.shuffle_after_epoch_seed)))
            self._order = h
        return self._order

    def __del__(self):
        try:
            if self._order is not None:
                _host().dalihSampleOrderDestroy(self._order)
                self._order = None
        except Exception:
            pass

    # ---- reference: reader_meta keys
    def meta(self):
        n = len(self.entries)
        return {"epoch_size": n, "epoch_size_padded": int(math.ceil(n / self.num_shards)) * self.num_shards,
                "number_of_shards": self.num_shards, "shard_id": self.shard_id, "pad_last_batch": self.pad_last_batch,
                "stick_to_shard": self.stick_to_shard}

    def _next_sample(self, first_in_batch):
        """Index (in discovery order) of the next returned sample: Loader<>::ReadOne, restated in host_random.cc."""
        idx = C.c_int64()
        _hcheck(_host().dalihSampleOrderNext(self._order_handle(), int(bool(first_in_batch)), C.byref(idx)))
        self.last = int(idx.value)
        return self.last

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
            try:
                self._pin_ring[k] = capi.pinned_empty(size) if dev is None else capi.pinned_empty(size, dev)
            except capi.DaliB200Error:
                if dev is None:
                    raise
                self._pin_ring[k] = capi.pinned_empty(size)         # still page-locked, allocated under the thread's current device
        return self._pin_ring[k]

    def enable_prefetch(self, ahead=2):
        """Read `ahead` batches ahead of the consumer on a background thread (the reference's loader fills its queue on a thread of
        its own, loader.h PrefetchWorker): file I/O then overlaps the GPU work instead of sitting on the thread that schedules the
        pipeline.  The sequence of batches is unchanged -- the reader's state is only advanced by that one thread.  With page-locked
        arenas the ring must hold the batches in flight in the pipeline, the `ahead` queued ones and the one being read."""
        if getattr(self, "_ahead", None) is None:
            self._ahead = max(1, int(ahead))            # the thread starts with the first batch request (the seed may still be assigned)

    def _start_prefetch(self):
        import queue
        import threading
        import weakref
        self._q = queue.Queue(maxsize=self._ahead)
        self._stop = False
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
        if getattr(self, "_ahead", None) is not None and getattr(self, "_q", None) is None:
            self._start_prefetch()
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


class _RandomSource:
    """Common part of the random number generators (rng_base.h OperatorWithRng): a master Philox state keyed by the operator seed whose
    sequence counter advances by the batch size after every iteration (:143-145)."""

    def __init__(self, batch_size, shape, seed, dtype):
        self.batch_size, self.shape, self.dtype = int(batch_size), shape, np.dtype(dtype)
        self.seed = -1 if seed is None or seed < 0 else int(seed)
        self._sequence = 0
        self._started = False

    def set_seed(self, seed):
        if self._started:
            raise RuntimeError("the generator has already produced numbers")
        self.seed = int(seed)

    def _begin(self):
        if not self._started:
            if self.seed < 0:
                self.seed = _clock_seed()               # the reference seeds an unseeded pipeline from the clock (pipeline.cc:303)
            self._started = True
        shp = tuple(int(v) for v in self.shape) if self.shape is not None else ()
        outs = [np.empty(shp, self.dtype) for _ in range(self.batch_size)]
        vols = (C.c_int64 * self.batch_size)(*[o.size for o in outs])
        ptrs = (C.c_void_p * self.batch_size)(*[o.ctypes.data for o in outs])
        return outs, vols, ptrs

    def _end(self):
        self._sequence += self.batch_size


class CoinFlip(_RandomSource):
    def __init__(self, batch_size, probability=0.5, shape=None, seed=-1, dtype=np.int32):
        super().__init__(batch_size, shape, seed, dtype)
        self.p = float(probability)

    def __call__(self, _iteration=None):
        from . import types
        outs, vols, ptrs = self._begin()
        prob = (C.c_float * self.batch_size)(*([self.p] * self.batch_size))
        _hcheck(_host().dalihRandomCoinFlip(C.c_int64(self.seed), C.c_uint64(self._sequence), self.batch_size, vols, prob,
                                            int(types.from_numpy_type(self.dtype)), ptrs))
        self._end()
        return outs


class Uniform(_RandomSource):
    def __init__(self, batch_size, range=(-1.0, 1.0), values=None, shape=None, seed=-1, dtype=np.float32):
        super().__init__(batch_size, shape, seed, dtype)
        self.values = None if values is None else np.ascontiguousarray(np.asarray(values, np.float32).ravel())
        self.lo, self.hi = float(range[0]), float(range[1])
        if self.values is None and not self.lo < self.hi:
            raise ValueError(f"Invalid range [{self.lo}, {self.hi}).")

    def __call__(self, _iteration=None):
        from . import types
        outs, vols, ptrs = self._begin()
        if self.values is not None:
            rng, vals, nvals = None, self.values.ctypes.data_as(C.POINTER(C.c_float)), int(self.values.size)
        else:
            rng, vals, nvals = (C.c_float * (2 * self.batch_size))(*([self.lo, self.hi] * self.batch_size)), None, 0
        _hcheck(_host().dalihRandomUniform(C.c_int64(self.seed), C.c_uint64(self._sequence), self.batch_size, vols, rng, vals,
                                           C.c_int64(nvals), int(types.from_numpy_type(self.dtype)), ptrs))
        self._end()
        return outs
