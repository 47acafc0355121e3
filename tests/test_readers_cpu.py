"""CPU tests of the host-side input staging: fn.readers.file (discovery order, labels, sharding, shard
rotation, pad_last_batch, shuffling as a per-epoch permutation), fn.random.coin_flip / uniform, reader_meta and the
DALIGenericIterator(reader_name=...) epoch logic (dali/operators/reader/loader/loader.h, plugin/base_iterator.py)."""
import os

import numpy as np
import pytest

from dali_b200 import fn, pipeline_def
from dali_b200.plugin.pytorch import DALIGenericIterator, LastBatchPolicy
from dali_b200.readers import FileReader, discover_files, start_index


@pytest.fixture()
def tree(tmp_path):
    """10 files, classes created in non-sorted order; file k holds 4 bytes of value k (reference order: a_cat 0-3, b_dog 4-6, c_eel 7-9)."""
    k = 0
    for c in ("b_dog", "a_cat", "c_eel"):
        (tmp_path / c).mkdir()
    for c, n in (("a_cat", 4), ("b_dog", 3), ("c_eel", 3)):
        for i in range(n):
            (tmp_path / c / f"img{i}.jpg").write_bytes(bytes([k] * 4))
            k += 1
    (tmp_path / "a_cat" / "notes.txt").write_text("not an image")
    (tmp_path / "loose.jpg").write_bytes(b"\x00")
    return str(tmp_path)


def _ids(reader, nbatches):
    out = []
    for _ in range(nbatches):
        data, labels = reader(0)
        out.append([int(d[0]) for d in data])
        assert all(l.dtype == np.int32 and l.shape == (1,) for l in labels)
    return out


def test_discovery_order_labels_and_lists(tree, tmp_path):
    e = discover_files(tree)
    assert [os.path.relpath(p, tree) for p, _ in e[:5]] == ["a_cat/img0.jpg", "a_cat/img1.jpg", "a_cat/img2.jpg", "a_cat/img3.jpg",
                                                            "b_dog/img0.jpg"]
    assert [l for _, l in e] == [0] * 4 + [1] * 3 + [2] * 3              # label = index of the sorted class directory
    lst = tmp_path / "list.txt"
    lst.write_text("c_eel/img1.jpg 7\nname with space/x.jpg 3\n\na_cat/img0.jpg 0\n")
    assert discover_files(tree, str(lst)) == [(os.path.join(tree, "c_eel/img1.jpg"), 7), (os.path.join(tree, "name with space/x.jpg"), 3),
                                              (os.path.join(tree, "a_cat/img0.jpg"), 0)]
    assert discover_files(files=["/x/a.jpg", "/x/b.jpg"]) == [("/x/a.jpg", 0), ("/x/b.jpg", 1)]
    assert discover_files(files=["a.jpg"], labels=[5], file_root="/r") == [("/r/a.jpg", 5)]
    with pytest.raises(ValueError):
        discover_files(files=["a"], labels=[1, 2])
    # file_filters / dir_filters / case_sensitive_filter (file_reader_op.cc:128-138, discover_files.cc:40-110): globs, matched
    # case-insensitively by default; labels are the indices of the directories that PASSED the filter
    os.rename(os.path.join(tree, "b_dog", "img1.jpg"), os.path.join(tree, "b_dog", "IMG1.JPG"))
    rel = lambda es: [(os.path.relpath(p, tree), l) for p, l in es]
    assert ("b_dog/IMG1.JPG", 1) in rel(discover_files(tree))
    assert ("b_dog/IMG1.JPG", 1) not in rel(discover_files(tree, case_sensitive_filter=True))
    assert rel(discover_files(tree, file_filters=["*.txt"])) == [("a_cat/notes.txt", 0)]
    assert rel(discover_files(tree, file_filters=["img0.*", "*2.jpg"], dir_filters=["?_[cd]*"])) == [
        ("a_cat/img0.jpg", 0), ("a_cat/img2.jpg", 0), ("b_dog/img0.jpg", 1), ("b_dog/img2.jpg", 1)]
    assert rel(discover_files(tree, dir_filters=["C_*"])) == [("c_eel/img0.jpg", 0), ("c_eel/img1.jpg", 0), ("c_eel/img2.jpg", 0)]
    assert discover_files(tree, dir_filters=["C_*"], case_sensitive_filter=True) == []
    os.rename(os.path.join(tree, "b_dog", "IMG1.JPG"), os.path.join(tree, "b_dog", "img1.jpg"))


def test_sharding_rotation_and_padding(tree):
    assert [start_index(s, 3, 10) for s in range(4)] == [0, 3, 6, 10]      # loader.cc:78-81
    assert _ids(FileReader(2, tree, shard_id=1, num_shards=3, stick_to_shard=True, pad_last_batch=True), 4) == [[3, 4], [5, 5], [3, 4], [5, 5]]
    assert _ids(FileReader(2, tree, shard_id=1, num_shards=3, stick_to_shard=True), 4) == [[3, 4], [5, 3], [4, 5], [3, 4]]
    assert _ids(FileReader(2, tree, shard_id=0, num_shards=3), 6) == [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9], [0, 1]]     # shards rotate
    assert _ids(FileReader(2, tree, shard_id=0, num_shards=3, pad_last_batch=True), 7) == [[0, 1], [2, 2], [3, 4], [5, 5], [6, 7], [8, 9], [0, 1]]
    assert _ids(FileReader(4, tree, shard_id=2, num_shards=3, stick_to_shard=True, pad_last_batch=True), 2) == [[6, 7, 8, 9], [6, 7, 8, 9]]
    with pytest.raises(ValueError):
        FileReader(2, tree, shard_id=3, num_shards=3)
    with pytest.raises(RuntimeError, match="number of shards"):
        FileReader(2, tree, shard_id=0, num_shards=11)


def test_shuffling_is_a_permutation_per_epoch(tree):
    for kw in (dict(random_shuffle=True, initial_fill=4, seed=3), dict(shuffle_after_epoch=True, seed=3)):
        r = FileReader(5, tree, **kw)
        a = _ids(r, 6)
        epochs = [sorted(a[2 * e][:] + a[2 * e + 1][:]) for e in range(3)]
        assert epochs == [list(range(10))] * 3
        assert a[0] + a[1] != list(range(10)) and a[0] + a[1] != a[2] + a[3]
        b = _ids(FileReader(5, tree, **kw), 6)
        assert a == b                                                    # deterministic for a given seed
    # random_shuffle shuffles the file list once with a seed every rank shares (file_label_loader.h:200-205): the shards are ranges
    # [5 k, 5 k + 5) of THAT list, so they still partition the data set whatever the per-rank seeds are
    kw = dict(num_shards=2, stick_to_shard=True, random_shuffle=True, initial_fill=100, pad_last_batch=True)
    sh = [_ids(FileReader(3, tree, shard_id=k, seed=1 + k, **kw), 4) for k in range(2)]
    parts = [sorted(x[0] + x[1][:2]) for x in sh]
    assert sorted(parts[0] + parts[1]) == list(range(10)) and parts[0] != [0, 1, 2, 3, 4]
    for x in sh:
        assert x[1][2] == x[1][1]                                            # each shard once, then the pad
        assert sorted(x[2] + x[3][:2]) == sorted(x[0] + x[1][:2])            # stick_to_shard: the same files in the next epoch


def test_random_generators():
    @pipeline_def(batch_size=64, num_threads=1, device_id=None)
    def p():
        return (fn.random.coin_flip(probability=0.25, seed=7), fn.random.uniform(range=(2.0, 3.0), seed=8),
                fn.random.uniform(values=[1, 5, 9], shape=[3], seed=9), fn.random.coin_flip(seed=7, probability=0.25))
    pipe = p()
    pipe.build()
    c, u, v, c2 = pipe.run()
    cs = np.array([c.at(i) for i in range(64)])
    assert cs.dtype == np.int32 and cs.shape == (64,) and set(cs.tolist()) <= {0, 1} and 3 <= cs.sum() <= 32
    us = np.array([u.at(i) for i in range(64)])
    assert us.dtype == np.float32 and (us >= 2).all() and (us < 3).all() and us.std() > 0.1
    vs = np.array([v.at(i) for i in range(64)])
    assert vs.shape == (64, 3) and set(vs.ravel().tolist()) <= {1.0, 5.0, 9.0}
    assert np.array_equal(cs, np.array([c2.at(i) for i in range(64)]))   # same seed, same stream


def test_pipeline_reader_meta_and_iterator_epochs(tree):
    def make(policy, pad, shard_id=1, num_shards=3, bs=2, stick=True):
        @pipeline_def(batch_size=bs, num_threads=1, device_id=None, prefetch_queue_depth=2)
        def p():
            data, label = fn.readers.file(file_root=tree, shard_id=shard_id, num_shards=num_shards, stick_to_shard=stick, pad_last_batch=pad,
                                          name="Reader")
            return data, label
        pipe = p()
        return pipe, DALIGenericIterator(pipe, ["data", "label"], reader_name="Reader", last_batch_policy=policy, auto_reset=True)
    pipe, it = make(LastBatchPolicy.FILL, True)
    assert pipe.reader_meta("Reader") == {"epoch_size": 10, "epoch_size_padded": 12, "number_of_shards": 3, "shard_id": 1,
                                          "pad_last_batch": True, "stick_to_shard": True}
    assert pipe.epoch_size("Reader") == 12 and it.size == 4
    for epoch in range(3):                                               # shard 1 = files 3, 4, 5 (labels 0, 1, 1), padded to 4
        batches = [(b[0]["data"][:, 0].tolist(), b[0]["label"][:, 0].tolist()) for b in it]
        assert batches == [([3, 4], [0, 1]), ([5, 5], [1, 1])], epoch
    _, it = make(LastBatchPolicy.PARTIAL, True)
    assert [b[0]["data"][:, 0].tolist() for b in it] == [[3, 4], [5]]
    assert [b[0]["data"][:, 0].tolist() for b in it] == [[3, 4], [5]]     # next epoch: same shard again (stick_to_shard)
    _, it = make(LastBatchPolicy.DROP, True)
    assert [b[0]["data"][:, 0].tolist() for b in it] == [[3, 4]]
    assert [b[0]["data"][:, 0].tolist() for b in it] == [[3, 4]]          # the dropped [5, 5] batch does not leak into the next epoch
    _, it = make(LastBatchPolicy.FILL, False, shard_id=0, num_shards=1, bs=4, stick=False)
    assert it.size == 12                                                 # ceil(10 / 4) * 4: the last batch wraps around
    assert [b[0]["data"][:, 0].tolist() for b in it] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 0, 1]]
    with pytest.raises(ValueError, match="size should not be set"):
        DALIGenericIterator(pipe, ["data", "label"], size=10, reader_name="Reader")


def test_iterator_fill_wraparound_keeps_epochs_aligned(tree):
    """plugin/base_iterator.py reset(): with FILL and pad_last_batch=False the reader runs on into the next epoch; the
    read-ahead is carried into the next epoch's counter and `size` is re-evaluated, so the epochs do not drift."""
    @pipeline_def(batch_size=4, num_threads=1, device_id=None, prefetch_queue_depth=2)
    def p():
        data, label = fn.readers.file(file_root=tree, name="Reader")
        return data, label
    it = DALIGenericIterator(p(), ["data", "label"], reader_name="Reader", last_batch_policy=LastBatchPolicy.FILL, auto_reset=True)
    e1 = [b[0]["data"][:, 0].tolist() for b in it]
    assert e1 == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 0, 1]] and len(e1) == 3
    assert it.size == 8 and len(it) == 2                               # 2 samples of the next epoch were already consumed
    e2 = [b[0]["data"][:, 0].tolist() for b in it]
    assert e2 == [[2, 3, 4, 5], [6, 7, 8, 9]]
    assert it.size == 12
    e3 = [b[0]["data"][:, 0].tolist() for b in it]
    assert e3 == e1


def test_shuffle_after_epoch_is_rank_independent_and_sticks_to_shard(tree):
    """file_label_loader.h:61-63,134-138,229-239: one global permutation per epoch, identical on every rank whatever the
    operator seed; the shard index does not rotate and reader_meta says so."""
    r0 = FileReader(5, tree, shuffle_after_epoch=True, shard_id=0, num_shards=2, seed=11)
    r1 = FileReader(5, tree, shuffle_after_epoch=True, shard_id=1, num_shards=2, seed=12)
    assert r0.meta()["stick_to_shard"] is True and r1.meta()["stick_to_shard"] is True
    for epoch in range(3):
        a, b = _ids(r0, 1)[0], _ids(r1, 1)[0]
        assert sorted(a + b) == list(range(10)), epoch                   # the two shards partition the data set in every epoch
    other = FileReader(5, tree, shuffle_after_epoch=True, shard_id=0, num_shards=2, shuffle_after_epoch_seed=99)
    assert _ids(other, 1) != _ids(FileReader(5, tree, shuffle_after_epoch=True, shard_id=0, num_shards=2), 1)


def test_reader_arena_ring_keeps_batches_in_flight_intact(tree, monkeypatch):
    """GPU pipelines read the files of a batch into a ring of page-locked arenas (one per batch in flight) that the mixed decoder copies
    from by DMA; here the arena allocator is replaced by plain numpy memory to check the bookkeeping on the CPU: every sample is a view
    of the batch's arena with the file's bytes, and the arenas of the last `num_buffers - 1` batches are not overwritten."""
    from dali_b200 import capi
    monkeypatch.setattr(capi, "pinned_empty", lambda n: np.zeros(max(1, int(n)), np.uint8))
    r = FileReader(4, tree)
    r.enable_pinned(3)
    batches = [r() for _ in range(5)]
    want = [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 0, 1], [2, 3, 4, 5], [6, 7, 8, 9]]
    for b, (data, labels) in enumerate(batches):
        assert len({d.base is None for d in data}) == 1 and all(d.base is not None for d in data)     # views of one arena
    # the two most recent batches still hold their own bytes (ring of 3: the arena of batch 2 was reused by batch 5 - 3 = 2 ... only older ones)
    for b in (3, 4):
        assert [int(d[0]) for d in batches[b][0]] == want[b] and all(d.size == 4 for d in batches[b][0])
    # batch 1's arena was recycled for batch 4 (1 + 3): its views now show batch 4's data -- which is why the ring must be as deep as the
    # number of batches in flight + 1 (prefetch_queue_depth + 1 in fn.readers.file)
    assert [int(d[0]) for d in batches[1][0]] == want[4]


def test_reader_read_ahead_thread_with_arena_ring(tree, monkeypatch):
    """The read-ahead thread may run `ahead` batches (+ the one it is reading) in front of the consumer; with the ring sized as
    fn.readers.file does (depth + 1 + ahead + 1 arenas) a consumer that keeps `depth` batches in flight never sees an arena overwritten,
    and the sequence of batches is the one of the synchronous reader."""
    import time
    from dali_b200 import capi
    monkeypatch.setattr(capi, "pinned_empty", lambda n: np.zeros(max(1, int(n)), np.uint8))
    depth, ahead = 2, 2
    ref = FileReader(3, tree, random_shuffle=True, initial_fill=5, seed=4)
    want = [[int(d[0]) for d in ref()[0]] for _ in range(12)]
    r = FileReader(3, tree, random_shuffle=True, initial_fill=5, seed=4)
    r.enable_pinned(depth + 1 + ahead + 1)
    r.enable_prefetch(ahead)
    inflight = []
    for k in range(12):
        inflight.append((k, r()[0]))
        time.sleep(0.01)                                         # let the producer run as far ahead as it is allowed to
        for kk, data in inflight[-depth:]:                       # the batches a pipeline of this depth still reads from
            assert [int(d[0]) for d in data] == want[kk], (k, kk)
    r.close()


def test_reader_read_ahead_reports_io_errors_on_the_consumer(tree):
    import os
    r = FileReader(4, tree)
    os.remove(os.path.join(tree, "b_dog", "img1.jpg"))            # file 5 disappears after discovery
    r.enable_prefetch(2)
    assert [int(d[0]) for d in r()[0]] == [0, 1, 2, 3]
    for _ in range(2):                                            # the error is raised where the batch is consumed, and stays
        with pytest.raises((FileNotFoundError, OSError)):
            r()


def test_reader_read_ahead_thread_ends_with_the_reader(tmp_path):
    """The read-ahead thread holds the reader only while it reads a batch: dropping the reader (with its pipeline) ends the thread."""
    import gc
    import time
    from dali_b200.readers import FileReader
    (tmp_path / "a").mkdir()
    for i in range(4):
        (tmp_path / "a" / f"{i}.jpg").write_bytes(bytes([i]) * 10)
    r = FileReader(2, file_root=str(tmp_path))
    r.enable_prefetch(2)
    r()
    thr = r._thr
    assert thr.is_alive()
    del r
    gc.collect()
    t0 = time.time()
    while thr.is_alive() and time.time() - t0 < 5:
        time.sleep(0.05)
    assert not thr.is_alive()


def test_shuffle_orders_are_the_standard_library_shuffles_the_reference_calls(tree, tmp_path):
    """random_shuffle = std::shuffle(entries, std::mt19937(524287)) once (file_label_loader.h:200-205); with initial_fill = 1 the
    buffer holds one sample, so the reader returns exactly that order.  shuffle_after_epoch = std::shuffle with
    std::mt19937_64(seed + (epoch << 32)), epoch counted from 1, applied to the order the previous epoch left (:229-240)."""
    import shutil
    import subprocess
    cxx = shutil.which(os.environ.get("CXX", "g++"))
    if cxx is None:
        pytest.skip("no C++ compiler")
    src = tmp_path / "s.cc"
    src.write_text(r"""
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>
int main() {
  std::vector<int> v(10);
  std::iota(v.begin(), v.end(), 0);
  std::mt19937 g(524287);
  std::shuffle(v.begin(), v.end(), g);
  for (int x : v) std::printf("%d ", x);
  std::printf("\n");
  std::iota(v.begin(), v.end(), 0);
  for (unsigned long long epoch = 1; epoch <= 3; epoch++) {
    std::mt19937_64 g64(77ull + (epoch << 32));
    std::shuffle(v.begin(), v.end(), g64);
    for (int x : v) std::printf("%d ", x);
    std::printf("\n");
  }
}
""")
    subprocess.run([cxx, "-std=c++17", "-O1", str(src), "-o", str(tmp_path / "s")], check=True)
    lines = [[int(t) for t in ln.split()] for ln in subprocess.run([str(tmp_path / "s")], check=True, capture_output=True, text=True).stdout.splitlines()]
    got = _ids(FileReader(5, tree, random_shuffle=True, initial_fill=1, seed=5), 4)
    assert got[0] + got[1] == lines[0] and got[2] + got[3] == lines[0]           # the same global order in every epoch
    got = _ids(FileReader(5, tree, shuffle_after_epoch=True, shuffle_after_epoch_seed=77), 6)
    assert [got[0] + got[1], got[2] + got[3], got[4] + got[5]] == lines[1:4]


def test_legacy_ops_api_define_graph_and_iter_setup(tree):
    """nvidia.dali.ops object API + Pipeline subclass with define_graph() / iter_setup() (the style of the reference's older examples):
    the classes forward to the same functional wrappers, so the graph -- and the numbers for a given seed -- are those of fn.*."""
    from dali_b200 import Pipeline, ops
    import nvidia.dali.ops as nv_ops
    assert nv_ops is ops and ops.decoders.Image is not None and ops.ImageDecoder._fn_name == "decoders.image"
    assert {"Resize", "CropMirrorNormalize", "WarpAffine", "Hsv", "Spectrogram", "MelFilterBank", "FileReader", "ExternalSource"} <= set(dir(ops))

    class Legacy(Pipeline):
        def __init__(self):
            super().__init__(batch_size=4, num_threads=1, device_id=None, seed=21)
            self.reader = ops.readers.File(file_root=tree, name="Reader")
            self.flip = ops.random.CoinFlip(probability=0.3)
            self.src = ops.ExternalSource()
            self.fed = 0

        def define_graph(self):
            data, label = self.reader()
            self.extra = self.src()
            return data, label, self.flip(), self.extra

        def iter_setup(self):
            self.feed_input(self.extra, [np.full((2,), self.fed, np.int32)] * 4)
            self.fed += 1

    @pipeline_def(batch_size=4, num_threads=1, device_id=None, seed=21)
    def functional():
        data, label = fn.readers.file(file_root=tree, name="Reader")
        return data, label, fn.random.coin_flip(probability=0.3)
    a, b = Legacy(), functional()
    a.build()
    assert a.reader_meta("Reader")["epoch_size"] == 10
    for it in range(3):
        oa, ob = a.run(), b.run()
        for k in range(3):
            for i in range(4):
                assert np.array_equal(np.asarray(oa[k].at(i)), np.asarray(ob[k].at(i))), (it, k, i)
        assert all(int(np.asarray(oa[3].at(i))[0]) == it for i in range(4))
