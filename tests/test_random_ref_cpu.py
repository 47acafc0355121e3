"""Host-side randomness in front of the hot path (dali_b200/host/host_random.cc): the Philox generator, the
random.coin_flip / random.uniform operators and the per-operator seed table must give the reference's numbers.

Pinned three ways: published known answers of Philox4x32-10, the reference's own classes compiled into oracle/_ref (when present), and
golden vectors generated from them (tests/golden/random_ref.npz, tests/golden/make_random_golden.py) for boxes without the reference.
"""
import ctypes as C
import os

import numpy as np
import pytest

from dali_b200 import fn, pipeline_def, readers, types
from oracle import pyoracle as po

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "random_ref.npz")
needs_ref = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref is not built")


def _philox(key, seq, off, n):
    out = (C.c_uint32 * n)()
    assert readers._host().dalihTestPhilox(C.c_uint64(key), C.c_uint64(seq), C.c_uint64(off), n, out) == 0
    return np.array(out, np.uint32)


def _flip(seed, iterations, batch, volume, p, dtype=np.int32):
    g = readers.CoinFlip(batch, p, None if volume == 1 else [volume], seed, dtype)
    return np.stack([np.stack([np.asarray(s).reshape(volume) for s in g()]) for _ in range(iterations)])


def _uniform(seed, iterations, batch, volume, rng=(-1.0, 1.0), values=None, dtype=np.float32):
    g = readers.Uniform(batch, rng, values, None if volume == 1 else [volume], seed, dtype)
    return np.stack([np.stack([np.asarray(s).reshape(volume) for s in g()]) for _ in range(iterations)])


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds: counter = 0, key = 0
    assert [hex(v) for v in _philox(0, 0, 0, 4)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # the stream continues with counter + 1; an offset addresses single 32-bit outputs (offset = 4 * counter + phase)
    a = _philox(0x0123456789ABCDEF, 77, 0, 64)
    for off in (1, 2, 3, 4, 5, 30, 61):
        assert np.array_equal(_philox(0x0123456789ABCDEF, 77, off, 3), a[off:off + 3])
    assert not np.array_equal(_philox(0x0123456789ABCDEF, 78, 0, 4), a[:4])


@needs_ref
def test_philox_equals_the_reference_class():
    rng = np.random.default_rng(1)
    for _ in range(50):
        key, seq = (int(v) for v in rng.integers(0, 2 ** 63, 2))
        off = int(rng.integers(0, 2 ** 40))
        assert np.array_equal(_philox(key, seq, off, 23), po.ref_philox(key, seq, off, 23))


CASES_FLIP = [(7, 3, 5, 1, 0.5, np.int32), (123456789012345, 2, 4, 37, 0.25, np.int32), (0, 2, 3, 300, 0.9, np.uint8),
              (42, 1, 2, 5, 1.0, np.int32), (42, 1, 2, 5, 0.0, np.int32)]
CASES_UNIFORM = [(8, 3, 5, 1, (-1.0, 1.0), None, np.float32), (9, 2, 3, 41, (2.0, 3.0), None, np.float32),
                 (10, 2, 3, 17, (-5.5, 20.25), None, np.int32), (11, 2, 3, 17, (0.0, 255.0), None, np.uint8),
                 (12, 2, 3, 9, (-1e3, 1e3), None, np.float64), (13, 2, 2, 9, (-40000.0, 40000.0), None, np.int16),
                 (14, 2, 2, 9, (-3.0e9, 3.0e9), None, np.int64),
                 (15, 2, 4, 33, (0, 1), [1, 5, 9], np.float32), (16, 1, 4, 33, (0, 1), [0.25, -7.5, 300.0, 2.5, 3.5], np.uint8)]


@needs_ref
def test_random_operators_equal_the_reference():
    for seed, it, b, vol, p, dt in CASES_FLIP:
        assert np.array_equal(_flip(seed, it, b, vol, p, dt), po.ref_random_coin_flip(seed, it, b, vol, p, dt)), (seed, p)
    for seed, it, b, vol, rng, vals, dt in CASES_UNIFORM:
        got, want = _uniform(seed, it, b, vol, rng, vals, dt), po.ref_random_uniform(seed, it, b, vol, rng, vals, dt)
        assert got.dtype == want.dtype and np.array_equal(got, want), (seed, rng, vals, dt)


def test_random_operators_equal_the_golden_vectors():
    g = np.load(GOLDEN)
    for k, (seed, it, b, vol, p, dt) in enumerate(CASES_FLIP):
        assert np.array_equal(_flip(seed, it, b, vol, p, dt), g[f"flip{k}"]), k
    for k, (seed, it, b, vol, rng, vals, dt) in enumerate(CASES_UNIFORM):
        assert np.array_equal(_uniform(seed, it, b, vol, rng, vals, dt), g[f"uniform{k}"]), k


def test_random_operator_statistics_and_ranges():
    f = _flip(3, 8, 64, 100, 0.25)
    assert set(np.unique(f)) == {0, 1} and abs(f.mean() - 0.25) < 0.01
    u = _uniform(4, 8, 64, 100, (2.0, 3.0))
    assert u.min() >= 2.0 and u.max() < 3.0 and abs(u.mean() - 2.5) < 0.01
    d = _uniform(5, 4, 64, 100, values=[1, 5, 9])
    assert set(np.unique(d)) == {1.0, 5.0, 9.0}
    # samples and iterations are distinct streams of one seed; another seed is another stream
    assert not np.array_equal(u[0, 0], u[0, 1]) and not np.array_equal(u[0], u[1])
    assert not np.array_equal(u, _uniform(6, 8, 64, 100, (2.0, 3.0)))


def test_seed_table_and_assignment_order():
    """Pipeline seed -> operator seeds: std::seed_seq{seed}.generate over 1024 slots (pipeline.cc:303-308), handed out in the
    inputs-first order from the outputs (pipeline.py:2449-2461), only to operators that take a seed and were not given one."""
    t = readers.seed_table(7)
    assert len(t) == 1024 and all(0 <= v < 2 ** 32 for v in t) and len(set(t)) > 1000
    assert t == readers.seed_table(7) and t != readers.seed_table(8)
    g = np.load(GOLDEN)
    assert np.array_equal(np.array(readers.seed_table(7)[:16], np.int64), g["seed_table_7"])

    made = {}

    @pipeline_def(batch_size=4, num_threads=1, device_id=None, seed=7)
    def pipe():
        a = fn.random.uniform(range=(0.0, 1.0))              # defined first, reached last
        b = fn.random.coin_flip(probability=0.5)
        c = fn.random.coin_flip(probability=0.5, seed=99)     # user seed: takes no table entry
        d = fn.random.uniform(range=(0.0, 1.0))              # unreachable: pruned, takes no table entry
        made.update(a=a, b=b, c=c, d=d)
        return b, c, a
    p = pipe()
    p.build()
    src = {k: v.source.source for k, v in made.items()}
    assert src["b"].seed == t[0] and src["a"].seed == t[1] and src["c"].seed == 99 and src["d"].seed == -1
    outs = p.run()
    want_b = _flip(t[0], 1, 4, 1, 0.5)[0].ravel()
    assert np.array_equal(np.array([np.asarray(outs[0].at(i)).ravel()[0] for i in range(4)]), want_b)
    want_a = _uniform(t[1], 1, 4, 1, (0.0, 1.0))[0].ravel()
    assert np.array_equal(np.array([np.asarray(outs[2].at(i)).ravel()[0] for i in range(4)], np.float32), want_a)


def test_unsupported_types_are_reported():
    with pytest.raises(RuntimeError, match="not supported"):
        readers.CoinFlip(2, 0.5, None, 1, np.float32)()
    with pytest.raises(ValueError, match="Invalid range"):
        readers.Uniform(2, (1.0, 1.0), None, None, 1)
    assert types.FLOAT is not None
