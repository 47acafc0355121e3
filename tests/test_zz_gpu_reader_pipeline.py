"""-m gpu: SURVEY.md 8f rank 2 end to end -- fn.readers.file (page-locked arenas filled by the read-ahead thread) ->
decoders.image(mixed) -> resize -> crop_mirror_normalize(coin_flip mirror) through DALIGenericIterator(reader_name=...), against the
oracle.  (A file of its own, collected after the operator tests.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402
from test_gpu_pipeline import MEAN, STD, _jpegs, bits  # noqa: E402


def test_f2_file_reader_to_training_tensors(tmp_path):
    """SURVEY.md 8f rank 2, end to end: fn.readers.file (two shards, one pipeline per shard like one per GPU) -> decoders.image(mixed)
    -> resize -> crop_mirror_normalize(coin_flip mirror) through DALIGenericIterator(reader_name=...), two epochs with shard rotation.
    Every delivered sample is matched, through its label and file id, with the oracle applied to that file."""
    import torch
    from dali_b200 import fn, types, pipeline_def
    from dali_b200.plugin.pytorch import DALIGenericIterator, LastBatchPolicy
    files = {}
    k = 0
    for cls, n in (("a", 4), ("b", 3), ("c", 3)):
        (tmp_path / cls).mkdir()
        for i in range(n):
            s = _jpegs(1, 120 + 8 * k, 160 + 16 * (k % 3), 200 + k)[0]
            (tmp_path / cls / f"{i}.jpg").write_bytes(s.tobytes())
            files[k] = (s, ord(cls) - ord("a"))
            k += 1
    mean, inv = po.cmn_norm_args(MEAN, STD)
    want = {}
    for fid, (s, lab) in files.items():
        r = po.resample(po.jpeg_decode(s.tobytes()), (64, 80))
        want[fid] = {m: po.cmn(r, (0, 0), (64, 80), bool(m), mean, inv, np.float16, "CHW") for m in (0, 1)}

    def make(shard):
        @pipeline_def(batch_size=3, num_threads=2, device_id=0, seed=7 + shard)
        def pipe():
            data, label = fn.readers.file(file_root=str(tmp_path), shard_id=shard, num_shards=2, pad_last_batch=True, name="Reader")
            mir = fn.random.coin_flip(probability=0.5)
            img = fn.decoders.image(data, device="mixed", output_type=types.RGB)
            img = fn.resize(img, resize_x=80, resize_y=64)
            out = fn.crop_mirror_normalize(img, dtype=types.FLOAT16, output_layout="CHW", mean=MEAN, std=STD, mirror=mir)
            return out, label, mir
        return pipe()
    it = DALIGenericIterator([make(0), make(1)], ["data", "label", "mirror"], reader_name="Reader", last_batch_policy=LastBatchPolicy.PARTIAL,
                             auto_reset=True)
    # shard 0 = files 0..4, shard 1 = files 5..9 (loader.h:92-94 start_index); the pipelines swap shards in the second epoch
    label_of = [files[f][1] for f in range(10)]
    for epoch in range(2):
        seen = [[], []]
        for batch in it:
            for g, d in enumerate(batch):
                data, lab, mir = d["data"].cpu().numpy(), d["label"].cpu().numpy().ravel(), d["mirror"].cpu().numpy().ravel()
                assert d["data"].is_cuda and d["data"].dtype == torch.float16
                for j in range(len(lab)):
                    shard = (g + epoch) % 2
                    fid = 5 * shard + len(seen[g])
                    assert lab[j] == label_of[fid], (epoch, g, j)
                    assert np.array_equal(bits(data[j]), bits(want[fid][int(mir[j])])), (epoch, g, fid)
                    seen[g].append(fid)
        assert [len(s) for s in seen] == [5, 5], (epoch, seen)
