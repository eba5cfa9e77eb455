"""mccnn_b200.checkpoint: TensorFlow V2 checkpoints (tensor bundles) read and written without TensorFlow.

TensorFlow is not installed here, so the format is pinned three ways: CRC32C against the RFC 3720 vectors, the writer's
bytes for a one-variable checkpoint against the bytes assembled by hand from the format description (LevelDB table +
BundleHeaderProto / BundleEntryProto wire encoding), and reader(writer(x)) == x over many variables, blocks and dtypes
with corruption detected."""
import os
import struct

import numpy as np
import pytest

from mccnn_b200 import checkpoint as ck


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert ck.crc32c(b"") == 0


@pytest.mark.parametrize("n", [1 << 16, (1 << 16) + 1, 300_007, 4096 * 97 + 4095])
def test_crc32c_chunked_path_equals_the_byte_loop(n):
    data = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8).tobytes()
    assert ck.crc32c(data) == ck._crc32c_scalar(data)


def test_crc32c_and_mask_equal_tensorboards_implementation():
    """tensorboard ships an independent pure-Python CRC32C + the LevelDB / TFRecord mask (compat/tensorflow_stub)."""
    tb = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    rng = np.random.default_rng(4)
    for n in (0, 1, 7, 1000, 70_001):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert ck.crc32c(data) == tb.crc32c(data)
        assert ck._mask(ck.crc32c(data)) == tb.masked_crc32c(data)


def test_snappy_decoder():
    # literal "abcd", then a copy of 8 bytes from offset 4 (overlapping its own output), then literal "xy"
    raw = bytes([14]) + bytes([3 << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4]) + bytes([1 << 2]) + b"xy"
    assert ck._snappy_uncompress(raw) == b"abcdabcdabcdxy"


def test_writer_bytes_of_a_one_variable_checkpoint(tmp_path):
    prefix = str(tmp_path / "model.ckpt")
    ck.write_tf_checkpoint(prefix, {"a": np.array([1.0], np.float32)})
    raw = np.array([1.0], np.float32).tobytes()
    assert open(prefix + ".data-00000-of-00001", "rb").read() == raw
    # BundleHeaderProto{num_shards: 1, version{producer: 1}}; BundleEntryProto{dtype: DT_FLOAT, shape{dim{size: 1}},
    # size: 4, crc32c: masked} -- offset 0 is the proto3 default and is not written
    header = bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    crc = ck._mask(ck.crc32c(raw))
    entry = bytes([0x08, 0x01, 0x12, 0x04, 0x12, 0x02, 0x08, 0x01, 0x28, 0x04, 0x35]) + struct.pack("<I", crc)
    block = (bytes([0, 0, len(header)]) + header +             # key "": shared 0, non-shared 0
             bytes([0, 1, len(entry)]) + b"a" + entry +         # key "a"
             struct.pack("<II", 0, 1))                          # one restart point at 0
    def framed(contents):
        return contents + b"\x00" + struct.pack("<I", ck._mask(ck.crc32c(contents + b"\x00")))
    meta = struct.pack("<II", 0, 1)                             # empty metaindex block
    handle = bytes([0, len(block)])                             # data block: offset 0
    index = bytes([0, 1, len(handle)]) + b"a" + handle + struct.pack("<II", 0, 1)
    off_meta = len(block) + 5
    off_index = off_meta + len(meta) + 5
    footer = bytes([off_meta, len(meta), off_index, len(index)])
    footer += bytes(40 - len(footer)) + bytes([0x57, 0xFB, 0x80, 0x8B, 0x24, 0x75, 0x47, 0xDB])
    assert open(prefix + ".index", "rb").read() == framed(block) + framed(meta) + framed(index) + footer


def _many():
    rng = np.random.default_rng(0)
    t = {}
    for i in range(120):       # shared prefixes, several data blocks
        t["Conv_%d_weights" % i] = rng.normal(size=(3, 8 * (1 + i % 5))).astype(np.float32)
        t["Conv_%d_weights/Adam" % i] = rng.normal(size=(3, 8 * (1 + i % 5))).astype(np.float32)
        t["BN_%d/moving_variance" % i] = rng.uniform(0.5, 2.0, (16,)).astype(np.float32)
    t["global_step"] = np.array(12345, np.int64)
    t["beta1_power"] = np.array(0.5, np.float32)
    t["empty"] = np.zeros((0, 4), np.float32)
    t["big"] = rng.normal(size=(300, 257)).astype(np.float32)         # > 64 KB: chunked CRC path
    t["halves"] = rng.normal(size=(5, 3)).astype(np.float16)
    t["doubles"] = rng.normal(size=(2, 2, 2)).astype(np.float64)
    t["flags"] = np.array([True, False, True])
    t["ints"] = np.arange(-5, 5, dtype=np.int32)
    return t


def test_round_trip(tmp_path):
    prefix = str(tmp_path / "sub" / "model.ckpt")
    want = _many()
    ck.write_tf_checkpoint(prefix, want)
    got = ck.read_tf_checkpoint(prefix)
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        assert np.array_equal(got[k], want[k]), k
    assert os.path.getsize(prefix + ".index") > 3 * 4096        # the index really spans several blocks


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "model.ckpt")
    ck.write_tf_checkpoint(prefix, _many())
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[len(data) // 2] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    with pytest.raises(ValueError, match="checksum mismatch"):
        ck.read_tf_checkpoint(prefix)
    ck.read_tf_checkpoint(prefix, verify=False)                 # readable when asked not to check
    ck.write_tf_checkpoint(prefix, _many())
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[100] ^= 0x01
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError, match="block checksum"):
        ck.read_tf_checkpoint(prefix)
    open(prefix + ".index", "wb").write(bytes(idx[:-1]) + b"\x00")
    with pytest.raises(ValueError, match="magic"):
        ck.read_tf_checkpoint(prefix)


def test_bfloat16_entries_are_widened(tmp_path):
    prefix = str(tmp_path / "model.ckpt")
    vals = np.array([1.0, -2.5, 0.15625], np.float32)
    bits = (vals.view(np.uint32) >> 16).astype(np.uint16)       # exactly representable in bfloat16
    raw = bits.tobytes()
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    entry = ck._encode_entry(ck._DT_BFLOAT16, (3,), 0, len(raw), ck._mask(ck.crc32c(raw)))
    header = bytes([0x08, 0x01, 0x1A, 0x02, 0x08, 0x01])
    ck._write_table(prefix + ".index", [(b"", header), (b"w", entry)])
    got = ck.read_tf_checkpoint(prefix)["w"]
    assert got.dtype == np.float32 and np.array_equal(got, vals)


def test_load_into_a_variable_store(tmp_path):
    import torch
    from mccnn_b200.variables import VariableStore
    prefix = str(tmp_path / "model.ckpt")
    saved = _many()
    ck.write_tf_checkpoint(prefix, saved)
    store = VariableStore(device="cpu", seed=1)
    w0 = store.get_variable("Conv_0_weights", [3, 8], "glorot")
    ptr, version = w0.data_ptr(), w0._version
    loaded, skipped = ck.load_tf_checkpoint(store, prefix)
    assert "Conv_0_weights" in loaded and "BN_7/moving_variance" in loaded
    assert "Conv_0_weights/Adam" in skipped and "global_step" in skipped and "beta1_power" in skipped
    assert "flags" in skipped and "ints" in skipped              # not floating point: not model weights
    # in place: same storage, version bumped (graphs and folded batch-norm caches stay valid)
    assert w0.data_ptr() == ptr and w0._version > version
    assert torch.equal(w0, torch.from_numpy(saved["Conv_0_weights"]))
    assert torch.equal(store.vars["BN_7/moving_variance"], torch.from_numpy(saved["BN_7/moving_variance"]))
    assert store.vars["halves"].dtype == torch.float32
    assert not any(k.endswith("/Adam") for k in store.vars)
    # strict: a model variable the checkpoint does not have
    store.get_variable("Conv_999_weights", [3, 8], "glorot")
    with pytest.raises(KeyError, match="Conv_999_weights"):
        ck.load_tf_checkpoint(store, prefix, strict=True)
    # a shape mismatch is an error, not a silent reinterpretation
    store2 = VariableStore(device="cpu", seed=1)
    store2.get_variable("Conv_0_weights", [3, 16], "glorot")
    with pytest.raises(ValueError, match="Conv_0_weights"):
        ck.load_tf_checkpoint(store2, prefix)
    # and back: what the store holds can be written for the reference's scripts
    out = str(tmp_path / "export.ckpt")
    ck.write_tf_checkpoint(out, store.state_dict())
    again = ck.read_tf_checkpoint(out)
    assert np.array_equal(again["Conv_3_weights"], saved["Conv_3_weights"])
