"""TensorFlow-free SavedModel variable-bundle reader (CPU only).

PARITY UNPINNED against TensorFlow itself (not installable, no TF-written fixture in the reference): pinned by the
CRC-32C / snappy known answers, the LevelDB table layout rules and a round trip through the module's own writer."""
import os
import struct

import numpy as np
import pytest

from yolo3_b200 import tf_bundle as tb, weights


def test_crc32c_known_answers_and_chunked_path():
    assert tb.crc32c(b"") == 0
    assert tb.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value
    assert tb.crc32c(b"\x00" * 32) == 0x8A9136AA                        # RFC 3720 B.4
    assert tb.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tb.crc32c(bytes(range(32))) == 0x46DD794E
    x = np.random.default_rng(0).integers(0, 256, 300_007, dtype=np.uint8).tobytes()
    serial = tb._raw_update(0xffffffff, x) ^ 0xffffffff
    assert tb.crc32c(x) == serial == tb.crc32c(x[100_000:], tb.crc32c(x[:100_000]))
    assert tb.mask_crc(0) == 0xa282ead8


def test_snappy_decoder():
    lit = bytes([11, 4 << 2]) + b"abcde"
    assert tb.snappy_decompress(lit + bytes([(1 << 2) | 1, 5]) + bytes([0]) + b"z") == b"abcdeabcdez"
    # copy with 2-byte offset and an overlapping run (offset 1 = RLE)
    assert tb.snappy_decompress(bytes([9, 0]) + b"x" + bytes([(7 << 2) | 2, 1, 0])) == b"x" * 9
    raw = bytes(range(200)) * 3
    assert tb.snappy_decompress(tb._snappy_literal(raw)) == raw
    with pytest.raises(tb.BundleError):
        tb.snappy_decompress(bytes([4, (3 << 2) | 2, 9, 0]))


@pytest.mark.parametrize("snappy", [False, True])
def test_table_round_trip_prefix_compression_and_blocks(tmp_path, snappy):
    rng = np.random.default_rng(1)
    items = {b"": b"header"}
    for i in range(700):
        items[("layer_with_weights-%d/kernel/.ATTRIBUTES/VARIABLE_VALUE" % i).encode()] = rng.bytes(int(rng.integers(0, 90)))
    p = str(tmp_path / "t.index")
    tb.write_table(p, items.items(), block_size=1024, snappy=snappy)
    got = tb.read_table(p)
    assert [k for k, _ in got] == sorted(items) and dict(got) == items
    raw = open(p, "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == tb.TABLE_MAGIC and len(raw) > 20 * 1024
    bad = bytearray(raw)
    bad[100] ^= 1
    open(p, "wb").write(bytes(bad))
    with pytest.raises(tb.BundleError):
        tb.read_table(p)


def test_bundle_round_trip_dtypes_and_strings(tmp_path):
    rng = np.random.default_rng(2)
    tensors = {"a/f32": rng.standard_normal((3, 3, 5, 7)).astype(np.float32), "b/i64": np.arange(6, dtype=np.int64).reshape(2, 3),
               "c/scalar": np.float32(3.5).reshape(()), "d/u8": rng.integers(0, 255, (4,), dtype=np.uint8),
               "e/big": rng.standard_normal(200_000).astype(np.float32)}
    prefix = str(tmp_path / "variables" / "variables")
    tb.write_bundle(prefix, tensors, {"_s": b"hello \x00 world"})
    got = tb.read_bundle(prefix, verify=True)
    assert got.pop("_s") == [b"hello \x00 world"]
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v)
    assert list(tb.read_bundle(prefix, keys={"b/i64"})) == ["b/i64"]
    # a flipped data byte is caught by the per-tensor checksum
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[10] ^= 0x40
    open(data, "wb").write(bytes(raw))
    with pytest.raises(tb.BundleError):
        tb.read_bundle(prefix, verify=True)


def test_saved_model_directory_round_trip(tmp_path):
    """the layout tf.saved_model.save gives a Keras model (train.py:221) -> load_model_dir -> the Keras variables"""
    w = weights.random_init(1, 2, 3, seed=4, randomize_bn=True)
    d = str(tmp_path / "saved_model")
    tb.write_saved_model_variables(d, w, input_shape=[-1, 1, 384, 512], checksum_limit=1 << 16)
    assert os.path.exists(os.path.join(d, "saved_model.pb")) and os.path.exists(os.path.join(d, "variables", "variables.index"))
    assert tb.signature_input_shape(os.path.join(d, "saved_model.pb")) == [-1, 1, 384, 512]
    cfg, got = weights.load_model_dir(d)
    assert cfg["img_size"] == [384, 512, 1] and cfg["number_classes"] == 2 and len(cfg["anchors"]) == 3
    assert sorted(got) == sorted(w)
    for k in w:
        assert got[k].dtype == np.float32 and np.array_equal(got[k], w[k]), k
    keys = [k.decode() for k, _ in tb.read_table(os.path.join(d, "variables", "variables.index")) if k]
    assert "layer_with_weights-0/kernel/.ATTRIBUTES/VARIABLE_VALUE" in keys and tb.OBJECT_GRAPH_KEY in keys
    # anchors are graph constants: a y3_config.json next to saved_model.pb overrides the defaults
    import json
    json.dump({"anchors": [[10, 10], [20, 20], [30, 30], [40, 40], [50, 50], [60, 60], [70, 70]]}, open(os.path.join(d, "y3_config.json"), "w"))
    with pytest.raises(RuntimeError, match="anchors"):     # 21 detection channels cannot belong to 7 anchors
        weights.load_model_dir(d)
    json.dump({"anchors": [[12, 12], [24, 24], [48, 48]], "img_size": [256, 256, 1]}, open(os.path.join(d, "y3_config.json"), "w"))
    cfg3, _ = weights.load_model_dir(d)
    assert cfg3["anchors"] == [[12.0, 12.0], [24.0, 24.0], [48.0, 48.0]] and cfg3["img_size"] == [256, 256, 1]
