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
    with pytest.raises(RuntimeError, match="anchors of this model are unknown"):      # no silent default
        weights.load_model_dir(d)
    cfg, got = weights.load_model_dir(d, anchors=[(32, 32), (128, 128), (256, 256)])
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


def _small_weights(nc, na, seed):
    return weights.random_init(1, nc, na, seed=seed, randomize_bn=True)


def test_reference_trainer_style_saved_model(tmp_path, monkeypatch):
    """A directory laid out the way the reference trainer's export is (train.py:213-221), written by an independent
    encoder (tests/tf_fixture.py: google.protobuf messages, LevelDB table with really-snappy-compressed blocks, two data
    shards): Keras names shifted by the first model built in the process (conv2d_72 ..), `layer_with_weights-N` not in
    creation order, anchors only as graph constants.  2 anchors x (5 + 4 classes) = 18 channels - the case that used to
    load silently as 3 anchors x 1 class."""
    import tf_fixture
    anchors = [(64, 384), (384, 64)]
    w = _small_weights(4, 2, seed=9)
    d = str(tmp_path / "saved_model")
    tf_fixture.write_reference_style_saved_model(d, w, anchors, [-1, 1, 512, 512])
    raw_names = set(tb.read_keras_variables(os.path.join(d, "variables", "variables")))
    assert "conv2d_72/kernel" in raw_names and "conv2d/kernel" not in raw_names and "conv2d_transpose_3/kernel" in raw_names
    assert tb.saved_model_anchors(os.path.join(d, "saved_model.pb")) == [(64.0, 384.0), (384.0, 64.0)]
    cfg, got = weights.load_model_dir(d)
    assert cfg["anchors"] == [[64.0, 384.0], [384.0, 64.0]] and cfg["number_classes"] == 4 and cfg["anchors_source"] == "saved_model.pb"
    assert cfg["img_size"] == [512, 512, 1]
    assert sorted(got) == sorted(w)
    for k in w:
        assert np.array_equal(got[k], w[k]), k
    # the table really is snappy-compressed with copy tags, and every tensor checksum that was written verifies
    tab = open(os.path.join(d, "variables", "variables.index"), "rb").read()
    assert tab[-8:] == (0xdb4775248b80fb57).to_bytes(8, "little")
    small = {k for k, v in tb.read_bundle(os.path.join(d, "variables", "variables")).items() if isinstance(v, np.ndarray) and v.nbytes <= 1 << 16}
    assert len(small) > 300
    tb.read_bundle(os.path.join(d, "variables", "variables"), keys=small, verify=True)

    # graph constants unreadable -> loud failure, unless the anchors are supplied some other way
    d2 = str(tmp_path / "saved_model_no_consts")
    tf_fixture.write_reference_style_saved_model(d2, w, anchors, [-1, 1, 512, 512], with_anchor_consts=False)
    monkeypatch.delenv("Y3_ANCHORS", raising=False)
    with pytest.raises(RuntimeError, match="anchors of this model are unknown"):
        weights.load_model_dir(d2)
    monkeypatch.setenv("Y3_ANCHORS", "64,384;384,64")
    cfg2, _ = weights.load_model_dir(d2)
    assert cfg2["number_classes"] == 4 and cfg2["anchors_source"] == "Y3_ANCHORS"
    monkeypatch.setenv("Y3_ANCHORS", "32,32;128,128;256,256;64,64;8,8")        # 18 channels / 5 anchors: refused
    with pytest.raises(RuntimeError, match="do not fit"):
        weights.load_model_dir(d2)


def test_keras_name_renumbering():
    found = {"conv2d_72/kernel": 1, "conv2d_73/kernel": 2, "conv2d_143/bias": 3, "batch_normalization_72/gamma": 4,
             "batch_normalization_80/beta": 5, "conv2d_transpose_2/kernel": 6, "conv2d_transpose_3/bias": 7, "feature_map_2/kernel": 8}
    out = tb.normalize_keras_names(found)
    assert out == {"conv2d/kernel": 1, "conv2d_1/kernel": 2, "conv2d_71/bias": 3, "batch_normalization/gamma": 4,
                   "batch_normalization_8/beta": 5, "conv2d_transpose/kernel": 6, "conv2d_transpose_1/bias": 7, "feature_map_2/kernel": 8}
    same = {"conv2d/kernel": 1, "conv2d_1/kernel": 2, "conv2d_transpose/kernel": 3}
    assert tb.normalize_keras_names(same) == same
