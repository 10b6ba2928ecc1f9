"""TensorFlow-free reader (and minimal writer) for the variable bundle of a TF2 SavedModel / checkpoint.

The reference loads its weights with `tf.saved_model.load(saved_model_filepath)` (inference.py:35,
inference_tiled.py:325); the directory is written by `tf.saved_model.save(keras_model, ...)` (train.py:221):

    <dir>/saved_model.pb                              SavedModel proto (signatures, object graph, functions)
    <dir>/variables/variables.index                   LevelDB-style SSTable: key -> BundleEntryProto
    <dir>/variables/variables.data-00000-of-00001     raw little-endian tensor bytes

TensorFlow is not installable here, so this module restates the published on-disk formats:
  * table format (tensorflow/core/lib/io/table*.cc = LevelDB's): blocks of prefix-compressed entries
    [varint shared][varint non_shared][varint value_len][key delta][value], a restart array, a 1-byte
    compression type (0 none, 1 snappy) and a masked CRC32C per block; 48-byte footer with the metaindex and
    index block handles and the magic 0xdb4775248b80fb57;
  * tensor bundle (tensorflow/core/util/tensor_bundle/, protobuf/tensor_bundle.proto): key "" holds
    BundleHeaderProto, every other key a BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6};
  * object graph (protobuf/trackable_object_graph.proto) under the key `_CHECKPOINTABLE_OBJECT_GRAPH`: nodes
    with SerializedTensor attributes {name=1, full_name=2, checkpoint_key=3}; `full_name` is the Keras variable
    name ("conv2d_17/kernel"), `checkpoint_key` the bundle key
    ("layer_with_weights-34/kernel/.ATTRIBUTES/VARIABLE_VALUE").

PARITY UNPINNED: no TensorFlow-written file exists in this sandbox or in the reference repository, so the
reader is validated against this module's own writer, the CRC32C / snappy known-answer vectors and the
format documentation only (tests/test_tf_bundle.py).
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
OBJECT_GRAPH_KEY = "_CHECKPOINTABLE_OBJECT_GRAPH"
VALUE_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_STRING = 7
_DT_OF = {np.dtype(v): k for k, v in _DTYPES.items()}


class BundleError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------ primitives
def _varint(buf, pos):
    out = shift = 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7f) << shift
        if b < 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise BundleError("varint too long")


def _put_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _crc_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82f63b78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC = _crc_table()


_CRC_NP = np.array(_CRC, dtype=np.uint32)
_CHUNK = 4096
_ZERO_FEED = None          # 4 x 256 table of the linear map "feed _CHUNK zero bytes"


def _raw_update(c, data):
    for b in data:
        c = _CRC[(c ^ b) & 0xff] ^ (c >> 8)
    return c


def _zero_feed_tables():
    global _ZERO_FEED
    if _ZERO_FEED is None:
        basis = (np.uint32(1) << np.arange(32, dtype=np.uint32)).astype(np.uint32)
        for _ in range(_CHUNK):
            basis = _CRC_NP[basis & np.uint32(0xff)] ^ (basis >> np.uint32(8))
        tabs = np.zeros((4, 256), np.uint32)
        for j in range(4):
            for v in range(256):
                acc = 0
                for bit in range(8):
                    if v >> bit & 1:
                        acc ^= int(basis[8 * j + bit])
                tabs[j, v] = acc
        _ZERO_FEED = [[int(x) for x in tabs[j]] for j in range(4)]
    return _ZERO_FEED


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli), the checksum of the table blocks and of every tensor.  Large buffers are processed
    as 4 KB chunks in parallel with NumPy (the register update is linear over GF(2)) and combined serially."""
    data = bytes(data) if not isinstance(data, (bytes, bytearray, memoryview, np.ndarray)) else data
    mv = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    c = crc ^ 0xffffffff
    n_chunks = mv.shape[0] // _CHUNK
    if n_chunks >= 16:
        body = mv[:n_chunks * _CHUNK].reshape(n_chunks, _CHUNK)
        r = np.zeros(n_chunks, np.uint32)
        for i in range(_CHUNK):
            r = _CRC_NP[(r ^ body[:, i]) & np.uint32(0xff)] ^ (r >> np.uint32(8))
        z = _zero_feed_tables()
        for ri in r.tolist():
            c = z[0][c & 0xff] ^ z[1][(c >> 8) & 0xff] ^ z[2][(c >> 16) & 0xff] ^ z[3][c >> 24] ^ ri
        mv = mv[n_chunks * _CHUNK:]
    c = _raw_update(c, mv.tolist())
    return c ^ 0xffffffff


def mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xffffffff


def snappy_decompress(buf):
    """Raw snappy block format (preamble varint = uncompressed length; literal / copy-1 / copy-2 / copy-4 tags)."""
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise BundleError("corrupt snappy stream")
        for _ in range(ln):                      # overlapping copies are byte-serial by definition
            out.append(out[-off])
    if len(out) != n:
        raise BundleError("snappy length mismatch: %d != %d" % (len(out), n))
    return bytes(out)


# ------------------------------------------------------------------------------------------ protobuf (wire level)
def proto_fields(buf):
    """[(field_number, wire_type, value)]: varint -> int, 64-bit/32-bit -> bytes, length-delimited -> bytes."""
    out, pos = [], 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise BundleError("unsupported protobuf wire type %d" % wt)
        out.append((fn, wt, v))
    return out


def _pb(fn, wt, payload):
    return _put_varint((fn << 3) | wt) + payload


def _pb_bytes(fn, b):
    return _pb(fn, 2, _put_varint(len(b)) + b)


def _pb_int(fn, v):
    return _pb(fn, 0, _put_varint(v))


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_shape(buf):
    dims = []
    for fn, _, v in proto_fields(buf):
        if fn == 2:
            size = 0
            for f2, _, v2 in proto_fields(v):
                if f2 == 1:
                    size = _signed64(v2)
            dims.append(size)
    return tuple(dims)


def _parse_entry(buf):
    e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
    for fn, wt, v in proto_fields(buf):
        if fn == 1:
            e["dtype"] = v
        elif fn == 2:
            e["shape"] = _parse_shape(v)
        elif fn == 3:
            e["shard_id"] = v
        elif fn == 4:
            e["offset"] = v
        elif fn == 5:
            e["size"] = v
        elif fn == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif fn == 7:
            e["sliced"] = True
    return e


# ------------------------------------------------------------------------------------------ table
def _read_block(data, offset, size, verify):
    raw = data[offset:offset + size]
    trailer = data[offset + size:offset + size + 5]
    if len(raw) != size or len(trailer) != 5:
        raise BundleError("table block outside the file")
    if verify:
        want = struct.unpack("<I", trailer[1:5])[0]
        if mask_crc(crc32c(raw + trailer[:1])) != want:
            raise BundleError("table block checksum mismatch at offset %d" % offset)
    if trailer[0] == 1:
        raw = snappy_decompress(raw)
    elif trailer[0] != 0:
        raise BundleError("unknown block compression type %d" % trailer[0])
    n_restarts = struct.unpack("<I", raw[-4:])[0]
    end = len(raw) - 4 * (n_restarts + 1)
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(raw, pos)
        non_shared, pos = _varint(raw, pos)
        vlen, pos = _varint(raw, pos)
        key = key[:shared] + raw[pos:pos + non_shared]
        pos += non_shared
        out.append((key, raw[pos:pos + vlen]))
        pos += vlen
    return out


def read_table(path, verify=True):
    """All (key, value) pairs of an SSTable file, in key order."""
    with open(path, "rb") as fh:
        data = fh.read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != TABLE_MAGIC:
        raise BundleError("%s is not a TensorFlow/LevelDB table (bad magic)" % path)
    footer = data[-48:]
    _, pos = _varint(footer, 0)
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    out = []
    for _, handle in _read_block(data, idx_off, idx_size, verify):
        off, p = _varint(handle, 0)
        size, _ = _varint(handle, p)
        out += _read_block(data, off, size, verify)
    return out


class _BlockBuilder:
    def __init__(self, restart_interval):
        self.ri, self.buf, self.restarts, self.n, self.last = restart_interval, bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.n % self.ri == 0:
            if self.n:
                self.restarts.append(len(self.buf))
        else:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.n = key, self.n + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _snappy_literal(raw):
    """a valid (if pointless) snappy stream: literals only - exercises the reader's decompression path"""
    out = bytearray(_put_varint(len(raw)))
    for i in range(0, len(raw), 60):
        piece = raw[i:i + 60]
        out.append((len(piece) - 1) << 2)
        out += piece
    return bytes(out)


def write_table(path, items, block_size=4096, snappy=False):
    """items: iterable of (key bytes, value bytes); written sorted in the LevelDB table layout."""
    items = sorted(items)
    out = bytearray()
    index = _BlockBuilder(1)

    def emit(block):
        kind = b"\x01" if snappy else b"\x00"
        if snappy:
            block = _snappy_literal(block)
        off = len(out)
        out.extend(block)
        out.extend(kind + struct.pack("<I", mask_crc(crc32c(block + kind))))
        return _put_varint(off) + _put_varint(len(block))

    cur, last_key = _BlockBuilder(16), None
    for k, v in items:
        cur.add(k, v)
        last_key = k
        if len(cur.buf) >= block_size:
            index.add(last_key, emit(cur.finish()))
            cur = _BlockBuilder(16)
    if cur.n:
        index.add(last_key, emit(cur.finish()))
    meta = emit(_BlockBuilder(1).finish())
    idx = emit(index.finish())
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as fh:
        fh.write(bytes(out))


# ------------------------------------------------------------------------------------------ bundle
def read_bundle(prefix, keys=None, verify=False):
    """{key: ndarray} for the tensor bundle `<prefix>.index` + `<prefix>.data-XXXXX-of-YYYYY`.
    String tensors come back as lists of bytes.  keys: optional collection to restrict the read."""
    entries = read_table(prefix + ".index", verify=True)
    num_shards = 1
    meta = {}
    for k, v in entries:
        if k == b"":
            for fn, _, val in proto_fields(v):
                if fn == 1:
                    num_shards = val
                elif fn == 2 and val != 0:
                    raise BundleError("big-endian bundles are not supported")
            continue
        meta[k.decode("utf-8")] = _parse_entry(v)
    shards = {}

    def shard(i):
        if i not in shards:
            shards[i] = np.memmap("%s.data-%05d-of-%05d" % (prefix, i, num_shards), dtype=np.uint8, mode="r")
        return shards[i]

    out = {}
    for name, e in meta.items():
        if keys is not None and name not in keys:
            continue
        if e["sliced"]:
            raise BundleError("partitioned variable %s is not supported" % name)
        raw = shard(e["shard_id"])[e["offset"]:e["offset"] + e["size"]]
        if raw.shape[0] != e["size"]:
            raise BundleError("tensor %s lies outside its data shard" % name)
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if e["dtype"] == _DT_STRING:
            buf = bytes(raw)
            lens, pos = [], 0
            for _ in range(count):
                ln, pos = _varint(buf, pos)
                lens.append(ln)
            pos += 4                                         # masked crc32c of the length prefix
            vals = []
            for ln in lens:
                vals.append(buf[pos:pos + ln])
                pos += ln
            out[name] = vals
            continue
        if e["dtype"] not in _DTYPES:
            raise BundleError("tensor %s has unsupported dtype enum %d" % (name, e["dtype"]))
        dt = np.dtype(_DTYPES[e["dtype"]])
        if count * dt.itemsize != e["size"]:
            raise BundleError("tensor %s: %d bytes on disk, shape %s needs %d" % (name, e["size"], e["shape"], count * dt.itemsize))
        if verify and e["crc32c"] is not None and mask_crc(crc32c(raw)) != e["crc32c"]:
            raise BundleError("tensor %s checksum mismatch" % name)
        out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    return out


def object_graph_names(graph_bytes):
    """{checkpoint_key: full_name} from a serialized TrackableObjectGraph."""
    out = {}
    for fn, _, node in proto_fields(graph_bytes):
        if fn != 1:
            continue
        for f2, _, attr in proto_fields(node):
            if f2 != 2:
                continue
            full = key = None
            for f3, _, v in proto_fields(attr):
                if f3 == 2:
                    full = v.decode("utf-8")
                elif f3 == 3:
                    key = v.decode("utf-8")
            if full and key:
                out[key] = full
    return out


def read_keras_variables(prefix, verify=False):
    """{keras variable name ("conv2d_3/kernel"): fp32 array} of a SavedModel / tf.train.Checkpoint bundle."""
    graph = read_bundle(prefix, keys={OBJECT_GRAPH_KEY}).get(OBJECT_GRAPH_KEY)
    if not graph:
        raise BundleError("%s.index has no %s entry (not a TF2 object-based checkpoint)" % (prefix, OBJECT_GRAPH_KEY))
    names = object_graph_names(graph[0])
    wanted = {k: v for k, v in names.items() if k.endswith(VALUE_SUFFIX)}
    tensors = read_bundle(prefix, keys=set(wanted), verify=verify)
    out = {}
    for key, arr in tensors.items():
        full = wanted[key]
        full = full[:-2] if full.endswith(":0") else full
        # a model saved inside a name scope or as a sub-model carries a prefix: keep the last two path components
        parts = full.split("/")
        if len(parts) >= 2 and isinstance(arr, np.ndarray):
            out["/".join(parts[-2:])] = arr
    return out


def signature_input_shape(saved_model_pb):
    """Input shape (list, -1 for unknown) of the first signature of a saved_model.pb, or None.
    SavedModel{meta_graphs=2}; MetaGraphDef{signature_def=5 map<string,SignatureDef>};
    SignatureDef{inputs=1 map<string,TensorInfo>}; TensorInfo{tensor_shape=3}."""
    try:
        with open(saved_model_pb, "rb") as fh:
            buf = fh.read()
        best = None
        for fn, wt, mg in proto_fields(buf):
            if fn != 2 or wt != 2:
                continue
            for f2, w2, sig_entry in proto_fields(mg):
                if f2 != 5 or w2 != 2:
                    continue
                sig_name, sig = None, None
                for f3, _, v in proto_fields(sig_entry):
                    if f3 == 1:
                        sig_name = v
                    elif f3 == 2:
                        sig = v
                if sig is None or sig_name == b"__saved_model_init_op":
                    continue
                for f4, w4, inp in proto_fields(sig):
                    if f4 != 1 or w4 != 2:
                        continue
                    for f5, _, v in proto_fields(inp):
                        if f5 == 2:
                            for f6, _, ti in proto_fields(v):
                                if f6 == 3:
                                    shape = list(_parse_shape(ti))
                                    if len(shape) == 4 and (best is None or sig_name == b"serving_default"):
                                        best = shape
        return best
    except (OSError, BundleError, IndexError, struct.error):
        return None


_AUTO_NAME = None


def normalize_keras_names(found):
    """Keras auto-names count layers per PROCESS, not per model: the reference trainer builds a second YoloV3 in the
    same process before tf.saved_model.save (train.py:213-221), so its export holds conv2d_72 .. conv2d_143,
    batch_normalization_72 .., conv2d_transpose_2/_3 instead of conv2d .. conv2d_71.  Renumber every auto-named
    family relative to the smallest suffix present, so that names count from zero ("conv2d", "conv2d_1", ...).
    Explicitly named layers (feature_map_1..3) are left alone.  {name: array} -> {name: array}."""
    import re
    global _AUTO_NAME
    if _AUTO_NAME is None:
        _AUTO_NAME = re.compile(r"^(conv2d_transpose|conv2d|batch_normalization)(?:_(\d+))?$")
    lowest = {}
    for name in found:
        layer = name.split("/")[0]
        m = _AUTO_NAME.match(layer)
        if m:
            k = int(m.group(2)) if m.group(2) else 0
            lowest[m.group(1)] = min(lowest.get(m.group(1), k), k)
    out = {}
    for name, arr in found.items():
        layer, _, var = name.partition("/")
        m = _AUTO_NAME.match(layer)
        if m:
            k = (int(m.group(2)) if m.group(2) else 0) - lowest[m.group(1)]
            layer = m.group(1) if k == 0 else "%s_%d" % (m.group(1), k)
        key = layer + "/" + var
        if key in out:
            raise BundleError("two variables map to %s after renumbering the Keras auto-names" % key)
        out[key] = arr
    return out


def _tensor_proto(buf):
    """TensorProto{dtype=1, tensor_shape=2, tensor_content=4, float_val=5, double_val=6, int_val=7, int64_val=10}
    -> ndarray or None (only the numeric kinds an anchor table can have)."""
    dtype, shape, content, vals = 0, (), None, []
    for fn, wt, v in proto_fields(buf):
        if fn == 1:
            dtype = v
        elif fn == 2:
            shape = _parse_shape(v)
        elif fn == 4:
            content = bytes(v)
        elif fn in (5, 6, 7, 10):
            if wt == 2:                                   # packed
                if fn == 5:
                    vals += list(struct.unpack("<%df" % (len(v) // 4), v))
                elif fn == 6:
                    vals += list(struct.unpack("<%dd" % (len(v) // 8), v))
                else:
                    pos = 0
                    while pos < len(v):
                        x, pos = _varint(v, pos)
                        vals.append(_signed64(x))
            elif wt == 5:
                vals.append(struct.unpack("<f", v)[0])
            elif wt == 1:
                vals.append(struct.unpack("<d", v)[0])
            else:
                vals.append(_signed64(v))
    np_dt = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}.get(dtype)
    if np_dt is None:
        return None
    count = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if content is not None and len(content) == count * np.dtype(np_dt).itemsize:
        return np.frombuffer(content, dtype=np_dt).reshape(shape).astype(np.float64)
    if len(vals) == count:
        return np.asarray(vals, np.float64).reshape(shape)
    if len(vals) == 1:                                    # a splat constant
        return np.full(shape, vals[0], np.float64)
    return None


def _node_defs(buf):
    """[(name, op, [inputs], {attr: AttrValue bytes})] of the NodeDef messages in `buf` fields `field`."""
    name = op = None
    inputs, attrs = [], {}
    for fn, _, v in proto_fields(buf):
        if fn == 1:
            name = v.decode("utf-8", "replace")
        elif fn == 2:
            op = v.decode("utf-8", "replace")
        elif fn == 3:
            inputs.append(v.decode("utf-8", "replace"))
        elif fn == 5:
            k = val = None
            for f2, _, v2 in proto_fields(v):
                if f2 == 1:
                    k = v2.decode("utf-8", "replace")
                elif f2 == 2:
                    val = v2
            if k is not None:
                attrs[k] = val
    return name, op, inputs, attrs


def _anchor_tables(nodes):
    """anchor tables among the nodes of ONE graph / function body: the [A,2] constant operand of a Mul
    (model.py:163 `box_wh = tf.exp(box_wh) * self.anchors`), looked up through Cast / Identity."""
    by_name = {n[0]: n for n in nodes}

    def const_of(ref, depth=0):
        node = by_name.get(ref.lstrip("^").split(":")[0])
        if node is None or depth > 4:
            return None
        if node[1] == "Const" and node[3].get("value") is not None:
            for fn, _, v in proto_fields(node[3]["value"]):
                if fn == 8:
                    return _tensor_proto(v)
            return None
        if node[1] in ("Cast", "Identity") and node[2]:
            return const_of(node[2][0], depth + 1)
        return None

    out = []
    for _, op, inputs, _ in nodes:
        if op != "Mul" or len([i for i in inputs if not i.startswith("^")]) != 2:
            continue
        for ref in inputs:
            t = const_of(ref)
            if t is not None and t.ndim == 2 and t.shape[1] == 2 and 1 <= t.shape[0] <= 8 and np.all(np.isfinite(t)) and np.all(t > 0):
                out.append([(float(w), float(h)) for w, h in t])
    return out


def saved_model_anchors(saved_model_pb):
    """The anchor table [(w, h), ...] baked into the graph of a reference SavedModel, or None.
    The anchors are not variables: `YoloV3.reorg_layer` multiplies exp(t_wh) by the Python list `self.anchors`
    (model.py:163, 432-436), which TensorFlow embeds as a Const of shape [A, 2] feeding a Mul - once per scale, in the
    top-level graph or in a function of the library.  SavedModel{meta_graphs=2}; MetaGraphDef{graph_def=2};
    GraphDef{node=1, library=2}; FunctionDefLibrary{function=1}; FunctionDef{node_def=3}; NodeDef{name=1, op=2,
    input=3, attr=5}; AttrValue{tensor=8}.  All tables found must agree; disagreement or none -> None.
    PARITY UNPINNED (no TensorFlow-written file available): validated on hand-assembled protos only."""
    try:
        with open(saved_model_pb, "rb") as fh:
            buf = fh.read()
        tables = []
        for fn, wt, mg in proto_fields(buf):
            if fn != 2 or wt != 2:
                continue
            for f2, w2, gd in proto_fields(mg):
                if f2 != 2 or w2 != 2:
                    continue
                top = []
                for f3, w3, v in proto_fields(gd):
                    if f3 == 1 and w3 == 2:
                        top.append(_node_defs(v))
                    elif f3 == 2 and w3 == 2:
                        for f4, w4, fdef in proto_fields(v):
                            if f4 != 1 or w4 != 2:
                                continue
                            body = [_node_defs(nd) for f5, w5, nd in proto_fields(fdef) if f5 == 3 and w5 == 2]
                            tables += _anchor_tables(body)
                tables += _anchor_tables(top)
        if not tables or any(t != tables[0] for t in tables):
            return None
        return tables[0]
    except (OSError, BundleError, IndexError, struct.error, UnicodeDecodeError):
        return None


# ------------------------------------------------------------------------------------------ writer (tests, export)
def write_bundle(prefix, tensors, strings=None, checksum_limit=None):
    """tensors: {key: ndarray}; strings: {key: bytes} scalar string tensors.  One data shard.
    checksum_limit: tensors above this many bytes are written without their crc32c field (fast test fixtures;
    TensorFlow itself would reject such an entry) - None = checksum everything."""
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    items = [(b"", _pb_int(1, 1) + _pb_int(2, 0) + _pb_bytes(3, _pb_int(1, 1)))]   # num_shards, little endian, version{producer=1}
    data = bytearray()

    def shape_pb(shape):
        return b"".join(_pb_bytes(2, _pb_int(1, int(d))) for d in shape)

    for key in sorted(tensors):
        a = np.asarray(tensors[key])
        raw = a.tobytes(order="C")
        entry = _pb_int(1, _DT_OF[a.dtype]) + _pb_bytes(2, shape_pb(a.shape)) + _pb_int(4, len(data)) + _pb_int(5, len(raw))
        if checksum_limit is None or len(raw) <= checksum_limit:
            entry += _pb(6, 5, struct.pack("<I", mask_crc(crc32c(raw))))
        items.append((key.encode("utf-8"), entry))
        data += raw
    for key, val in sorted((strings or {}).items()):
        ln = _put_varint(len(val))
        raw = ln + struct.pack("<I", mask_crc(crc32c(ln))) + val
        entry = _pb_int(1, _DT_STRING) + _pb_bytes(2, b"") + _pb_int(4, len(data)) + _pb_int(5, len(raw)) \
            + _pb(6, 5, struct.pack("<I", mask_crc(crc32c(raw))))
        items.append((key.encode("utf-8"), entry))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as fh:
        fh.write(bytes(data))
    write_table(prefix + ".index", items)


def _anchor_function(anchors):
    """FunctionDef with the decode's `exp(t_wh) * anchors` nodes (what saved_model_anchors looks for)."""
    a = np.asarray(anchors, np.float32).reshape(-1, 2)
    shape = b"".join(_pb_bytes(2, _pb_int(1, int(d))) for d in a.shape)
    tensor = _pb_int(1, 1) + _pb_bytes(2, shape) + _pb_bytes(4, a.tobytes())

    def attr(key, val):
        return _pb_bytes(5, _pb_bytes(1, key) + _pb_bytes(2, val))

    def node(name, op, inputs, attrs=b""):
        return _pb_bytes(3, _pb_bytes(1, name) + _pb_bytes(2, op) + b"".join(_pb_bytes(3, i) for i in inputs) + attrs)

    body = node(b"Exp", b"Exp", [b"inputs"], attr(b"T", _pb_int(6, 1)))
    body += node(b"mul/y", b"Const", [], attr(b"dtype", _pb_int(6, 1)) + attr(b"value", _pb_bytes(8, tensor)))
    body += node(b"mul", b"Mul", [b"Exp:y:0", b"mul/y:output:0"], attr(b"T", _pb_int(6, 1)))
    return _pb_bytes(1, _pb_bytes(1, b"__inference_reorg_layer")) + body


def write_saved_model_variables(path, weights, input_shape=None, checksum_limit=None, anchors=None):
    """Writes `<path>/variables/variables.{index,data-*}` the way tf.saved_model.save lays a Keras model out
    (layer_with_weights-<i>/<attr>/.ATTRIBUTES/VARIABLE_VALUE keys + object graph) and, when input_shape
    ([-1, C, H, W]) is given, a minimal saved_model.pb that carries the serving signature's input shape and - when
    anchors are given - the decode's anchor constant as a library function (`exp(t_wh) * anchors`)."""
    layers = []
    for name in weights:
        layer = name.split("/")[0]
        if layer not in layers:
            layers.append(layer)
    tensors, nodes = {}, [b""]
    for i, layer in enumerate(layers):
        attrs = b""
        for name, arr in weights.items():
            if name.split("/")[0] != layer:
                continue
            var = name.split("/")[1]
            key = "layer_with_weights-%d/%s%s" % (i, var, VALUE_SUFFIX)
            tensors[key] = np.asarray(arr, np.float32)
            attrs += _pb_bytes(2, _pb_bytes(1, b"VARIABLE_VALUE") + _pb_bytes(2, name.encode()) + _pb_bytes(3, key.encode()))
        nodes.append(attrs)
    root = b"".join(_pb_bytes(1, _pb_int(1, i + 1) + _pb_bytes(2, ("layer_with_weights-%d" % i).encode())) for i in range(len(layers)))
    nodes[0] = root
    graph = b"".join(_pb_bytes(1, n) for n in nodes)
    write_bundle(os.path.join(path, "variables", "variables"), tensors, {OBJECT_GRAPH_KEY: graph}, checksum_limit)
    if input_shape is not None:
        shape = b"".join(_pb_bytes(2, _pb_int(1, int(d))) for d in input_shape)
        tinfo = _pb_bytes(1, b"serving_default_input_1:0") + _pb_int(2, 1) + _pb_bytes(3, shape)
        sig = _pb_bytes(1, _pb_bytes(1, b"input_1") + _pb_bytes(2, tinfo)) + _pb_bytes(3, b"tensorflow/serving/predict")
        mg = _pb_bytes(5, _pb_bytes(1, b"serving_default") + _pb_bytes(2, sig))
        if anchors is not None:                               # graph_def{library{function{...}}}
            mg = _pb_bytes(2, _pb_bytes(2, _pb_bytes(1, _anchor_function(anchors)))) + mg
        with open(os.path.join(path, "saved_model.pb"), "wb") as fh:
            fh.write(_pb_int(1, 1) + _pb_bytes(2, mg))
