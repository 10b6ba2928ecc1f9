"""Independent writer of a TF2 SavedModel directory for the reader tests (test infrastructure).

Nothing here uses yolo3_b200.tf_bundle: the protocol buffers are encoded by the official `google.protobuf` runtime
from message descriptors declared below with the field numbers of TensorFlow's published .proto files
(tensor_bundle.proto, trackable_object_graph.proto, saved_model.proto, meta_graph.proto, graph.proto, function.proto,
node_def.proto, attr_value.proto, tensor.proto, tensor_shape.proto), the table file is laid out per LevelDB's
table_format.md with blocks compressed by a real snappy encoder (pyarrow's: copy tags, not just literals), the data is
split over two shards, CRC-32C is computed by a plain bitwise loop, and the Keras names are what the reference trainer
produces (train.py:213-221 builds the model twice in one process, so the exported layers are conv2d_72 .. and
`layer_with_weights-N` follows the functional model's depth order, not creation order).
"""
import os
import struct

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_T = {"int32": _F.TYPE_INT32, "int64": _F.TYPE_INT64, "string": _F.TYPE_STRING, "bytes": _F.TYPE_BYTES, "bool": _F.TYPE_BOOL,
      "fixed32": _F.TYPE_FIXED32, "float": _F.TYPE_FLOAT, "enum": _F.TYPE_INT32}


def _build_pool():
    fd = descriptor_pb2.FileDescriptorProto(name="tf_subset.proto", package="tfs", syntax="proto3")

    def msg(name, *fields, maps=()):
        m = fd.message_type.add(name=name)
        for fname, num, typ, rep in fields:
            f = m.field.add(name=fname, number=num, label=_F.LABEL_REPEATED if rep else _F.LABEL_OPTIONAL)
            if typ in _T:
                f.type = _T[typ]
            else:
                f.type, f.type_name = _F.TYPE_MESSAGE, ".tfs." + typ
        for fname, num, vtyp in maps:
            e = m.nested_type.add(name=fname.title().replace("_", "") + "Entry")
            e.options.map_entry = True
            e.field.add(name="key", number=1, type=_F.TYPE_STRING, label=_F.LABEL_OPTIONAL)
            v = e.field.add(name="value", number=2, label=_F.LABEL_OPTIONAL)
            v.type, v.type_name = _F.TYPE_MESSAGE, ".tfs." + vtyp
            f = m.field.add(name=fname, number=num, label=_F.LABEL_REPEATED, type=_F.TYPE_MESSAGE)
            f.type_name = ".tfs.%s.%s" % (name, e.name)
        return m

    msg("Dim", ("size", 1, "int64", 0), ("name", 2, "string", 0))
    msg("TensorShapeProto", ("dim", 2, "Dim", 1), ("unknown_rank", 3, "bool", 0))
    msg("VersionDef", ("producer", 1, "int32", 0), ("min_consumer", 2, "int32", 0))
    msg("BundleHeaderProto", ("num_shards", 1, "int32", 0), ("endianness", 2, "enum", 0), ("version", 3, "VersionDef", 0))
    msg("BundleEntryProto", ("dtype", 1, "enum", 0), ("shape", 2, "TensorShapeProto", 0), ("shard_id", 3, "int32", 0),
        ("offset", 4, "int64", 0), ("size", 5, "int64", 0), ("crc32c", 6, "fixed32", 0))
    msg("ObjectReference", ("node_id", 1, "int32", 0), ("local_name", 2, "string", 0))
    msg("SerializedTensor", ("name", 1, "string", 0), ("full_name", 2, "string", 0), ("checkpoint_key", 3, "string", 0))
    msg("TrackableObject", ("children", 1, "ObjectReference", 1), ("attributes", 2, "SerializedTensor", 1))
    msg("TrackableObjectGraph", ("nodes", 1, "TrackableObject", 1))
    msg("TensorProto", ("dtype", 1, "enum", 0), ("tensor_shape", 2, "TensorShapeProto", 0), ("tensor_content", 4, "bytes", 0),
        ("float_val", 5, "float", 1), ("int_val", 7, "int32", 1))
    msg("AttrValue", ("type", 6, "enum", 0), ("tensor", 8, "TensorProto", 0))
    msg("NodeDef", ("name", 1, "string", 0), ("op", 2, "string", 0), ("input", 3, "string", 1), maps=[("attr", 5, "AttrValue")])
    msg("OpDef", ("name", 1, "string", 0))
    msg("FunctionDef", ("signature", 1, "OpDef", 0), ("node_def", 3, "NodeDef", 1))
    msg("FunctionDefLibrary", ("function", 1, "FunctionDef", 1))
    msg("GraphDef", ("node", 1, "NodeDef", 1), ("library", 2, "FunctionDefLibrary", 0))
    msg("TensorInfo", ("name", 1, "string", 0), ("dtype", 2, "enum", 0), ("tensor_shape", 3, "TensorShapeProto", 0))
    msg("SignatureDef", ("method_name", 3, "string", 0), maps=[("inputs", 1, "TensorInfo"), ("outputs", 2, "TensorInfo")])
    msg("MetaGraphDef", ("graph_def", 2, "GraphDef", 0), maps=[("signature_def", 5, "SignatureDef")])
    msg("SavedModel", ("saved_model_schema_version", 1, "int64", 0), ("meta_graphs", 2, "MetaGraphDef", 1))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


_POOL = _build_pool()


def M(name):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName("tfs." + name))


def crc32c_bitwise(data):
    """CRC-32C straight from the definition (reflected polynomial 0x82F63B78), one bit at a time per byte table."""
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
        tab.append(c)
    tab = np.array(tab, np.uint32)
    c = 0xFFFFFFFF
    for b in bytes(data):
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked(c):
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _block(entries, restart_interval=16):
    buf, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(buf))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        buf += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    return bytes(buf) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def write_table(path, items, entries_per_block=40, snappy=True):
    import pyarrow as pa
    codec = pa.Codec("snappy")
    items = sorted(items)
    out = bytearray()

    def emit(block, compress):
        kind = 1 if compress else 0
        body = codec.compress(block, asbytes=True) if compress else block
        off = len(out)
        out.extend(body)
        out.append(kind)
        out.extend(struct.pack("<I", _masked(crc32c_bitwise(body + bytes([kind])))))
        return _varint(off) + _varint(len(body))

    index = []
    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        index.append((chunk[-1][0], emit(_block(chunk), snappy)))
    meta = emit(_block([]), False)
    idx = emit(_block(index, 1), False)
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57))
    with open(path, "wb") as fh:
        fh.write(bytes(out))


def keras_export_order(names):
    """layer names in the order Keras numbers `layer_with_weights-N` for a functional model: by decreasing depth from
    the outputs.  For this network the backbone is a chain (creation order); behind the first yolo block the three
    branches interleave.  The exact interleave does not matter to the reader - it must not depend on it - so the
    branches are simply rotated: a deterministic order that is NOT creation order."""
    names = list(names)
    head = [n for n in names if not n.startswith("feature_map")]
    dets = [n for n in names if n.startswith("feature_map")]
    cut = len(head) * 3 // 4
    tail = head[cut:]
    return head[:cut] + tail[1::2] + tail[0::2] + dets[::-1]


def write_reference_style_saved_model(path, weights, anchors, input_shape, name_offset=72, convt_offset=2, with_anchor_consts=True,
                                      checksum_limit=1 << 16):
    """weights: {keras_name_counted_from_zero: array}.  Layer names are shifted the way a second model in the same
    process is named (conv2d_72 .., batch_normalization_72 .., conv2d_transpose_2 ..)."""
    def shifted(layer):
        for base, off in (("conv2d_transpose", convt_offset), ("conv2d", name_offset), ("batch_normalization", name_offset)):
            if layer == base:
                return "%s_%d" % (base, off) if off else base
            if layer.startswith(base + "_") and layer[len(base) + 1:].isdigit():
                return "%s_%d" % (base, int(layer[len(base) + 1:]) + off)
        return layer

    layers = []
    for name in weights:
        layer = name.split("/")[0]
        if layer not in layers:
            layers.append(layer)
    order = keras_export_order(layers)
    os.makedirs(os.path.join(path, "variables"), exist_ok=True)
    graph = M("TrackableObjectGraph")()
    root = graph.nodes.add()
    shard_data = [bytearray(), bytearray()]
    items = []
    hdr = M("BundleHeaderProto")(num_shards=2, endianness=0)
    hdr.version.producer = 1
    items.append((b"", hdr.SerializeToString()))
    for i, layer in enumerate(order):
        root.children.add(node_id=i + 1, local_name="layer_with_weights-%d" % i)
        node = graph.nodes.add()
        for name, arr in weights.items():
            if name.split("/")[0] != layer:
                continue
            var = name.split("/")[1]
            key = "layer_with_weights-%d/%s/.ATTRIBUTES/VARIABLE_VALUE" % (i, var)
            node.attributes.add(name="VARIABLE_VALUE", full_name="%s/%s" % (shifted(layer), var), checkpoint_key=key)
            a = np.ascontiguousarray(arr, np.float32)
            raw = a.tobytes()
            shard = i % 2
            e = M("BundleEntryProto")(dtype=1, shard_id=shard, offset=len(shard_data[shard]), size=len(raw))
            for d in a.shape:
                e.shape.dim.add(size=int(d))
            if len(raw) <= checksum_limit:
                e.crc32c = _masked(crc32c_bitwise(raw))
            shard_data[shard] += raw
            items.append((key.encode(), e.SerializeToString()))
    g = graph.SerializeToString()
    ln = _varint(len(g))
    raw = ln + struct.pack("<I", _masked(crc32c_bitwise(ln))) + g
    e = M("BundleEntryProto")(dtype=7, shard_id=0, offset=len(shard_data[0]), size=len(raw), crc32c=_masked(crc32c_bitwise(raw)))
    shard_data[0] += raw
    items.append((b"_CHECKPOINTABLE_OBJECT_GRAPH", e.SerializeToString()))
    for s in range(2):
        with open(os.path.join(path, "variables", "variables.data-%05d-of-00002" % s), "wb") as fh:
            fh.write(bytes(shard_data[s]))
    write_table(os.path.join(path, "variables", "variables.index"), items)

    sm = M("SavedModel")(saved_model_schema_version=1)
    mg = sm.meta_graphs.add()
    sig = mg.signature_def["serving_default"]
    sig.method_name = "tensorflow/serving/predict"
    ti = sig.inputs["input_1"]
    ti.name, ti.dtype = "serving_default_input_1:0", 1
    for d in input_shape:
        ti.tensor_shape.dim.add(size=int(d))
    if with_anchor_consts:
        a = np.asarray(anchors, np.float32).reshape(-1, 2)
        for scale in range(3):                               # one decode per scale, each with its own constant
            fn = mg.graph_def.library.function.add()
            fn.signature.name = "__inference_tf_op_layer_mul_%d_layer_call_fn" % scale
            n_exp = fn.node_def.add(name="Exp", op="Exp")
            n_exp.input.append("inputs")
            n_exp.attr["T"].type = 1
            c = fn.node_def.add(name="mul/y", op="Const")
            c.attr["dtype"].type = 1
            t = c.attr["value"].tensor
            t.dtype = 1
            for d in a.shape:
                t.tensor_shape.dim.add(size=int(d))
            if scale == 1:
                t.float_val.extend([float(v) for v in a.reshape(-1)])     # the other encoding TF uses for small tensors
            else:
                t.tensor_content = a.tobytes()
            m = fn.node_def.add(name="mul", op="Mul")
            m.input.extend(["Exp:y:0", "mul/y:output:0"])
            # a decoy of the same rank: the (h, w) stride constant is rank 1 and must not be taken for anchors
            s_ = fn.node_def.add(name="mul_1/y", op="Const")
            s_.attr["dtype"].type = 1
            ts = s_.attr["value"].tensor
            ts.dtype = 1
            ts.tensor_shape.dim.add(size=2)
            ts.float_val.extend([32.0, 32.0])
            m2 = fn.node_def.add(name="mul_1", op="Mul")
            m2.input.extend(["add:z:0", "mul_1/y:output:0"])
    with open(os.path.join(path, "saved_model.pb"), "wb") as fh:
        fh.write(sm.SerializeToString())
