"""TensorFlow checkpoints without TensorFlow: read (and write) the V2 "tensor bundle" format that the reference's
training scripts produce (`saver = tf.train.Saver()` / `saver.save(sess, logFolder + "/model.ckpt")`,
ModelNet/ModelNet.py:199,269; restored by ModelNet/ModelNetEval.py), so that weights trained with the reference load
into a VariableStore -- whose variable names ARE the reference's (`Conv_1_weights`, `BNRELUDROP_final_BN/moving_mean`,
`Final_Logits_weights1`, ...: utils/MCConvBuilder.py:394-419, utils/MCNetworkUtils.py:20-142).

    tensors = read_tf_checkpoint("log/model.ckpt")          # {name: numpy array}, TensorFlow shapes and dtypes
    load_tf_checkpoint(store, "log/model.ckpt")             # model variables only (optimizer slots skipped), in place
    write_tf_checkpoint("out/model.ckpt", store.state_dict())

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table*: a port of LevelDB's table):
  <prefix>.index                    an SSTable: key "" -> BundleHeaderProto, key <variable name> -> BundleEntryProto
                                    (dtype, shape, shard, offset, size, masked CRC32C of the bytes)
  <prefix>.data-SSSSS-of-NNNNN      the tensors' bytes (little endian, row-major) at those offsets
SSTable: blocks of prefix-compressed (key, value) entries with a restart array, each followed by a 1-byte compression
type and a masked CRC32C; an index block maps separator keys to block handles; a 48-byte footer holds the handles of
the metaindex and index blocks and the magic number 0xdb4775248b80fb57.

Host-side utility: NumPy only, nothing here touches the GPU path.
"""
import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}
_DT_BFLOAT16 = 14

# ---- CRC32C (Castagnoli), as LevelDB / TensorFlow use it ---------------------------------------------------------
_POLY = 0x82F63B78


def _make_table():
    t = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ (_POLY if c & 1 else 0)
        t[i] = c
    return t


_TABLE = _make_table()
_TABLE_LIST = [int(x) for x in _TABLE]


def _crc32c_scalar(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    tab = _TABLE_LIST
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def _zeros_operator(nbytes):
    """32 x 32 GF(2) matrix that advances a CRC register over `nbytes` zero bytes (zlib's crc32_combine, CRC32C
    polynomial), as four 256-entry lookup tables."""
    odd = [_POLY] + [1 << (n - 1) for n in range(1, 32)]   # one zero bit
    even = _gf2_square(odd)                                  # two
    odd = _gf2_square(even)                                  # four
    op = None                                                # identity so far
    n = nbytes
    while True:
        even = _gf2_square(odd)                              # first pass: eight bits = one byte
        if n & 1:
            op = even if op is None else [_gf2_times(even, col) for col in op]
        n >>= 1
        if not n:
            break
        odd = _gf2_square(even)
        if n & 1:
            op = odd if op is None else [_gf2_times(odd, col) for col in op]
        n >>= 1
        if not n:
            break
    if op is None:
        op = [1 << k for k in range(32)]
    return [[_gf2_times(op, v << (8 * byte)) for v in range(256)] for byte in range(4)]


def crc32c(data):
    """CRC32C of a bytes-like object.  Large buffers are cut into equal chunks whose CRCs advance in lock-step through
    NumPy and are then combined (CRC is linear over GF(2)), so that tens of megabytes check in well under a second."""
    buf = np.frombuffer(memoryview(data).cast("B"), dtype=np.uint8)
    n = buf.shape[0]
    if n < (1 << 16):
        return _crc32c_scalar(buf.tobytes())
    lanes = 4096
    length = n // lanes
    body = buf[: lanes * length].reshape(lanes, length)
    c = np.full(lanes, 0xFFFFFFFF, dtype=np.uint32)
    cols = np.ascontiguousarray(body.T)                      # [length, lanes]: one byte of every lane per step
    for pos in range(length):
        c = _TABLE[(c ^ cols[pos]) & 0xFF] ^ (c >> np.uint32(8))
    c ^= np.uint32(0xFFFFFFFF)
    t0, t1, t2, t3 = _zeros_operator(length)
    total = 0
    for v in c.tolist():                                     # crc(A || B) = shift_len(B)(crc(A)) ^ crc(B)
        total = t0[total & 0xFF] ^ t1[(total >> 8) & 0xFF] ^ t2[(total >> 16) & 0xFF] ^ t3[total >> 24] ^ v
    tail = buf[lanes * length:].tobytes()
    return _crc32c_scalar(tail, total) if tail else total


def _mask(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


# ---- varints / the two protos ------------------------------------------------------------------------------------
def _get_varint(b, pos):
    shift = result = 0
    while True:
        byte = b[pos]
        pos += 1
        result |= (byte & 0x7F) << shift
        if not byte & 0x80:
            return result, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        byte = v & 0x7F
        v >>= 7
        if v:
            out.append(byte | 0x80)
        else:
            out.append(byte)
            return bytes(out)


def _parse_fields(b):
    """protobuf wire format -> list of (field number, wire type, value)"""
    pos, out = 0, []
    while pos < len(b):
        key, pos = _get_varint(b, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(b, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", b, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _get_varint(b, pos)
            v = bytes(b[pos: pos + ln])
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", b, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _parse_entry(b):
    """BundleEntryProto -> dict(dtype, shape, shard, offset, size, crc, sliced)"""
    e = {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0, "crc": None, "sliced": False}
    for field, _, v in _parse_fields(b):
        if field == 1:
            e["dtype"] = v
        elif field == 2:
            dims = []
            for f2, _, v2 in _parse_fields(v):
                if f2 == 2:                                  # TensorShapeProto.Dim
                    size = 0
                    for f3, _, v3 in _parse_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif field == 3:
            e["shard"] = v
        elif field == 4:
            e["offset"] = v
        elif field == 5:
            e["size"] = v
        elif field == 6:
            e["crc"] = v
        elif field == 7:
            e["sliced"] = True
    return e


def _field(num, wt, payload):
    return _put_varint((num << 3) | wt) + payload


def _encode_entry(dtype_id, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in
                    (_field(1, 0, _put_varint(int(s))) for s in shape))
    out = _field(1, 0, _put_varint(dtype_id)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


# ---- snappy (raw format), in case a writer compressed the index blocks --------------------------------------------
def _snappy_uncompress(b):
    n, pos = _get_varint(b, 0)
    out = bytearray()
    while pos < len(b):
        tag = b[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                extra = ln - 59
                ln = int.from_bytes(b[pos: pos + extra], "little")
                pos += extra
            ln += 1
            out += b[pos: pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | b[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(b[pos: pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(b[pos: pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy block")
        for _ in range(ln):                                  # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("corrupt snappy block (length)")
    return bytes(out)


# ---- SSTable ---------------------------------------------------------------------------------------------------
def _read_block(f, offset, size, verify):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError("truncated checkpoint index")
    contents, ctype, crc = raw[:size], raw[size], struct.unpack_from("<I", raw, size + 1)[0]
    if verify and _mask(crc32c(raw[: size + 1])) != crc:
        raise ValueError("checkpoint index: block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        contents = _snappy_uncompress(contents)
    elif ctype != 0:
        raise ValueError("checkpoint index: unknown block compression %d" % ctype)
    return contents


def _block_entries(block):
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 * (num_restarts + 1)
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos: pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos: pos + vlen])
        pos += vlen


def _read_table(path, verify):
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        size = f.tell()
        if size < 48:
            raise ValueError("%s is too short to be a checkpoint index" % path)
        f.seek(size - 48)
        footer = f.read(48)
        if struct.unpack_from("<Q", footer, 40)[0] != _MAGIC:
            raise ValueError("%s is not a TensorFlow V2 checkpoint index (bad magic number)" % path)
        pos = 0
        _, pos = _get_varint(footer, pos)        # metaindex handle
        _, pos = _get_varint(footer, pos)
        idx_off, pos = _get_varint(footer, pos)
        idx_size, pos = _get_varint(footer, pos)
        out = []
        for _, handle in _block_entries(_read_block(f, idx_off, idx_size, verify)):
            off, p = _get_varint(handle, 0)
            sz, p = _get_varint(handle, p)
            out.extend(_block_entries(_read_block(f, off, sz, verify)))
        return out


class _BlockBuilder:
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last, self.interval = bytearray(), [0], 0, b"", restart_interval

    def add(self, key, value):
        shared = 0
        if self.count % self.interval == 0:
            if self.count:
                self.restarts.append(len(self.buf))
        else:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        self.buf += key[shared:] + value
        self.last, self.count = key, self.count + 1

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
            struct.pack("<I", len(self.restarts))


def _write_table(path, items, block_size=4096):
    """items: (key bytes, value bytes) in ascending key order"""
    with open(path, "wb") as f:
        def emit(contents):
            off = f.tell()
            f.write(contents + b"\x00" + struct.pack("<I", _mask(crc32c(contents + b"\x00"))))
            return _put_varint(off) + _put_varint(len(contents))
        index, blk = _BlockBuilder(restart_interval=1), _BlockBuilder()
        for key, value in items:
            blk.add(key, value)
            if len(blk.buf) >= block_size:
                index.add(blk.last, emit(blk.finish()))
                blk = _BlockBuilder()
        if blk.count:
            index.add(blk.last, emit(blk.finish()))
        meta = emit(_BlockBuilder().finish())
        idx = emit(index.finish())
        footer = meta + idx
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))


# ---- public ----------------------------------------------------------------------------------------------------
def read_tf_checkpoint(prefix, verify=True):
    """All tensors of the checkpoint `prefix` (the path handed to saver.save / saver.restore, without extension):
    {variable name: numpy array}.  verify: check the block and tensor checksums."""
    entries = _read_table(prefix + ".index", verify)
    if not entries or entries[0][0] != b"":
        raise ValueError("%s.index has no bundle header" % prefix)
    num_shards, endian = 1, 0
    for field, _, v in _parse_fields(entries[0][1]):
        if field == 1:
            num_shards = v
        elif field == 2:
            endian = v
    if endian != 0:
        raise ValueError("big-endian checkpoints are not supported")
    files, out = {}, {}
    try:
        for key, value in entries[1:]:
            e = _parse_entry(value)
            name = key.decode("utf-8")
            if e["sliced"]:
                raise ValueError("variable %s is stored in slices (partitioned variable): not supported" % name)
            if e["dtype"] == _DT_BFLOAT16:
                dtype, bf16 = np.dtype(np.uint16), True
            elif e["dtype"] in _DTYPES:
                dtype, bf16 = np.dtype(_DTYPES[e["dtype"]]), False
            else:
                continue                                     # strings, resources, variants: not model weights
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dtype.itemsize != e["size"]:
                raise ValueError("variable %s: %d bytes stored, shape %s needs %d" %
                                 (name, e["size"], e["shape"], count * dtype.itemsize))
            f = files.get(e["shard"])
            if f is None:
                f = files[e["shard"]] = open("%s.data-%05d-of-%05d" % (prefix, e["shard"], num_shards), "rb")
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError("variable %s: data file truncated" % name)
            if verify and e["crc"] is not None and _mask(crc32c(raw)) != e["crc"]:
                raise ValueError("variable %s: checksum mismatch" % name)
            arr = np.frombuffer(raw, dtype=dtype).reshape(e["shape"]).copy()
            if bf16:
                arr = (arr.astype(np.uint32) << 16).view(np.float32)
            out[name] = arr
    finally:
        for f in files.values():
            f.close()
    return out


_OPTIMIZER_SUFFIXES = ("/Adam", "/Adam_1", "/Momentum", "/RMSProp", "/RMSProp_1", "/Adagrad", "/ExponentialMovingAverage")
_OPTIMIZER_NAMES = ("beta1_power", "beta2_power", "global_step")


def is_model_variable(name):
    """False for what a TF1 training script saves next to the weights: optimizer slots, Adam's power accumulators and
    the step counter (ModelNet/ModelNet.py:49-55,158)."""
    return not (name.endswith(_OPTIMIZER_SUFFIXES) or name in _OPTIMIZER_NAMES)


def load_tf_checkpoint(store, prefix, strict=False, verify=True):
    """Load the model variables of a reference checkpoint into `store` (a VariableStore): variables that already
    exist are overwritten in place (captured graphs and folded batch-norm caches stay valid), others are created, so
    the checkpoint can be loaded before or after the first forward.  strict: every variable already in the store must
    be in the checkpoint with the same shape.  Returns (names loaded, names skipped)."""
    tensors = read_tf_checkpoint(prefix, verify=verify)
    model = {k: v for k, v in tensors.items() if is_model_variable(k) and v.dtype.kind == "f"}
    if strict:
        missing = sorted(set(store.vars) - set(model))
        if missing:
            raise KeyError("checkpoint %s lacks %d variable(s) of the model, e.g. %s" % (prefix, len(missing), missing[:5]))
    store.load_state_dict({k: v.astype(np.float32) for k, v in model.items()})
    return sorted(model), sorted(set(tensors) - set(model))


def write_tf_checkpoint(prefix, tensors):
    """Write {name: array-like} as a TensorFlow V2 checkpoint (one shard) that tf.train.Saver / tf.train.load_checkpoint
    can read back: the way from this library to the reference's evaluation scripts."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items, offset = [], 0
    with open("%s.data-00000-of-00001" % prefix, "wb") as data:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            t = tensors[name]
            arr = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
            if arr.dtype not in _DTYPE_IDS:
                raise TypeError("variable %s: dtype %s cannot be written" % (name, arr.dtype))
            raw = arr.tobytes()                              # C order; a 0-d array is a scalar entry (no dims)
            data.write(raw)
            items.append((name.encode("utf-8"),
                          _encode_entry(_DTYPE_IDS[arr.dtype], arr.shape, offset, len(raw), _mask(crc32c(raw)))))
            offset += len(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, _put_varint(2) + _field(1, 0, _put_varint(1)))
    _write_table(prefix + ".index", [(b"", header)] + items)
