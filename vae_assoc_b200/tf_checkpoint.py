"""Reader / writer of TensorFlow's V1 checkpoint file -- the single `<name>.ckpt` file `tf.train.Saver` wrote up to
TF 0.11 and the one the reference's `restore_model` looks for (`f.endswith('.ckpt')`, /root/reference/vae_assoc.py:437-463;
the shipped model `model_batchsize64_nz4_lambda8_weight50.ckpt`, baxter_vae_assoc_writer.py:598).

TensorFlow is not vendored in the reference and not installable here, so the format is RESTATED from TensorFlow's
published sources (unpinned; no real TF-written file was available to check against -- tests/test_tf_checkpoint.py
round-trips through the writer below, which follows the same sources):

  * the file is a LevelDB-style sorted table (tensorflow/core/lib/io/table*.cc, format.cc): data blocks of
    prefix-compressed (shared, non_shared, value_len varint32 | key delta | value) entries followed by a restart array
    (uint32 offsets + uint32 count); every block is followed by a 5-byte trailer (compression type, masked crc32c);
    an index block maps last-key -> BlockHandle(offset, size varint64); the 48-byte footer holds the metaindex and index
    handles and the magic 0xdb4775248b80fb57.  tensor_slice_writer.cc builds it with `kNoCompression`.
  * key ""  -> SavedTensorSlices{meta = SavedTensorSliceMeta{tensor[] = {name, shape, type, slice[]}, versions}}
    other   -> SavedTensorSlices{data = SavedSlice{name, slice, data = TensorProto}}   (saved_tensor_slice.proto);
    the key itself is OrderedCode(0, name, dims, start/length per dim) (saved_tensor_slice_util.cc) -- the reader does
    not need to decode it because every value carries its tensor name and slice.
  * DT_FLOAT tensors are stored in TensorProto.float_val (field 5, packed).

Only what a Saver over float variables writes is supported: full (unsliced) DT_FLOAT / DT_DOUBLE / DT_INT32 / DT_INT64
tensors.  Host-side file I/O only; nothing here is on the train step.
"""
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
BLOCK_SIZE = 262144          # table::Options::block_size default
RESTART_INTERVAL = 16

# ---------------------------------------------------------------------------------------------------------
# crc32c (Castagnoli), masked as in tensorflow/core/lib/hash/crc32c.h
# ---------------------------------------------------------------------------------------------------------
_CRC_TABLE = None


def _crc_table():
    global _CRC_TABLE
    if _CRC_TABLE is None:
        poly = 0x82F63B78
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    return _CRC_TABLE


def crc32c(data, crc=0):
    tab = _crc_table()
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------
# varints / protobuf wire format (just enough for saved_tensor_slice.proto and tensor.proto)
# ---------------------------------------------------------------------------------------------------------
def _put_varint(out, v):
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)


def _get_varint(buf, pos):
    shift, v = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf):
    """Yields (field number, wire type, value) of one protobuf message; value = int (varint / fixed) or a memoryview."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
            if len(v) != ln:
                raise ValueError("truncated length-delimited field")
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield num, wt, v


def _msg(num, payload):
    out = bytearray()
    _put_varint(out, (num << 3) | 2)
    _put_varint(out, len(payload))
    return bytes(out) + bytes(payload)


def _vint(num, v):
    out = bytearray()
    _put_varint(out, (num << 3) | 0)
    _put_varint(out, v)
    return bytes(out)


def _shape_proto(shape):
    return b"".join(_msg(2, _vint(1, int(d))) for d in shape)          # TensorShapeProto.dim[].size


def _parse_shape(buf):
    dims = []
    for num, wt, v in _fields(buf):
        if num == 2 and wt == 2:
            size = 0
            for n2, w2, v2 in _fields(v):
                if n2 == 1 and w2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return tuple(dims)


def _full_slice_proto(ndim):
    # TensorSliceProto.extent[] with neither start nor length set = "everything" along that dimension
    return b"".join(_msg(1, b"") for _ in range(ndim))


# ---------------------------------------------------------------------------------------------------------
# OrderedCode pieces used by EncodeTensorNameSlice (keys sort by tensor name)
# ---------------------------------------------------------------------------------------------------------
def _oc_num_increasing(v):
    body = b"" if v == 0 else v.to_bytes((v.bit_length() + 7) // 8, "big")
    return bytes([len(body)]) + body


def _oc_string(s):
    """OrderedCode::WriteString: 0x00 -> 00 ff, 0xff -> ff 00, terminator 00 01."""
    b = s.encode() if isinstance(s, str) else bytes(s)
    out = bytearray()
    for c in b:
        if c == 0:
            out += b"\x00\xff"
        elif c == 0xFF:
            out += b"\xff\x00"
        else:
            out.append(c)
    return bytes(out) + b"\x00\x01"


def _oc_signed_increasing(v):
    """OrderedCode::WriteSignedNumIncreasing, one-byte form (|v| < 64): 0x80 ^ v.  Full slices only use -1."""
    if -64 <= v < 64:
        return bytes([(0x80 ^ v) & 0xFF])
    raise ValueError("slice extents beyond one byte are not written by this module")


def encode_tensor_name_slice(name, ndim):
    key = _oc_num_increasing(0) + _oc_string(name) + _oc_num_increasing(ndim)
    for _ in range(ndim):
        key += _oc_signed_increasing(-1) + _oc_signed_increasing(-1)
    return key


# ---------------------------------------------------------------------------------------------------------
# table reader
# ---------------------------------------------------------------------------------------------------------
def _read_block(buf, offset, size, verify):
    data = buf[offset:offset + size]
    if len(data) != size or offset + size + 5 > len(buf):
        raise ValueError("block handle out of range")
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed table blocks (type %d) are not supported; tf.train.Saver V1 writes none" % ctype)
    if verify:
        want = struct.unpack_from("<I", buf, offset + size + 1)[0]
        got = mask_crc(crc32c(bytes(buf[offset:offset + size + 1])))
        if want != got:
            raise ValueError("block checksum mismatch at offset %d" % offset)
    return data


def _block_entries(block):
    n = len(block)
    if n < 4:
        raise ValueError("block too small")
    num_restarts = struct.unpack_from("<I", block, n - 4)[0]
    limit = n - 4 - 4 * num_restarts
    if limit < 0:
        raise ValueError("bad restart array")
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _handle(buf, pos=0):
    off, pos = _get_varint(buf, pos)
    size, pos = _get_varint(buf, pos)
    return off, size, pos


def _table_items(buf, verify):
    if len(buf) < 48:
        raise ValueError("file too short for a table footer")
    footer = buf[len(buf) - 48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError("not a TensorFlow V1 checkpoint (table magic missing)")
    _, _, pos = _handle(footer, 0)                    # metaindex (unused)
    ioff, isize, _ = _handle(footer, pos)
    for _, hv in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, size, _ = _handle(hv, 0)
        for kv in _block_entries(_read_block(buf, off, size, verify)):
            yield kv


_NP = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64}


def _parse_tensor_proto(buf, dtype_hint):
    dtype, shape, content, vals = dtype_hint, None, None, []
    for num, wt, v in _fields(buf):
        if num == 1 and wt == 0:
            dtype = v
        elif num == 2 and wt == 2:
            shape = _parse_shape(v)
        elif num == 4 and wt == 2:
            content = bytes(v)
        elif num == 5:                                 # float_val: packed (wt 2) or one fixed32 per element
            vals.append(np.frombuffer(bytes(v), "<f4") if wt == 2 else np.array([v], "<u4").view("<f4"))
        elif num == 6:                                 # double_val
            vals.append(np.frombuffer(bytes(v), "<f8") if wt == 2 else np.array([v], "<u8").view("<f8"))
        elif num in (7, 10):                           # int_val / int64_val (varints)
            if wt == 2:
                p, out = 0, []
                while p < len(v):
                    x, p = _get_varint(v, p)
                    out.append(x if x < (1 << 63) else x - (1 << 64))
                vals.append(np.array(out, np.int64))
            else:
                vals.append(np.array([v if v < (1 << 63) else v - (1 << 64)], np.int64))
    if dtype not in _NP:
        raise ValueError("unsupported tensor dtype %r" % dtype)
    if content is not None:
        arr = np.frombuffer(content, _NP[dtype])
    else:
        arr = np.concatenate(vals).astype(_NP[dtype]) if vals else np.zeros(0, _NP[dtype])
    return arr, shape


def read_v1(path, verify_checksums=False):
    """{variable name: numpy array} of a TensorFlow V1 checkpoint file (every variable a tf.train.Saver saved: the model's
    `<scope>/Variable_k`, their `.../Adam`, `.../Adam_1` slots and `beta1_power`, `beta2_power`)."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    meta, data = {}, {}
    for key, value in _table_items(buf, verify_checksums):
        for num, wt, v in _fields(value):
            if num == 1 and wt == 2 and key == b"":          # SavedTensorSliceMeta
                for n2, w2, v2 in _fields(v):
                    if n2 == 1 and w2 == 2:                  # SavedSliceMeta
                        name, shape, dtype = None, (), DT_FLOAT
                        for n3, w3, v3 in _fields(v2):
                            if n3 == 1: name = bytes(v3).decode()
                            elif n3 == 2: shape = _parse_shape(v3)
                            elif n3 == 3: dtype = v3
                        meta[name] = (shape, dtype)
            elif num == 2 and wt == 2:                       # SavedSlice
                name, tensor, extents = None, None, []
                for n2, w2, v2 in _fields(v):
                    if n2 == 1: name = bytes(v2).decode()
                    elif n2 == 2:
                        for n3, w3, v3 in _fields(v2):
                            if n3 == 1: extents.append(bytes(v3))
                    elif n2 == 3: tensor = v2
                if any(len(e) for e in extents):
                    raise ValueError("partitioned variable slices are not supported (%s)" % name)
                data.setdefault(name, []).append(tensor)
    if not meta:
        raise ValueError("%s holds no SavedTensorSliceMeta entry" % path)
    out = {}
    for name, (shape, dtype) in meta.items():
        if name not in data:
            raise ValueError("checkpoint lists %s but holds no data for it" % name)
        arr, _ = _parse_tensor_proto(data[name][0], dtype)
        n = int(np.prod(shape)) if len(shape) else 1
        if arr.size != n:
            raise ValueError("%s: %d values for shape %s" % (name, arr.size, shape))
        out[name] = arr.reshape(shape).copy()
    return out


# ---------------------------------------------------------------------------------------------------------
# table writer (export a model for a TensorFlow-0.x consumer; also what the round-trip test uses)
# ---------------------------------------------------------------------------------------------------------
class _BlockBuilder(object):
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count < RESTART_INTERVAL:
            m = min(len(self.last), len(key))
            while shared < m and self.last[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf)); self.count = 0
        _put_varint(self.buf, shared); _put_varint(self.buf, len(key) - shared); _put_varint(self.buf, len(value))
        self.buf += key[shared:]; self.buf += value
        self.last, self.count = key, self.count + 1

    def empty(self):
        return not self.buf

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _write_block(f, offset, contents):
    trailer = b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00")))
    f.write(contents); f.write(trailer)
    return offset, len(contents), offset + len(contents) + 5


def _encode_handle(off, size):
    out = bytearray()
    _put_varint(out, off); _put_varint(out, size)
    return bytes(out)


def write_v1(path, tensors):
    """Writes {name: array} as a TensorFlow V1 checkpoint (one table file, no compression)."""
    items = []
    meta = b""
    for name in sorted(tensors):
        a = np.asarray(tensors[name])
        if a.dtype == np.float64: dt, a = DT_DOUBLE, a
        elif a.dtype == np.int32: dt, a = DT_INT32, a
        elif a.dtype == np.int64: dt, a = DT_INT64, a
        else: dt, a = DT_FLOAT, a.astype(np.float32)
        slice_proto = _full_slice_proto(a.ndim)
        meta += _msg(1, _msg(1, name.encode()) + _msg(2, _shape_proto(a.shape)) + _vint(3, dt) + _msg(4, slice_proto))
        if dt == DT_FLOAT: vals = _msg(5, a.astype("<f4").tobytes())
        elif dt == DT_DOUBLE: vals = _msg(6, a.astype("<f8").tobytes())
        else:
            packed = bytearray()
            for x in a.reshape(-1).tolist():
                _put_varint(packed, int(x))
            vals = _msg(7 if dt == DT_INT32 else 10, packed)
        tensor = _vint(1, dt) + _msg(2, _shape_proto(a.shape)) + vals
        saved = _msg(1, name.encode()) + _msg(2, slice_proto) + _msg(3, tensor)
        items.append((encode_tensor_name_slice(name, a.ndim), _msg(2, saved)))
    meta += _msg(2, _vint(1, 0))                                   # VersionDef{producer = 0}
    items.append((b"", _msg(1, meta)))
    items.sort(key=lambda kv: kv[0])
    with open(path, "wb") as f:
        offset, index, blk = 0, _BlockBuilder(), _BlockBuilder()
        pending = None

        def flush():
            nonlocal offset, blk, pending
            if blk.empty():
                return
            last = blk.last
            off, size, offset = _write_block(f, offset, blk.finish())
            pending = (last, _encode_handle(off, size))
            blk = _BlockBuilder()

        for key, value in items:
            if pending is not None:
                index.add(pending[0], pending[1]); pending = None
            blk.add(key, value)
            if blk.size() >= BLOCK_SIZE:
                flush()
        flush()
        if pending is not None:
            index.add(pending[0], pending[1])
        moff, msize, offset = _write_block(f, offset, _BlockBuilder().finish())
        ioff, isize, offset = _write_block(f, offset, index.finish())
        footer = _encode_handle(moff, msize) + _encode_handle(ioff, isize)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
