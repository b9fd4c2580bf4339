"""TensorFlow V1 checkpoint table reader / writer (vae_assoc_b200/tf_checkpoint.py): format restated from TensorFlow's
sources (unpinned: no TF-written file is available here), checked by a round trip, by the table's own structural
invariants (footer magic, block checksums, sorted keys) and by hand-computed known answers of the primitives."""
import struct

import numpy as np
import pytest

from vae_assoc_b200 import tf_checkpoint as tfc


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors for CRC32C (Castagnoli)
    assert tfc.crc32c(b"\x00" * 32) == 0x8A9136AA
    assert tfc.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E
    assert tfc.crc32c(b"123456789") == 0xE3069283


def test_ordered_code_key():
    # OrderedCode: num 0 -> "\0"; string "a" -> "a\0\1"; num 2 -> "\1\2"; -1 -> 0x7f
    assert tfc.encode_tensor_name_slice("a", 2) == b"\x00" + b"a\x00\x01" + b"\x01\x02" + b"\x7f\x7f\x7f\x7f"
    assert tfc.encode_tensor_name_slice("beta1_power", 0) == b"\x00beta1_power\x00\x01\x00"
    # keys sort like the tensor names
    names = ["image/Variable", "image/Variable_1", "image/Variable/Adam", "joint_1/Variable_5", "beta1_power"]
    keys = [tfc.encode_tensor_name_slice(n, 2) for n in names]
    assert [n for _, n in sorted(zip(keys, names))] == sorted(names)


def test_round_trip_and_structure(tmp_path):
    rng = np.random.RandomState(0)
    tensors = {"image/Variable": rng.normal(size=(784, 500)).astype(np.float32),     # > one 256 KB block
               "image/Variable_1": rng.normal(size=(500,)).astype(np.float32),
               "image/Variable/Adam": rng.normal(size=(784, 500)).astype(np.float32),
               "joint_1/deconv2d/weights": rng.normal(size=(3, 3, 8, 4)).astype(np.float32),
               "beta1_power": np.float32(0.9 ** 7), "beta2_power": np.float32(0.999 ** 7),
               "global_step": np.int64(7), "d": rng.normal(size=(3, 2))}
    path = tmp_path / "model.ckpt"
    tfc.write_v1(str(path), tensors)
    raw = path.read_bytes()
    assert struct.unpack("<Q", raw[-8:])[0] == tfc.TABLE_MAGIC
    back = tfc.read_v1(str(path), verify_checksums=True)
    assert sorted(back) == sorted(tensors)
    for k, v in tensors.items():
        assert back[k].shape == np.asarray(v).shape
        np.testing.assert_array_equal(back[k], np.asarray(v))
    assert back["global_step"].dtype == np.int64 and back["d"].dtype == np.float64
    # a flipped byte inside a data block is caught by the block checksum
    bad = bytearray(raw); bad[100] ^= 0x40
    (tmp_path / "bad.ckpt").write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        tfc.read_v1(str(tmp_path / "bad.ckpt"), verify_checksums=True)
    with pytest.raises(ValueError):
        (tmp_path / "junk.ckpt").write_bytes(b"not a checkpoint" * 10)
        tfc.read_v1(str(tmp_path / "junk.ckpt"))


def test_unpacked_float_val_and_tensor_content_are_read(tmp_path):
    """Old protobuf writers emit repeated floats unpacked (one fixed32 per element); newer ones use tensor_content."""
    name = "v"
    vals = np.array([1.5, -2.25, 3.0], np.float32)
    shape = tfc._shape_proto(vals.shape)
    meta = tfc._msg(1, tfc._msg(1, name.encode()) + tfc._msg(2, shape) + tfc._vint(3, tfc.DT_FLOAT) + tfc._msg(4, tfc._full_slice_proto(1)))
    for tensor in (tfc._vint(1, tfc.DT_FLOAT) + b"".join(bytes([(5 << 3) | 5]) + struct.pack("<f", x) for x in vals),
                   tfc._vint(1, tfc.DT_FLOAT) + tfc._msg(4, vals.tobytes())):
        saved = tfc._msg(1, name.encode()) + tfc._msg(2, tfc._full_slice_proto(1)) + tfc._msg(3, tensor)
        items = sorted([(b"", tfc._msg(1, meta)), (tfc.encode_tensor_name_slice(name, 1), tfc._msg(2, saved))])
        blk = tfc._BlockBuilder()
        for k, v in items:
            blk.add(k, v)
        p = tmp_path / "t.ckpt"
        with open(p, "wb") as f:
            off, size, nxt = tfc._write_block(f, 0, blk.finish())
            idx = tfc._BlockBuilder(); idx.add(items[-1][0], tfc._encode_handle(off, size))
            moff, msize, nxt = tfc._write_block(f, nxt, tfc._BlockBuilder().finish())
            ioff, isize, nxt = tfc._write_block(f, nxt, idx.finish())
            footer = tfc._encode_handle(moff, msize) + tfc._encode_handle(ioff, isize)
            f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", tfc.TABLE_MAGIC))
        np.testing.assert_array_equal(tfc.read_v1(str(p), verify_checksums=True)[name], vals)
