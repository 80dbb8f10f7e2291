"""Reader / writer for the TensorFlow checkpoint *tensor bundle* (``variables.index`` +
``variables.data-00000-of-00001``) — the weight interchange format of the reference's
SavedModel directories.

Reference artefacts this reads:
  * ``OverlapDetection/timit/models/timit{1.0,2.0}/variables/variables.index``
  * ``SpeakerIdentification/timit/model/variables/variables.index``
loaded in the reference by ``tf.keras.models.load_model(dir)``
(``OverlapDetection/scripts/record_on_pc.py:87-88``,
``SpeakerIdentification/scripts/record_on_pc.py:76-77``).

The ``.data`` shards are stripped from the reference mount (``.MISSING_LARGE_BLOBS``), so the
reader is exercised against (a) the real ``.index`` files for names/shapes/offsets and (b)
bundles produced by :func:`write_bundle` with seeded synthetic weights.

Format (no TensorFlow needed): the index is a LevelDB-style SSTable — data blocks of
prefix-compressed ``(key, value)`` entries, an index block, and a 48-byte footer ending in the
magic ``0xdb4775248b80fb57``.  Key ``""`` maps to a ``BundleHeaderProto``; every other key maps
to a ``BundleEntryProto`` (dtype, shape, shard_id, offset, size, crc32c).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_FOOTER_LEN = 48

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8,
           9: np.int64, 10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}
DT_STRING = 7


@dataclass(frozen=True)
class BundleEntry:
    key: str
    dtype: int
    shape: Tuple[int, ...]
    shard_id: int
    offset: int
    size: int
    crc32c: int


# ---------------------------------------------------------------------------------------------
# varint / protobuf helpers
# ---------------------------------------------------------------------------------------------
def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _write_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Minimal protobuf wire parser: returns {field_number: [raw values]}."""
    fields: Dict[int, list] = {}
    pos = 0
    n = len(buf)
    while pos < n:
        tag, pos = _read_varint(buf, pos)
        fnum, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        fields.setdefault(fnum, []).append(v)
    return fields


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for d in _parse_proto(buf).get(2, []):      # TensorShapeProto.dim
        size = _parse_proto(d).get(1, [0])[0]   # Dim.size
        if size >= 1 << 63:
            size -= 1 << 64
        dims.append(int(size))
    return tuple(dims)


def _parse_entry(key: str, buf: bytes) -> BundleEntry:
    f = _parse_proto(buf)
    return BundleEntry(
        key=key,
        dtype=f.get(1, [0])[0],
        shape=_parse_shape(f[2][0]) if 2 in f else (),
        shard_id=f.get(3, [0])[0],
        offset=f.get(4, [0])[0],
        size=f.get(5, [0])[0],
        crc32c=f.get(6, [0])[0],
    )


# ---------------------------------------------------------------------------------------------
# SSTable reading
# ---------------------------------------------------------------------------------------------
def _block_entries(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos = 0
    key = b""
    while pos < limit:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_block(buf: bytes, offset: int, size: int) -> bytes:
    ctype = buf[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable blocks are not supported (TF bundles write none)")
    return buf[offset:offset + size]


def read_index(index_path: str) -> Tuple[Dict[str, int], List[BundleEntry]]:
    """Parse ``variables.index``.  Returns (header dict, entries in key order)."""
    with open(index_path, "rb") as f:
        buf = f.read()
    if len(buf) < _FOOTER_LEN:
        raise ValueError(f"{index_path}: too short for an SSTable")
    footer = buf[-_FOOTER_LEN:]
    if struct.unpack_from("<Q", footer, _FOOTER_LEN - 8)[0] != _MAGIC:
        raise ValueError(f"{index_path}: bad SSTable magic")
    pos = 0
    _, pos = _read_varint(footer, pos)          # metaindex offset
    _, pos = _read_varint(footer, pos)          # metaindex size
    idx_off, pos = _read_varint(footer, pos)
    idx_size, pos = _read_varint(footer, pos)
    header: Dict[str, int] = {}
    entries: List[BundleEntry] = []
    for _, handle in _block_entries(_read_block(buf, idx_off, idx_size)):
        boff, p = _read_varint(handle, 0)
        bsize, p = _read_varint(handle, p)
        for key, value in _block_entries(_read_block(buf, boff, bsize)):
            if key == b"":
                h = _parse_proto(value)
                header = {"num_shards": h.get(1, [0])[0], "endianness": h.get(2, [0])[0]}
            else:
                entries.append(_parse_entry(key.decode("utf-8"), value))
    return header, entries


def read_bundle(prefix: str, keys=None, verify_crc: bool = False) -> Dict[str, np.ndarray]:
    """Read tensors from ``<prefix>.index`` + ``<prefix>.data-0000x-of-0000n``.

    ``prefix`` is e.g. ``<model_dir>/variables/variables``.  Raises FileNotFoundError when a
    data shard is absent (the case for the stripped reference mount) — there is no fallback.
    """
    header, entries = read_index(prefix + ".index")
    nshards = max(1, header.get("num_shards", 1))
    out: Dict[str, np.ndarray] = {}
    shards: Dict[int, np.memmap] = {}
    for e in entries:
        if keys is not None and e.key not in keys:
            continue
        if e.dtype == DT_STRING or e.dtype not in _DTYPES:
            continue
        if e.shard_id not in shards:
            path = f"{prefix}.data-{e.shard_id:05d}-of-{nshards:05d}"
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            shards[e.shard_id] = np.memmap(path, dtype=np.uint8, mode="r")
        raw = np.asarray(shards[e.shard_id][e.offset:e.offset + e.size])
        if verify_crc and masked_crc32c(raw.tobytes()) != e.crc32c:
            raise ValueError(f"crc32c mismatch for {e.key}")
        out[e.key] = raw.view(_DTYPES[e.dtype]).reshape(e.shape).copy()
    return out


# ---------------------------------------------------------------------------------------------
# crc32c (Castagnoli), masked the way TF/LevelDB store it
# ---------------------------------------------------------------------------------------------
def _make_crc_table() -> np.ndarray:
    poly = 0x82F63B78
    tab = np.zeros(256, dtype=np.uint32)
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab[i] = c
    return tab


_CRC_TABLE = _make_crc_table()


def _make_slice8_tables():
    t = [[int(v) for v in _CRC_TABLE]]
    for k in range(1, 8):
        prev = t[k - 1]
        t.append([(prev[i] >> 8) ^ t[0][prev[i] & 0xFF] for i in range(256)])
    return t


_CRC_T8 = _make_slice8_tables()
_crc_native = None


def _native_crc():
    """``mmla_crc32c_host`` from libmmla_b200.so (host code, no GPU needed) when the library is built."""
    global _crc_native
    if _crc_native is None:
        try:
            from . import _lib
            _crc_native = _lib.load().mmla_crc32c_host
        except Exception:
            _crc_native = False
    return _crc_native


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli).  Weight shards are megabytes: the library's C routine is used when
    libmmla_b200.so is present, otherwise a slice-by-8 loop (8 bytes per Python iteration)."""
    fn = _native_crc()
    if fn and len(data) >= 64:
        buf = bytes(data)
        return int(fn(buf, len(buf))) & 0xFFFFFFFF
    t0, t1, t2, t3, t4, t5, t6, t7 = _CRC_T8
    c = 0xFFFFFFFF
    n8 = len(data) // 8
    if n8:
        words = struct.unpack_from("<%dQ" % n8, data)
        for w in words:
            w ^= c
            c = (t7[w & 0xFF] ^ t6[(w >> 8) & 0xFF] ^ t5[(w >> 16) & 0xFF] ^ t4[(w >> 24) & 0xFF] ^
                 t3[(w >> 32) & 0xFF] ^ t2[(w >> 40) & 0xFF] ^ t1[(w >> 48) & 0xFF] ^ t0[w >> 56])
    for b in data[n8 * 8:]:
        c = t0[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------
# writing (used to persist seeded synthetic weights in the reference's own format)
# ---------------------------------------------------------------------------------------------
def _shape_proto(shape) -> bytes:
    out = b""
    for d in shape:
        dim = b"\x08" + _write_varint(int(d))
        out += b"\x12" + _write_varint(len(dim)) + dim
    return out


def _entry_proto(dtype: int, shape, offset: int, size: int, crc: int) -> bytes:
    sp = _shape_proto(shape)
    out = b"\x08" + _write_varint(dtype)
    out += b"\x12" + _write_varint(len(sp)) + sp
    if offset:
        out += b"\x20" + _write_varint(offset)
    out += b"\x28" + _write_varint(size)
    out += b"\x35" + struct.pack("<I", crc)
    return out


def _build_block(items: List[Tuple[bytes, bytes]], restart_interval: int = 16) -> bytes:
    out = bytearray()
    restarts = []
    prev = b""
    for i, (k, v) in enumerate(items):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            m = min(len(prev), len(k))
            while shared < m and prev[shared] == k[shared]:
                shared += 1
        out += _write_varint(shared) + _write_varint(len(k) - shared) + _write_varint(len(v))
        out += k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _emit_block(f_out: bytearray, block: bytes) -> Tuple[int, int]:
    off = len(f_out)
    f_out += block
    f_out += b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00"))
    return off, len(block)


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray], with_crc: bool = True) -> None:
    """Write ``tensors`` as a single-shard TF tensor bundle readable by :func:`read_bundle`
    (and by TensorFlow's ``BundleReader``)."""
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    keys = sorted(tensors.keys(), key=lambda s: s.encode("utf-8"))
    items: List[Tuple[bytes, bytes]] = []
    header = b"\x08\x01" + b"\x1a\x02\x08\x01"     # num_shards=1, version{producer=1}
    items.append((b"", header))
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as fd:
        for k in keys:
            a = np.ascontiguousarray(tensors[k])
            raw = a.tobytes()
            crc = masked_crc32c(raw) if with_crc else 0
            items.append((k.encode("utf-8"),
                          _entry_proto(_DTYPE_IDS[a.dtype], a.shape, offset, len(raw), crc)))
            fd.write(raw)
            offset += len(raw)
    f_out = bytearray()
    # data blocks of ~4 KiB
    index_items: List[Tuple[bytes, bytes]] = []
    cur: List[Tuple[bytes, bytes]] = []
    cur_bytes = 0
    for kv in items:
        cur.append(kv)
        cur_bytes += len(kv[0]) + len(kv[1]) + 3
        if cur_bytes >= 4096:
            off, size = _emit_block(f_out, _build_block(cur))
            index_items.append((cur[-1][0] + b"\x00", _write_varint(off) + _write_varint(size)))
            cur, cur_bytes = [], 0
    if cur:
        off, size = _emit_block(f_out, _build_block(cur))
        index_items.append((cur[-1][0] + b"\x00", _write_varint(off) + _write_varint(size)))
    meta_off, meta_size = _emit_block(f_out, _build_block([]))
    idx_off, idx_size = _emit_block(f_out, _build_block(index_items, restart_interval=1))
    footer = _write_varint(meta_off) + _write_varint(meta_size)
    footer += _write_varint(idx_off) + _write_varint(idx_size)
    footer += b"\x00" * (_FOOTER_LEN - 8 - len(footer))
    footer += struct.pack("<Q", _MAGIC)
    f_out += footer
    with open(prefix + ".index", "wb") as fi:
        fi.write(bytes(f_out))
