"""Rows of reth's trie tables from the engine's stored-node records.

`account_trie_rows` / `storage_trie_rows` return `[(key_bytes, value_bytes)]` in MDBX key order, byte for byte what
`write_trie_updates_sorted` (crates/storage/provider/src/providers/database/provider.rs:3125-3160) puts into
`AccountsTrie` / `StoragesTrie` (crates/storage/db-api/src/tables/mod.rs:484-494; storage v2 views :542-572).
The encoding itself is done by the C ABI (b200_account_trie_rows / b200_storage_trie_rows, host-only code in
csrc/table_rows.cu); this module marshals records into a `b200_updates` and slices the result.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import B200Error, Rows, Updates

KEYS_LEGACY = 0  # StoredNibbles / StoredNibblesSubKey (crates/trie/common/src/nibbles.rs:27-141)
KEYS_PACKED = 1  # PackedStoredNibbles / PackedStoredNibblesSubKey (nibbles.rs:143-300)


def _pack(path: bytes) -> bytes:
    out = bytearray(32)
    for i, nib in enumerate(path):
        out[i >> 1] |= (nib << 4) if not (i & 1) else nib
    return bytes(out)


class _Marshalled:
    """A b200_updates view over numpy arrays built from records `(trie_id, path_nibbles, state, tree, hash, [hashes])`."""

    def __init__(self, records):
        n = len(records)
        self.trie_id = np.array([r[0] for r in records], dtype=np.uint32).reshape(n)
        self.path_len = np.array([len(r[1]) for r in records], dtype=np.uint8).reshape(n)
        self.path_packed = np.frombuffer(b"".join(_pack(r[1]) for r in records) or b"\0", dtype=np.uint8).copy()
        self.state = np.array([r[2] for r in records], dtype=np.uint16).reshape(n)
        self.tree = np.array([r[3] for r in records], dtype=np.uint16).reshape(n)
        self.hash = np.array([r[4] for r in records], dtype=np.uint16).reshape(n)
        offs = np.zeros(n + 1, dtype=np.uint64)
        for i, r in enumerate(records):
            offs[i + 1] = offs[i] + len(r[5])
        self.offs = offs
        self.hashes = np.frombuffer(b"".join(h for r in records for h in r[5]) or b"\0", dtype=np.uint8).copy()
        u = Updates()
        u.n_nodes = n
        cast = lambda a, t: C.cast(a.ctypes.data, C.POINTER(t))
        u.trie_id, u.path_len, u.path_packed = cast(self.trie_id, C.c_uint32), cast(self.path_len, C.c_uint8), cast(self.path_packed, C.c_uint8)
        u.state_mask, u.tree_mask, u.hash_mask = cast(self.state, C.c_uint16), cast(self.tree, C.c_uint16), cast(self.hash, C.c_uint16)
        u.hash_offset, u.hashes = cast(self.offs, C.c_uint64), cast(self.hashes, C.c_uint8)
        self.struct = u


def _slice_rows(rows: Rows, lib) -> list:
    n = int(rows.n_rows)
    out = []
    if n:
        offs = np.ctypeslib.as_array(rows.row_offset, (n + 1,))
        kl = np.ctypeslib.as_array(rows.key_len, (n,))
        blob = np.ctypeslib.as_array(rows.bytes, (int(offs[n]),)).tobytes()
        for r in range(n):
            lo, hi, k = int(offs[r]), int(offs[r + 1]), int(kl[r])
            out.append((blob[lo:lo + k], blob[lo + k:hi]))
    lib.b200_rows_release(C.byref(rows))
    return out


class TableRows:
    """Rows straight from the C ABI (no per-row Python objects): numpy views on library-owned memory."""

    def __init__(self, rows: Rows, lib):
        self._rows, self._lib = rows, lib
        n = self.n_rows = int(rows.n_rows)
        self.row_offset = np.ctypeslib.as_array(rows.row_offset, (n + 1,)) if n else np.zeros(1, np.uint64)
        self.key_len = np.ctypeslib.as_array(rows.key_len, (n,)) if n else np.zeros(0, np.uint32)
        total = int(self.row_offset[n]) if n else 0
        self.bytes = np.ctypeslib.as_array(rows.bytes, (total,)) if total else np.zeros(0, np.uint8)

    def to_list(self) -> list:
        blob = self.bytes.tobytes()
        out = []
        for r in range(self.n_rows):
            lo, hi, k = int(self.row_offset[r]), int(self.row_offset[r + 1]), int(self.key_len[r])
            out.append((blob[lo:lo + k], blob[lo + k:hi]))
        return out

    def release(self):
        if self._rows is not None:
            self.row_offset = self.key_len = self.bytes = None
            self._lib.b200_rows_release(C.byref(self._rows))
            self._rows = None

    def __len__(self):
        return self.n_rows

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def rows_from_updates(account_updates=None, storage_updates=None, acct_keys=None, key_format: int = KEYS_LEGACY):
    """Raw path: `b200_updates` structs as the engine filled them (before release) -> (TableRows|None, TableRows|None)."""
    lib = _lib.load()
    res = []
    if account_updates is not None:
        rows = Rows()
        rc = lib.b200_account_trie_rows(C.cast(C.byref(account_updates), C.POINTER(Updates)), key_format, C.byref(rows))
        if rc:
            raise B200Error(rc, "b200_account_trie_rows")
        res.append(TableRows(rows, lib))
    else:
        res.append(None)
    if storage_updates is not None:
        keys = np.ascontiguousarray(acct_keys, dtype=np.uint8).reshape(-1, 32)
        rows = Rows()
        rc = lib.b200_storage_trie_rows(C.cast(C.byref(storage_updates), C.POINTER(Updates)), keys.ctypes.data,
                                        len(keys), key_format, C.byref(rows))
        if rc:
            raise B200Error(rc, "b200_storage_trie_rows")
        res.append(TableRows(rows, lib))
    else:
        res.append(None)
    return tuple(res)


def account_trie_rows(records, key_format: int = KEYS_LEGACY) -> list:
    """records: account-trie `TrieUpdates.account_nodes` as the engine returns them."""
    lib = _lib.load()
    m, rows = _Marshalled(records), Rows()
    rc = lib.b200_account_trie_rows(C.byref(m.struct), key_format, C.byref(rows))
    if rc:
        raise B200Error(rc, "b200_account_trie_rows")
    return _slice_rows(rows, lib)


def storage_trie_rows(records, acct_keys, key_format: int = KEYS_LEGACY) -> list:
    """records: storage-trie nodes with trie_id = account index; acct_keys: the (n,32) hashed addresses."""
    lib = _lib.load()
    keys = np.ascontiguousarray(acct_keys, dtype=np.uint8).reshape(-1, 32)
    m, rows = _Marshalled(records), Rows()
    rc = lib.b200_storage_trie_rows(C.byref(m.struct), keys.ctypes.data, len(keys), key_format, C.byref(rows))
    if rc:
        raise B200Error(rc, "b200_storage_trie_rows")
    return _slice_rows(rows, lib)


# ------------------------------------------------------------------------------------------------ checkpoint codecs (host)
def branch_node_compact_to_bytes(node) -> bytes:
    """`Compact for BranchNodeCompact` (alloy-trie; the value bytes of an AccountsTrie / StoragesTrie row): the three masks
    as big-endian u16, the root hash if present, then the child hashes."""
    out = bytearray()
    for m in (node.state_mask, node.tree_mask, node.hash_mask):
        out += int(m).to_bytes(2, "big")
    if getattr(node, "root_hash", None):
        out += node.root_hash
    for h in node.hashes:
        out += h
    return bytes(out)


def branch_node_compact_from_bytes(buf: bytes):
    """Inverse; the root hash is present iff there is one more 32-byte word than hash_mask has bits (alloy-trie's rule)."""
    from .trie import BranchNodeCompact
    if len(buf) < 6 or (len(buf) - 6) % 32:
        raise ValueError("BranchNodeCompact: 6 bytes of masks followed by 32-byte hashes")
    sm, tm, hm = (int.from_bytes(buf[i:i + 2], "big") for i in (0, 2, 4))
    words = [buf[6 + 32 * i:38 + 32 * i] for i in range((len(buf) - 6) // 32)]
    n = bin(hm).count("1")
    if len(words) == n + 1:
        return BranchNodeCompact(sm, tm, hm, tuple(words[1:]), words[0])
    if len(words) != n:
        raise ValueError("BranchNodeCompact: hash count does not match hash_mask")
    return BranchNodeCompact(sm, tm, hm, tuple(words), None)


class StoredSubNode:
    """crates/trie/common/src/subnode.rs:5-14 — one element of the walker stack inside reth's MerkleCheckpoint, with its
    Compact codec (:16-67): u16 key length, key, option flag + nibble, option flag + BranchNodeCompact.  This engine's own
    checkpoint is b200_stream_checkpoint; the codec is here so that a host can read and write reth's rows."""

    def __init__(self, key: bytes = b"", nibble=None, node=None):
        self.key, self.nibble, self.node = bytes(key), nibble, node

    def __eq__(self, o):
        return (self.key, self.nibble, self.node) == (o.key, o.nibble, o.node)

    def to_compact(self) -> bytes:
        out = bytearray(len(self.key).to_bytes(2, "big")) + self.key
        out += bytes([1, self.nibble]) if self.nibble is not None else b"\x00"
        if self.node is not None:
            out += b"\x01" + branch_node_compact_to_bytes(self.node)
        else:
            out += b"\x00"
        return bytes(out)

    @classmethod
    def from_compact(cls, buf: bytes) -> "StoredSubNode":
        n = int.from_bytes(buf[:2], "big")
        key, p = buf[2:2 + n], 2 + n
        nibble = None
        if buf[p]:
            nibble = buf[p + 1]
            p += 2
        else:
            p += 1
        node = branch_node_compact_from_bytes(buf[p + 1:]) if buf[p] else None
        return cls(key, nibble, node)


def _varuint(n: int) -> bytes:
    """reth-codecs `encode_varuint`: LEB128, low 7 bits first."""
    out = bytearray()
    while n >= 0x80:
        out.append(0x80 | (n & 0x7F))
        n >>= 7
    out.append(n)
    return bytes(out)


def _read_varuint(buf: bytes, p: int):
    n = shift = 0
    while True:
        b = buf[p]
        p += 1
        n |= (b & 0x7F) << shift
        if not b & 0x80:
            return n, p
        shift += 7


def _vec_u8_to_compact(v: bytes) -> bytes:
    """reth-codecs `Vec<T>::to_compact` instantiated at T = u8: varuint element count, then per element a varuint length
    and the element's own Compact bytes — a u8 drops its leading zero byte, so 0 is (length 0, no byte)."""
    out = bytearray(_varuint(len(v)))
    for x in v:
        out += b"\x00" if x == 0 else bytes([1, x])
    return bytes(out)


def _vec_u8_from_compact(buf: bytes, p: int):
    n, p = _read_varuint(buf, p)
    v = bytearray()
    for _ in range(n):
        ln, p = _read_varuint(buf, p)
        v.append(int.from_bytes(buf[p:p + ln], "big") if ln else 0)
        p += ln
    return bytes(v), p


class HashBuilderState:
    """crates/trie/common/src/hash_builder/state.rs:15-32 — alloy-trie's HashBuilder between two `add_leaf` calls, the form
    reth's MerkleCheckpoint stores — with its Compact codec (:66-140): key (nibbles, one per byte) as a Vec<u8>, u16 stack
    length + (u16 length, RlpNode bytes) per entry, the pending value (tag 0 + 32-byte hash | tag 1 + Vec<u8> of a leaf
    value), three u16-counted lists of big-endian u16 masks (groups, tree, hash), one byte stored_in_database.

    The element codecs (Vec<u8>, HashBuilderValue, TrieMask) live in the external crate reth-codecs 0.3.1 (Cargo.toml:328; not
    vendored under the reference) and are restated here from its published source: parity of this codec is unpinned — the
    reference tree holds no byte vector for it (state.rs:149-170 round-trips only), and so does tests/test_table_rows.py.
    This engine's resumable state is the frontier checkpoint (b200_root_stream_checkpoint), which a
    HashBuilder stack cannot express at a bucket boundary (the builder folds a key only when the next one arrives); the
    codec lets a host read and write the rows reth itself left in `StageCheckpoints`."""

    def __init__(self, key: bytes = b"", value=("bytes", b""), stack=(), groups=(), tree_masks=(), hash_masks=(),
                 stored_in_database: bool = False):
        self.key, self.value, self.stack = bytes(key), (value[0], bytes(value[1])), [bytes(s) for s in stack]
        self.groups, self.tree_masks, self.hash_masks = list(groups), list(tree_masks), list(hash_masks)
        self.stored_in_database = bool(stored_in_database)
        if any(n > 15 for n in self.key):
            raise ValueError("key holds one nibble per byte")
        if self.value[0] not in ("hash", "bytes") or (self.value[0] == "hash" and len(self.value[1]) != 32):
            raise ValueError("value is ('hash', 32 bytes) or ('bytes', leaf value)")

    def _tuple(self):
        return (self.key, self.value, self.stack, self.groups, self.tree_masks, self.hash_masks, self.stored_in_database)

    def __eq__(self, o):
        return self._tuple() == o._tuple()

    def to_compact(self) -> bytes:
        out = bytearray(_vec_u8_to_compact(self.key))
        out += len(self.stack).to_bytes(2, "big")
        for item in self.stack:
            out += len(item).to_bytes(2, "big") + item
        out += (b"\x00" + self.value[1]) if self.value[0] == "hash" else (b"\x01" + _vec_u8_to_compact(self.value[1]))
        for masks in (self.groups, self.tree_masks, self.hash_masks):
            out += len(masks).to_bytes(2, "big")
            for m in masks:
                out += int(m).to_bytes(2, "big")
        out.append(1 if self.stored_in_database else 0)
        return bytes(out)

    @classmethod
    def from_compact(cls, buf: bytes) -> "HashBuilderState":
        key, p = _vec_u8_from_compact(buf, 0)
        n = int.from_bytes(buf[p:p + 2], "big")
        p += 2
        stack = []
        for _ in range(n):
            ln = int.from_bytes(buf[p:p + 2], "big")
            stack.append(buf[p + 2:p + 2 + ln])
            p += 2 + ln
        tag = buf[p]
        p += 1
        if tag == 0:
            value = ("hash", buf[p:p + 32])
            p += 32
        elif tag == 1:
            v, p = _vec_u8_from_compact(buf, p)
            value = ("bytes", v)
        else:
            raise ValueError("HashBuilderValue tag %d" % tag)
        lists = []
        for _ in range(3):
            n = int.from_bytes(buf[p:p + 2], "big")
            p += 2
            lists.append([int.from_bytes(buf[p + 2 * i:p + 2 * i + 2], "big") for i in range(n)])
            p += 2 * n
        return cls(key, value, stack, lists[0], lists[1], lists[2], buf[p] != 0)
