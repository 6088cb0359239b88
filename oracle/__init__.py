"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by anything under reth_b200/.

The C sources restate reth's state-commitment algorithm (see oracle/oracle.h for the
reference file:line each function follows).  Parity status: pinned by the reference's golden
vectors, tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

ACCOUNT_DTYPE = np.dtype([("nonce", "<u8"), ("balance", "u1", (32,)), ("code_hash", "u1", (32,))])
assert ACCOUNT_DTYPE.itemsize == 72

KECCAK_EMPTY = bytes.fromhex("c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470")
EMPTY_ROOT_HASH = bytes.fromhex("56e81f171bcc55a6ff8345e692c0f86e5b48e01b996cadc001622fb5e363b421")


def build(force: bool = False) -> str:
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("keccak.c", "keccak_avx512.c", "hash_builder.c", "state_root.c", "ordered_root.c", "oracle.h",
                                             "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


class _Updates(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint64),
        ("trie_id", C.POINTER(C.c_uint32)),
        ("path_len", C.POINTER(C.c_uint8)),
        ("path_packed", C.POINTER(C.c_uint8)),
        ("state_mask", C.POINTER(C.c_uint16)),
        ("tree_mask", C.POINTER(C.c_uint16)),
        ("hash_mask", C.POINTER(C.c_uint16)),
        ("hash_offset", C.POINTER(C.c_uint64)),
        ("hashes", C.POINTER(C.c_uint8)),
    ]


class _BranchNode(C.Structure):
    _fields_ = [
        ("path", C.c_uint8 * 64),
        ("path_len", C.c_uint8),
        ("state_mask", C.c_uint16),
        ("tree_mask", C.c_uint16),
        ("hash_mask", C.c_uint16),
        ("n_hashes", C.c_uint8),
        ("hashes", (C.c_uint8 * 32) * 16),
        ("has_root_hash", C.c_uint8),
        ("root_hash", C.c_uint8 * 32),
    ]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("leaves", "branch_nodes", "extension_nodes", "hashed_nodes", "keccak_f", "rlp_bytes_hashed")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u8p, u64p = C.c_void_p, C.c_void_p, C.c_void_p
        L.orc_keccak256.argtypes = [u8p, C.c_size_t, u8p]
        L.orc_keccak256_fixed.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint64, u8p, C.c_int]
        L.orc_keccak256_var.argtypes = [u8p, u64p, C.c_uint64, u8p, C.c_int]
        L.orc_keccak256_fixed_simd.restype = C.c_int
        L.orc_keccak256_fixed_simd.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_uint64, u8p, C.c_int]
        L.orc_hb_new.restype = vp
        L.orc_hb_new.argtypes = [C.c_int]
        L.orc_hb_free.argtypes = [vp]
        L.orc_hb_add_leaf.argtypes = [vp, u8p, C.c_size_t, u8p, C.c_size_t]
        L.orc_hb_add_branch.argtypes = [vp, u8p, C.c_size_t, u8p, C.c_int]
        L.orc_hb_root.argtypes = [vp, u8p]
        L.orc_hb_updates_len.restype = C.c_size_t
        L.orc_hb_updates_len.argtypes = [vp]
        L.orc_hb_update_at.restype = C.POINTER(_BranchNode)
        L.orc_hb_update_at.argtypes = [vp, C.c_size_t]
        L.orc_hb_nodes_len.restype = C.c_size_t
        L.orc_hb_nodes_len.argtypes = [vp]
        L.orc_hb_node_at.restype = C.POINTER(C.c_uint8)
        L.orc_hb_node_at.argtypes = [vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_hb_retain_nodes.argtypes = [vp, C.c_int]
        L.orc_encode_trie_account.restype = C.c_size_t
        L.orc_encode_trie_account.argtypes = [vp, u8p, u8p]
        L.orc_encode_u256.restype = C.c_size_t
        L.orc_encode_u256.argtypes = [u8p, u8p]
        L.orc_updates_free.argtypes = [C.POINTER(_Updates)]
        L.orc_storage_roots.argtypes = [u8p, u8p, u64p, C.c_uint64, u8p, C.POINTER(_Updates), C.c_int]
        L.orc_state_root.argtypes = [u8p, vp, u8p, C.c_uint64, u8p, C.POINTER(_Updates)]
        L.orc_state_root_full.argtypes = [u8p, vp, C.c_uint64, u8p, u8p, u64p, u8p,
                                          C.POINTER(_Updates), C.POINTER(_Updates), C.c_int]
        L.orc_trie_root_recursive.argtypes = [u8p, u8p, u64p, C.c_uint64, u8p]
        L.orc_ordered_roots.argtypes = [u8p, u64p, u64p, C.c_uint64, u8p]
        L.orc_stats_reset.argtypes = []
        L.orc_stats_get.argtypes = [C.POINTER(_Stats)]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _c(a, dtype=np.uint8):
    return np.ascontiguousarray(a, dtype=dtype)


# ------------------------------------------------------------------ keccak
def keccak256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    lib().orc_keccak256(C.cast(C.c_char_p(bytes(data)), C.c_void_p), len(data), C.cast(out, C.c_void_p))
    return out.raw


def keccak256_fixed(msgs: np.ndarray, msg_len: int | None = None, threads: int = 1) -> np.ndarray:
    """msgs: uint8[n, stride]; hashes the first msg_len bytes of each row."""
    msgs = _c(msgs)
    n, stride = msgs.shape
    out = np.empty((n, 32), np.uint8)
    lib().orc_keccak256_fixed(_p(msgs), msg_len or stride, stride, n, _p(out), threads)
    return out


def keccak256_fixed_simd(msgs: np.ndarray, msg_len: int | None = None, threads: int = 1):
    """8-way AVX-512 multi-buffer variant (best-effort CPU figure); None when the CPU lacks AVX-512F."""
    msgs = _c(msgs)
    n, stride = msgs.shape
    out = np.empty((n, 32), np.uint8)
    ok = lib().orc_keccak256_fixed_simd(_p(msgs), msg_len or stride, stride, n, _p(out), threads)
    return out if ok else None


def keccak256_var(data: np.ndarray, offsets: np.ndarray, threads: int = 1) -> np.ndarray:
    data = _c(data)
    offsets = _c(offsets, np.uint64)
    n = len(offsets) - 1
    out = np.empty((n, 32), np.uint8)
    lib().orc_keccak256_var(_p(data), _p(offsets), n, _p(out), threads)
    return out


# ------------------------------------------------------------------ encodings
def encode_u256(value: int) -> bytes:
    v = np.frombuffer(int(value).to_bytes(32, "big"), np.uint8).copy()
    out = np.empty(33, np.uint8)
    n = lib().orc_encode_u256(_p(v), _p(out))
    return out[:n].tobytes()


def make_accounts(rows) -> np.ndarray:
    """rows: iterable of (nonce:int, balance:int, code_hash:bytes|None) -> ACCOUNT_DTYPE array."""
    rows = list(rows)
    a = np.zeros(len(rows), ACCOUNT_DTYPE)
    for i, (nonce, balance, code_hash) in enumerate(rows):
        a[i]["nonce"] = nonce
        a[i]["balance"] = np.frombuffer(int(balance).to_bytes(32, "big"), np.uint8)
        a[i]["code_hash"] = np.frombuffer(code_hash or KECCAK_EMPTY, np.uint8)
    return a


def encode_trie_account(nonce: int, balance: int, storage_root: bytes = EMPTY_ROOT_HASH,
                        code_hash: bytes | None = None) -> bytes:
    a = make_accounts([(nonce, balance, code_hash)])
    sr = np.frombuffer(storage_root, np.uint8).copy()
    out = np.empty(112, np.uint8)
    n = lib().orc_encode_trie_account(_p(a), _p(sr), _p(out))
    return out[:n].tobytes()


def unpack_nibbles(key: bytes) -> bytes:
    return bytes(x for b in key for x in (b >> 4, b & 15))


# ------------------------------------------------------------------ HashBuilder
class HashBuilder:
    """alloy-trie HashBuilder restatement (oracle/hash_builder.c)."""

    def __init__(self, retain_updates: bool = False, retain_nodes: bool = False):
        self._h = lib().orc_hb_new(int(retain_updates))
        if retain_nodes:
            lib().orc_hb_retain_nodes(self._h, 1)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_hb_free(self._h)
            self._h = None

    def add_leaf(self, key_nibbles: bytes, value: bytes):
        k = np.frombuffer(bytes(key_nibbles), np.uint8).copy()
        v = np.frombuffer(bytes(value), np.uint8).copy()
        if lib().orc_hb_add_leaf(self._h, _p(k), len(k), _p(v), len(v)) != 0:
            raise ValueError("add_leaf: keys must be strictly increasing")

    def add_branch(self, key_nibbles: bytes, hash32: bytes, stored_in_database: bool):
        k = np.frombuffer(bytes(key_nibbles), np.uint8).copy()
        hh = np.frombuffer(bytes(hash32), np.uint8).copy()
        if lib().orc_hb_add_branch(self._h, _p(k), len(k), _p(hh), int(stored_in_database)) != 0:
            raise ValueError("add_branch: keys must be strictly increasing")

    def root(self) -> bytes:
        out = np.empty(32, np.uint8)
        lib().orc_hb_root(self._h, _p(out))
        return out.tobytes()

    def updates(self) -> dict:
        """{path nibbles (bytes): dict(state_mask, tree_mask, hash_mask, hashes, root_hash)} incl. empty path."""
        res = {}
        for i in range(lib().orc_hb_updates_len(self._h)):
            bn = lib().orc_hb_update_at(self._h, i).contents
            path = bytes(bn.path[: bn.path_len])
            res[path] = dict(
                state_mask=bn.state_mask, tree_mask=bn.tree_mask, hash_mask=bn.hash_mask,
                hashes=[bytes(bn.hashes[j]) for j in range(bn.n_hashes)],
                root_hash=bytes(bn.root_hash) if bn.has_root_hash else None,
            )
        return res

    def nodes(self) -> list:
        out = []
        ln = C.c_size_t()
        for i in range(lib().orc_hb_nodes_len(self._h)):
            p = lib().orc_hb_node_at(self._h, i, C.byref(ln))
            out.append(bytes(p[: ln.value]))
        return out


# ------------------------------------------------------------------ roots
def _updates_to_py(u: _Updates) -> list:
    """-> sorted list of (trie_id, path_nibbles, state, tree, hash, [hashes])."""
    n = u.n_nodes
    res = []
    if n:
        tid = np.ctypeslib.as_array(u.trie_id, (n,)).copy()
        pl = np.ctypeslib.as_array(u.path_len, (n,)).copy()
        pp = np.ctypeslib.as_array(u.path_packed, (n, 32)).copy()
        sm = np.ctypeslib.as_array(u.state_mask, (n,)).copy()
        tm = np.ctypeslib.as_array(u.tree_mask, (n,)).copy()
        hm = np.ctypeslib.as_array(u.hash_mask, (n,)).copy()
        ho = np.ctypeslib.as_array(u.hash_offset, (n + 1,)).copy()
        nh = int(ho[n])
        hs = np.ctypeslib.as_array(u.hashes, (max(nh, 1), 32)).copy()
        for i in range(n):
            path = unpack_nibbles(pp[i].tobytes())[: int(pl[i])]
            res.append((int(tid[i]), path, int(sm[i]), int(tm[i]), int(hm[i]),
                        [hs[j].tobytes() for j in range(int(ho[i]), int(ho[i + 1]))]))
    lib().orc_updates_free(C.byref(u))
    res.sort(key=lambda r: (r[0], r[1]))
    return res


def storage_roots(slot_keys, values, seg_offsets, want_updates=False, threads=1):
    slot_keys = _c(slot_keys).reshape(-1, 32)
    values = _c(values).reshape(-1, 32)
    seg_offsets = _c(seg_offsets, np.uint64)
    m = len(seg_offsets) - 1
    roots = np.empty((m, 32), np.uint8)
    u = _Updates()
    rc = lib().orc_storage_roots(_p(slot_keys), _p(values), _p(seg_offsets), m, _p(roots),
                                 C.byref(u) if want_updates else None, threads)
    if rc:
        raise ValueError(f"orc_storage_roots rc={rc}")
    return (roots, _updates_to_py(u)) if want_updates else roots


def pack_lists(lists):
    """[[bytes, ...], ...] -> (values blob u8, value_offsets u64 [n+1], seg_offsets u64 [n_lists+1])."""
    items = [it for l in lists for it in l]
    value_offsets = np.zeros(len(items) + 1, np.uint64)
    if items:
        value_offsets[1:] = np.cumsum([len(it) for it in items], dtype=np.uint64)
    seg_offsets = np.zeros(len(lists) + 1, np.uint64)
    if lists:
        seg_offsets[1:] = np.cumsum([len(l) for l in lists], dtype=np.uint64)
    values = np.frombuffer(b"".join(items), np.uint8).copy() if items else np.zeros(0, np.uint8)
    return values, value_offsets, seg_offsets


def ordered_roots(values, value_offsets, seg_offsets) -> np.ndarray:
    """Transactions / receipts / withdrawals roots of lists of pre-encoded items (ordered_root.rs:202-257)."""
    values = _c(values)
    value_offsets = _c(value_offsets, np.uint64)
    seg_offsets = _c(seg_offsets, np.uint64)
    m = len(seg_offsets) - 1
    roots = np.empty((m, 32), np.uint8)
    keep = values if len(values) else np.zeros(1, np.uint8)
    rc = lib().orc_ordered_roots(_p(keep), _p(value_offsets), _p(seg_offsets), m, _p(roots))
    if rc:
        raise ValueError(f"orc_ordered_roots rc={rc}")
    return roots


def state_root(acct_keys, accounts, storage_roots32=None, want_updates=False):
    acct_keys = _c(acct_keys).reshape(-1, 32)
    accounts = np.ascontiguousarray(accounts, ACCOUNT_DTYPE)
    sr = None if storage_roots32 is None else _c(storage_roots32).reshape(-1, 32)
    root = np.empty(32, np.uint8)
    u = _Updates()
    rc = lib().orc_state_root(_p(acct_keys), _p(accounts), _p(sr), len(acct_keys), _p(root),
                              C.byref(u) if want_updates else None)
    if rc:
        raise ValueError(f"orc_state_root rc={rc}")
    return (root.tobytes(), _updates_to_py(u)) if want_updates else root.tobytes()


def state_root_full(acct_keys, accounts, slot_keys, values, seg_offsets, want_updates=False, threads=1):
    acct_keys = _c(acct_keys).reshape(-1, 32)
    accounts = np.ascontiguousarray(accounts, ACCOUNT_DTYPE)
    slot_keys = _c(slot_keys).reshape(-1, 32)
    values = _c(values).reshape(-1, 32)
    seg_offsets = _c(seg_offsets, np.uint64)
    assert len(seg_offsets) == len(acct_keys) + 1
    root = np.empty(32, np.uint8)
    ua, us = _Updates(), _Updates()
    rc = lib().orc_state_root_full(_p(acct_keys), _p(accounts), len(acct_keys), _p(slot_keys), _p(values),
                                   _p(seg_offsets), _p(root),
                                   C.byref(ua) if want_updates else None,
                                   C.byref(us) if want_updates else None, threads)
    if rc:
        raise ValueError(f"orc_state_root_full rc={rc}")
    if want_updates:
        return root.tobytes(), _updates_to_py(ua), _updates_to_py(us)
    return root.tobytes()


def trie_root_recursive(keys, values: list) -> bytes:
    """Independent recursive implementation: keys uint8[n,32] sorted, values list of bytes."""
    keys = _c(keys).reshape(-1, 32)
    offs = np.zeros(len(values) + 1, np.uint64)
    offs[1:] = np.cumsum([len(v) for v in values])
    data = np.frombuffer(b"".join(values) or b"\0", np.uint8).copy()
    root = np.empty(32, np.uint8)
    rc = lib().orc_trie_root_recursive(_p(keys), _p(data), _p(offs), len(keys), _p(root))
    if rc:
        raise ValueError("orc_trie_root_recursive: keys not strictly increasing")
    return root.tobytes()


def stats_reset():
    lib().orc_stats_reset()


def stats() -> dict:
    s = _Stats()
    lib().orc_stats_get(C.byref(s))
    return {n: getattr(s, n) for n, _ in _Stats._fields_}
