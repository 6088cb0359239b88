"""Shared helpers for the parity tests: synthetic states in the flat layout of include/b200trie.h."""
from __future__ import annotations

import numpy as np

import oracle

MASK64 = (1 << 64) - 1


def splitmix64_stream(seed: int, n_words: int) -> np.ndarray:
    """Vectorised splitmix64: word i is the i-th output of the generator seeded with `seed`."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n_words + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_keys(seed: int, n: int) -> np.ndarray:
    """n x 32 pseudo-random bytes (stand-ins for keccak outputs)."""
    return splitmix64_stream(seed, 4 * n).view(np.uint8).reshape(n, 32).copy()


def sort_rows(keys: np.ndarray) -> np.ndarray:
    """argsort of uint8[n,32] rows in big-endian lexicographic order."""
    v = keys.reshape(-1, 32).view(">u8")  # [n,4]
    return np.lexsort((v[:, 3], v[:, 2], v[:, 1], v[:, 0]))


def u256_be(values) -> np.ndarray:
    vals = list(values)
    out = np.zeros((len(vals), 32), np.uint8)
    for i, v in enumerate(vals):
        out[i] = np.frombuffer(int(v).to_bytes(32, "big"), np.uint8)
    return out


def alloc_to_flat(alloc: dict, keccak_rows=None):
    """Genesis-style alloc {addr_hex: {balance, nonce?, code?, storage?}} -> flat sorted arrays.

    Mirrors alloy-trie `state_root_ref_unhashed` as called at crates/chainspec/src/spec.rs:98: keys are
    keccak(address) / keccak(slot), zero-valued slots are skipped, code_hash = keccak(code).
    Returns (acct_keys[n,32], accounts[n], slot_keys[m,32], slot_values[m,32], seg_offsets[n+1]).
    """
    kr = keccak_rows or (lambda a: oracle.keccak256_fixed(a))
    addrs = sorted(alloc.keys())
    n = len(addrs)
    addr_bytes = np.zeros((n, 20), np.uint8)
    for i, a in enumerate(addrs):
        addr_bytes[i] = np.frombuffer(bytes.fromhex(a), np.uint8)
    hashed = kr(addr_bytes) if n else np.zeros((0, 32), np.uint8)
    order = sort_rows(hashed) if n else np.zeros(0, np.int64)
    rows, slot_keys, slot_vals, offs = [], [], [], [0]
    for i in order:
        e = alloc[addrs[i]]
        code = bytes.fromhex(e.get("code", "0x")[2:]) if e.get("code") else b""
        code_hash = oracle.keccak256(code) if code else None
        rows.append((int(e.get("nonce", "0x0"), 16), int(e.get("balance", "0x0"), 16), code_hash))
        st = e.get("storage") or {}
        items = [(bytes.fromhex(k[2:].rjust(64, "0")), int(v, 16)) for k, v in st.items()]
        items = [(k, v) for k, v in items if v != 0]
        if items:
            sk = np.frombuffer(b"".join(k for k, _ in items), np.uint8).reshape(-1, 32)
            hk = kr(sk)
            o2 = sort_rows(hk)
            slot_keys.append(hk[o2])
            slot_vals.append(u256_be([items[j][1] for j in o2]))
        offs.append(offs[-1] + len(items))
    acct_keys = hashed[order] if n else hashed
    accounts = oracle.make_accounts(rows)
    sk = np.concatenate(slot_keys) if slot_keys else np.zeros((0, 32), np.uint8)
    sv = np.concatenate(slot_vals) if slot_vals else np.zeros((0, 32), np.uint8)
    return acct_keys, accounts, sk, sv, np.array(offs, np.uint64)


def synth_accounts(seed: int, n: int, with_code=True):
    """Sorted random account keys + accounts (nonce<2^16, balance<2^80, random code hash)."""
    keys = random_keys(seed, n)
    keys = keys[sort_rows(keys)]
    w = splitmix64_stream(seed ^ 0xACC0, 8 * n).reshape(n, 8)
    acc = np.zeros(n, oracle.ACCOUNT_DTYPE)
    acc["nonce"] = w[:, 0] & np.uint64(0xFFFF)
    bal = np.zeros((n, 32), np.uint8)
    bal[:, 22:24] = (w[:, 1] & np.uint64(0xFFFF)).astype(">u2").view(np.uint8).reshape(n, 2)
    bal[:, 24:32] = w[:, 2].astype(">u8").view(np.uint8).reshape(n, 8)
    acc["balance"] = bal
    if with_code:
        acc["code_hash"] = w[:, 4:8].copy().view(np.uint8).reshape(n, 32)
    else:
        acc["code_hash"] = np.frombuffer(oracle.KECCAK_EMPTY, np.uint8)
    return keys, acc


def synth_storage(seed: int, counts, value_mode="u64"):
    """Per-account sorted random slot keys and non-zero values. counts: slots per account."""
    counts = np.asarray(counts, np.int64)
    m = int(counts.sum())
    offs = np.zeros(len(counts) + 1, np.uint64)
    offs[1:] = np.cumsum(counts)
    keys = random_keys(seed ^ 0x5107, m)
    seg = np.repeat(np.arange(len(counts)), counts)
    v = keys.view(">u8")
    order = np.lexsort((v[:, 3], v[:, 2], v[:, 1], v[:, 0], seg))
    keys = keys[order]
    w = splitmix64_stream(seed ^ 0x7A1, 5 * m).reshape(m, 5)
    vals = np.zeros((m, 32), np.uint8)
    if value_mode == "u64":
        x = w[:, 0] | np.uint64(1)
        vals[:, 24:32] = x.astype(">u8").view(np.uint8).reshape(m, 8)
    else:  # "mixed": 50% < 2^8, 25% 20-byte, 25% full 32-byte
        sel = w[:, 4] % np.uint64(4)
        full = w[:, :4].copy().astype(">u8").view(np.uint8).reshape(m, 32)
        full[:, 31] |= 1
        vals[:] = full
        small = sel < 2
        vals[small, :31] = 0
        addr = sel == 2
        vals[addr, :12] = 0
        vals[addr, 12] |= 1
        vals[sel == 3, 0] |= 1
    return keys, vals, offs


def to_device_ptrs(arrays):
    """Raw device addresses of copies of `arrays` (+ the objects that keep them alive): CUDA tensors on a GPU; under
    tools/emu device memory is host memory, so the numpy arrays themselves."""
    import os
    arrays = [np.ascontiguousarray(a) for a in arrays]
    if not os.environ.get("B200_EMU"):
        import torch
        if torch.cuda.is_available():
            hold = [torch.from_numpy(a.view(np.uint8).reshape(-1)).cuda() for a in arrays]
            return [t.data_ptr() for t in hold], hold
    return [a.ctypes.data for a in arrays], arrays


def read_device(held, dtype=np.uint8):
    """Host copy of one of the objects `to_device_ptrs` returned in its second list."""
    if hasattr(held, "cpu"):
        held = held.cpu().numpy()
    return np.asarray(held).view(np.uint8).view(dtype).copy()
