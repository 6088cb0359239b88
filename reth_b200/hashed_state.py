"""Host-side mirror of reth's hashed-state types (crates/trie/common/src/hashed_state.rs, prefix_set.rs, key.rs).

Same names, argument meaning and deletion semantics as the reference; every keccak goes to the device in one
batch (`Engine.keccak256_fixed`) instead of one `keccak256` call per key.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from .engine import ACCOUNT_DTYPE, KECCAK_EMPTY, Engine

B256 = bytes


def unpack_nibbles(key: bytes) -> bytes:
    """Nibbles::unpack."""
    return bytes(x for b in key for x in (b >> 4, b & 15))


@dataclass(frozen=True)
class Account:
    """reth_primitives_traits::Account (nonce, balance, bytecode_hash: None = no code)."""
    nonce: int = 0
    balance: int = 0
    bytecode_hash: Optional[bytes] = None

    def code_hash(self) -> bytes:
        # into_trie_account: bytecode_hash.unwrap_or(KECCAK_EMPTY), crates/trie/common/src/account.rs:16-31
        return self.bytecode_hash or KECCAK_EMPTY


class KeccakKeyHasher:
    """crates/trie/common/src/key.rs:4-18, plus the batch form the device wants."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def hash_key(self, key: bytes) -> bytes:
        return self.hash_keys([key])[0]

    def hash_keys(self, keys: List[bytes]) -> List[bytes]:
        if not keys:
            return []
        ln = len(keys[0])
        if any(len(k) != ln for k in keys):
            raise ValueError("hash_keys: all keys of a batch must have the same length")
        arr = np.frombuffer(b"".join(keys), np.uint8).reshape(len(keys), ln)
        return [d.tobytes() for d in self.engine.keccak256_fixed(arr)]


# ---------------------------------------------------------------------------------------------- prefix sets
class PrefixSetMut:
    """crates/trie/common/src/prefix_set.rs:100-177."""

    def __init__(self, keys: Iterable[bytes] = (), all: bool = False):
        self.all = all
        self.keys: List[bytes] = list(keys)

    @classmethod
    def all_(cls):
        return cls(all=True)

    def insert(self, nibbles: bytes):
        self.keys.append(bytes(nibbles))

    def extend(self, other: "PrefixSetMut"):
        self.all |= other.all
        self.keys.extend(other.keys)

    def is_empty(self) -> bool:
        return not self.all and not self.keys

    def freeze(self) -> "PrefixSet":
        if self.all:
            return PrefixSet([], all=True)
        return PrefixSet(sorted(set(self.keys)))


class PrefixSet:
    """Frozen, sorted, de-duplicated; `contains` keeps a cursor like the reference (:205-231)."""

    def __init__(self, keys: List[bytes], all: bool = False):
        self.keys, self.all, self.index = keys, all, 0

    def contains(self, prefix: bytes) -> bool:
        if self.all:
            return True
        while self.index > 0 and self.keys[self.index] > prefix:
            self.index -= 1
        for idx in range(self.index, len(self.keys)):
            key = self.keys[idx]
            if key.startswith(prefix):
                self.index = idx
                return True
            if key > prefix:
                self.index = idx
                return False
        return False

    def __len__(self):
        return len(self.keys)

    def is_empty(self) -> bool:
        return not self.all and not self.keys


@dataclass
class TriePrefixSetsMut:
    account_prefix_set: PrefixSetMut = field(default_factory=PrefixSetMut)
    storage_prefix_sets: Dict[B256, PrefixSetMut] = field(default_factory=dict)
    destroyed_accounts: set = field(default_factory=set)

    def freeze(self) -> "TriePrefixSets":
        return TriePrefixSets(self.account_prefix_set.freeze(),
                              {k: v.freeze() for k, v in self.storage_prefix_sets.items()},
                              set(self.destroyed_accounts))


@dataclass
class TriePrefixSets:
    account_prefix_set: PrefixSet = field(default_factory=lambda: PrefixSet([]))
    storage_prefix_sets: Dict[B256, PrefixSet] = field(default_factory=dict)
    destroyed_accounts: set = field(default_factory=set)


# ---------------------------------------------------------------------------------------------- hashed state
@dataclass
class HashedStorage:
    """hashed_state.rs:423-428: wiped flag + hashed slot -> value (0 = deleted)."""
    wiped: bool = False
    storage: Dict[B256, int] = field(default_factory=dict)

    def is_empty(self) -> bool:
        return not self.wiped and not self.storage

    @classmethod
    def from_iter(cls, wiped: bool, items: Iterable[Tuple[B256, int]]):
        return cls(wiped, dict(items))

    def construct_prefix_set(self) -> PrefixSetMut:
        if self.wiped:
            return PrefixSetMut.all_()
        return PrefixSetMut(unpack_nibbles(k) for k in self.storage)

    def extend(self, other: "HashedStorage"):
        if other.wiped:
            self.wiped = True
            self.storage.clear()
        self.storage.update(other.storage)

    def into_sorted(self) -> "HashedStorageSorted":
        return HashedStorageSorted(sorted(self.storage.items()), self.wiped)


@dataclass
class HashedStorageSorted:
    """hashed_state.rs:710-715: storage_slots sorted by hashed slot; value 0 = deletion."""
    storage_slots: List[Tuple[B256, int]] = field(default_factory=list)
    wiped: bool = False


@dataclass
class HashedPostState:
    """hashed_state.rs:29-34: hashed address -> Account | None (destroyed); hashed address -> HashedStorage."""
    accounts: Dict[B256, Optional[Account]] = field(default_factory=dict)
    storages: Dict[B256, HashedStorage] = field(default_factory=dict)

    @classmethod
    def from_bundle_state(cls, engine: Engine, state: Iterable[Tuple[bytes, dict]]) -> "HashedPostState":
        """from_bundle_state (hashed_state.rs:49-69).  `state`: (address20, {"info": Account|None,
        "was_destroyed": bool, "storage": {slot_int: present_value_int}}).  All addresses are hashed in one
        device batch, all slots in another — the reference hashes them one by one (and in parallel with rayon,
        crates/trie/db/src/state.rs:408)."""
        state = list(state)
        hasher = KeccakKeyHasher(engine)
        hashed_addrs = hasher.hash_keys([bytes(a) for a, _ in state])
        slot_keys, owners = [], []
        for i, (_, acc) in enumerate(state):
            for slot in acc.get("storage", {}):
                slot_keys.append(int(slot).to_bytes(32, "big"))
                owners.append(i)
        hashed_slots = hasher.hash_keys(slot_keys)
        res = cls()
        per_acc: Dict[int, Dict[bytes, int]] = {}
        for hs, owner, sk in zip(hashed_slots, owners, slot_keys):
            per_acc.setdefault(owner, {})[hs] = int(state[owner][1]["storage"][int.from_bytes(sk, "big")])
        for i, (_, acc) in enumerate(state):
            res.accounts[hashed_addrs[i]] = acc.get("info")
            hs = HashedStorage(bool(acc.get("was_destroyed", False)), per_acc.get(i, {}))
            if not hs.is_empty():
                res.storages[hashed_addrs[i]] = hs
        return res

    @classmethod
    def from_hashed_storage(cls, hashed_address: B256, storage: HashedStorage):
        return cls({}, {hashed_address: storage})

    def with_accounts(self, accounts: Iterable[Tuple[B256, Optional[Account]]]):
        self.accounts = dict(accounts)
        return self

    def with_storages(self, storages: Iterable[Tuple[B256, HashedStorage]]):
        self.storages = dict(storages)
        return self

    def is_empty(self) -> bool:
        return not self.accounts and not self.storages

    def construct_prefix_sets(self) -> TriePrefixSetsMut:
        """hashed_state.rs:105-126."""
        ps = TriePrefixSetsMut()
        for addr, acc in self.accounts.items():
            ps.account_prefix_set.insert(unpack_nibbles(addr))
            if acc is None:
                ps.destroyed_accounts.add(addr)
        for addr, st in self.storages.items():
            ps.account_prefix_set.insert(unpack_nibbles(addr))
            ps.storage_prefix_sets[addr] = st.construct_prefix_set()
        return ps

    def extend(self, other: "HashedPostState"):
        self.accounts.update(other.accounts)
        for addr, st in other.storages.items():
            if addr in self.storages:
                self.storages[addr].extend(st)
            else:
                self.storages[addr] = HashedStorage(st.wiped, dict(st.storage))

    def into_sorted(self) -> "HashedPostStateSorted":
        """hashed_state.rs:329-340."""
        return HashedPostStateSorted(sorted(self.accounts.items()),
                                     {k: v.into_sorted() for k, v in self.storages.items()})


@dataclass
class HashedPostStateSorted:
    """hashed_state.rs:519-524."""
    accounts: List[Tuple[B256, Optional[Account]]] = field(default_factory=list)
    storages: Dict[B256, HashedStorageSorted] = field(default_factory=dict)

    @classmethod
    def from_reverts(cls, engine: Engine, account_changesets, storage_changesets) -> "HashedPostStateSorted":
        """DatabaseHashedPostState::from_reverts (crates/trie/db/src/state.rs:289-347): the state to go back to when the blocks
        of a range are unwound.  account_changesets: (address20, Account-before or None) rows in changeset order (block, then
        address); storage_changesets: (address20, slot (int or 32 bytes), value-before) rows likewise.  The value before the
        FIRST change of every address / (address, slot) pair is kept, addresses and slots are keccak-hashed, everything comes
        back in trie order — one b200_hash_changesets call instead of the HashSets and the sorts on one core."""
        acct = [(bytes(a), info) for a, info in account_changesets]
        stor = [(bytes(a), sl if isinstance(sl, (bytes, bytearray)) else int(sl).to_bytes(32, "big"), int(v))
                for a, sl, v in storage_changesets]
        blob = lambda rows, w: np.frombuffer(b"".join(rows), np.uint8).reshape(-1, w) if rows else np.zeros((0, w), np.uint8)
        cs = engine.hash_changesets(blob([a for a, _ in acct], 20), blob([a for a, _, _ in stor], 20),
                                    blob([bytes(sl) for _, sl, _ in stor], 32))
        accounts = [(k.tobytes(), acct[int(i)][1]) for k, i in zip(cs["account_keys"], cs["account_first"])]
        offs = cs["storage_seg_offsets"]
        storages = {}
        for t, hk in enumerate(cs["storage_account_keys"]):
            seg = range(int(offs[t]), int(offs[t + 1]))
            storages[hk.tobytes()] = HashedStorageSorted([(cs["slot_keys"][j].tobytes(), stor[int(cs["slot_first"][j])][2]) for j in seg],
                                                         False)
        return cls(accounts, storages)

    def to_flat(self):
        """The flat layout of include/b200trie.h for a state that IS the whole state (no database underneath):
        destroyed accounts (None) and zero-valued slots are dropped exactly where the reference's cursors skip
        them (crates/trie/trie/src/hashed_cursor/post_state.rs:260-297); storage of an address without an
        account entry is never visited by StateRoot::calculate and is dropped too."""
        live = [(k, a) for k, a in self.accounts if a is not None]
        n = len(live)
        keys = np.frombuffer(b"".join(k for k, _ in live), np.uint8).reshape(n, 32) if n else np.zeros((0, 32), np.uint8)
        accts = np.zeros(n, ACCOUNT_DTYPE)
        sk, sv, offs = [], [], [0]
        for i, (k, a) in enumerate(live):
            accts[i]["nonce"] = a.nonce
            accts[i]["balance"] = np.frombuffer(int(a.balance).to_bytes(32, "big"), np.uint8)
            accts[i]["code_hash"] = np.frombuffer(a.code_hash(), np.uint8)
            st = self.storages.get(k)
            cnt = 0
            if st is not None:
                for slot, val in st.storage_slots:
                    if val != 0:
                        sk.append(slot)
                        sv.append(int(val).to_bytes(32, "big"))
                        cnt += 1
            offs.append(offs[-1] + cnt)
        m = len(sk)
        slot_keys = np.frombuffer(b"".join(sk), np.uint8).reshape(m, 32) if m else np.zeros((0, 32), np.uint8)
        slot_vals = np.frombuffer(b"".join(sv), np.uint8).reshape(m, 32) if m else np.zeros((0, 32), np.uint8)
        return keys, accts, slot_keys, slot_vals, np.array(offs, np.uint64)
