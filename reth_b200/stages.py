"""Host-side mirror of the three hashing/merkle stages (crates/stages/stages/src/stages/{hashing_account,
hashing_storage,merkle}.rs) over in-memory tables.  The MDBX tables, changesets, ETL files and the pipeline are
out of scope; what is mirrored is each stage's full-pass data transformation:

  AccountHashingStage : PlainAccountState[address]           -> HashedAccounts[keccak(address)]      (sorted)
  StorageHashingStage : PlainStorageState[address][slot]     -> HashedStorages[keccak(addr)][keccak(slot)] (sorted)
  MerkleStage         : HashedAccounts + HashedStorages      -> state root (+ AccountsTrie/StoragesTrie updates),
                        validated against the header's state root (merkle.rs:437-453)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .engine import Engine
from .hashed_state import Account, HashedPostStateSorted, HashedStorageSorted
from .trie import StateRoot, TrieUpdates


class StageError(RuntimeError):
    pass


@dataclass
class Tables:
    plain_accounts: Dict[bytes, Account] = field(default_factory=dict)        # PlainAccountState
    plain_storage: Dict[bytes, Dict[int, int]] = field(default_factory=dict)  # PlainStorageState (slot -> value)
    hashed_accounts: List[Tuple[bytes, Account]] = field(default_factory=list)          # HashedAccounts, key order
    hashed_storages: Dict[bytes, List[Tuple[bytes, int]]] = field(default_factory=dict)  # HashedStorages, key order
    trie_updates: Optional[TrieUpdates] = None                                # AccountsTrie + StoragesTrie rows


class AccountHashingStage:
    """Full pass of hashing_account.rs:176-238: hash every address, emit rows sorted by digest (the ETL
    collector's job, done by the device radix sort)."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def execute(self, t: Tables) -> int:
        addrs = list(t.plain_accounts.keys())
        if not addrs:
            t.hashed_accounts = []
            return 0
        arr = np.frombuffer(b"".join(addrs), np.uint8).reshape(len(addrs), 20)
        digests, perm = self.engine.hash_sort_keys(arr)
        t.hashed_accounts = [(digests[i].tobytes(), t.plain_accounts[addrs[int(perm[i])]]) for i in range(len(addrs))]
        return len(addrs)


class StorageHashingStage:
    """Full pass of hashing_storage.rs:106-178: composite key keccak(address) ‖ keccak(slot); the address digest
    is computed once per address (the reference caches it across consecutive entries, :129-134)."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def execute(self, t: Tables) -> int:
        addrs = [a for a, st in t.plain_storage.items() if st]
        t.hashed_storages = {}
        if not addrs:
            return 0
        slot_rows, owners, values = [], [], []
        for i, a in enumerate(addrs):
            for slot, val in t.plain_storage[a].items():
                if val == 0:
                    continue  # zero-valued slots are not stored
                slot_rows.append(int(slot).to_bytes(32, "big"))
                owners.append(i)
                values.append(val)
        total = len(slot_rows)
        if total:
            # one device call: hash each address once, each slot once, sort by keccak(address) || keccak(slot)
            keys, perm = self.engine.hash_sort_storage(
                np.frombuffer(b"".join(addrs), np.uint8).reshape(len(addrs), 20), np.array(owners, np.uint32),
                np.frombuffer(b"".join(slot_rows), np.uint8).reshape(total, 32))
            for i in range(total):  # rows arrive in table order (append_dup, hashing_storage.rs:150-170)
                t.hashed_storages.setdefault(keys[i, :32].tobytes(), []).append(
                    (keys[i, 32:].tobytes(), values[int(perm[i])]))
        return total


class MerkleStage:
    """Rebuild path of merkle.rs:210-310: StateRoot over the hashed tables with updates retained, then
    validate_state_root against the expected header root."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def execute(self, t: Tables, expected_state_root: Optional[bytes] = None) -> bytes:
        state = HashedPostStateSorted(list(t.hashed_accounts),
                                      {k: HashedStorageSorted(list(v)) for k, v in t.hashed_storages.items()})
        root, updates = StateRoot(self.engine, state).root_with_updates()
        if expected_state_root is not None and root != expected_state_root:
            # merkle.rs:437-453 -> StageError::Block{BodyStateRootDiff}
            raise StageError(f"state root mismatch: got {root.hex()}, expected {expected_state_root.hex()}")
        t.trie_updates = updates
        return root
