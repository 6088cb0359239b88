"""Host-side mirror of the three hashing/merkle stages (crates/stages/stages/src/stages/{hashing_account,
hashing_storage,merkle}.rs) over in-memory tables.  The MDBX tables, changesets, ETL files and the pipeline are
out of scope; what is mirrored is each stage's full-pass data transformation:

  AccountHashingStage : PlainAccountState[address]           -> HashedAccounts[keccak(address)]      (sorted)
  StorageHashingStage : PlainStorageState[address][slot]     -> HashedStorages[keccak(addr)][keccak(slot)] (sorted)
  MerkleStage         : HashedAccounts + HashedStorages      -> state root (+ AccountsTrie/StoragesTrie updates),
                        validated against the header's state root (merkle.rs:437-453)

and the incremental legs the pipeline takes for short block ranges (merkle.rs:255-300 with the incremental hashing of
crates/storage/provider/src/providers/database/provider.rs:3206-3221,3266-3280), in two forms:

  * nothing resident — the pipeline's own shape: `AccountHashingStage.execute_incremental` / `StorageHashingStage.
    execute_incremental` take the range's changeset rows (b200_hash_changesets: every address and slot hashed once, first
    occurrence kept) and bring the hashed tables up to date; `MerkleStage.execute_incremental_from_changesets` turns the same
    rows into prefix sets (load_prefix_sets_with_provider) and folds the changed leaves with the stored hashes of every
    untouched subtree (IncrementalStateRoot over b200_root_from_items);
  * state resident on the device — `MerkleStage.execute_incremental`: the changed plain accounts / slots of the range are
    hashed in one device batch, folded into the hashed tables and committed to a `DynamicStateRoot`.

Either way the root is validated before the trie tables are written, the way write_trie_updates does it.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .engine import Engine
from .hashed_state import (Account, HashedPostState, HashedPostStateSorted, HashedStorage, HashedStorageSorted, KeccakKeyHasher,
                           PrefixSet, TriePrefixSets, unpack_nibbles)
from .trie import DynamicStateRoot, StateRoot, StorageTrieUpdates, TrieUpdates


class StageError(RuntimeError):
    pass


@dataclass
class Tables:
    plain_accounts: Dict[bytes, Account] = field(default_factory=dict)        # PlainAccountState
    plain_storage: Dict[bytes, Dict[int, int]] = field(default_factory=dict)  # PlainStorageState (slot -> value)
    hashed_accounts: List[Tuple[bytes, Account]] = field(default_factory=list)          # HashedAccounts, key order
    hashed_storages: Dict[bytes, List[Tuple[bytes, int]]] = field(default_factory=dict)  # HashedStorages, key order
    trie_updates: Optional[TrieUpdates] = None                                # AccountsTrie + StoragesTrie rows
    merkle_checkpoint: Optional[Tuple[bytes, bytes]] = None   # (last_account_key, stream checkpoint): MerkleCheckpoint
    merkle_inner_state: Optional[object] = None               # the live IntermediateStateRootState between execute() calls


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


    def execute_incremental(self, t: Tables, changed_addresses) -> int:
        """The stage's changeset leg (hashing_account.rs:240-262 -> insert_account_for_hashing, provider.rs:3206-3221):
        `changed_addresses` are the addresses of the range's AccountChangeSets rows (repeats allowed); every distinct one is
        hashed once on the device and HashedAccounts takes the account's CURRENT plain value — or loses the row."""
        addrs = [bytes(a) for a in changed_addresses]
        if not addrs:
            return 0
        cs = self.engine.hash_changesets(np.frombuffer(b"".join(addrs), np.uint8).reshape(-1, 20),
                                         np.zeros((0, 20), np.uint8), np.zeros((0, 32), np.uint8))
        hashed = dict(t.hashed_accounts)
        for k, first in zip(cs["account_keys"], cs["account_first"]):
            acc = t.plain_accounts.get(addrs[int(first)])
            if acc is None:
                hashed.pop(k.tobytes(), None)
            else:
                hashed[k.tobytes()] = acc
        t.hashed_accounts = sorted(hashed.items())
        return len(cs["account_keys"])


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


    def execute_incremental(self, t: Tables, changed_rows) -> int:
        """The changeset leg (hashing_storage.rs:180-206 -> insert_storage_for_hashing, provider.rs:3266-3280): `changed_rows` =
        (address, slot) of the range's StorageChangeSets rows in changeset order (repeats allowed; a destroyed account lists
        every slot it had).  One device call hashes each address once per run and each slot once and returns the unique pairs
        as a CSR; HashedStorages takes every pair's CURRENT plain value, zero / absent = the row goes."""
        rows = [(bytes(a), int(s)) for a, s in changed_rows]
        if not rows:
            return 0
        sa = np.frombuffer(b"".join(a for a, _ in rows), np.uint8).reshape(-1, 20)
        ss = np.frombuffer(b"".join(s.to_bytes(32, "big") for _, s in rows), np.uint8).reshape(-1, 32)
        cs = self.engine.hash_changesets(np.zeros((0, 20), np.uint8), sa, ss)
        offs = cs["storage_seg_offsets"]
        for i, hk in enumerate(cs["storage_account_keys"]):
            cur = dict(t.hashed_storages.get(hk.tobytes(), []))
            for j in range(int(offs[i]), int(offs[i + 1])):
                addr, slot = rows[int(cs["slot_first"][j])]
                val = t.plain_storage.get(addr, {}).get(slot, 0)
                if val:
                    cur[cs["slot_keys"][j].tobytes()] = val
                else:
                    cur.pop(cs["slot_keys"][j].tobytes(), None)
            if cur:
                t.hashed_storages[hk.tobytes()] = sorted(cur.items())
            else:
                t.hashed_storages.pop(hk.tobytes(), None)
        return len(cs["slot_keys"])


class MerkleStage:
    """Rebuild path of merkle.rs:210-310: StateRoot over the hashed tables with updates retained, then
    validate_state_root against the expected header root."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.resident: Optional[DynamicStateRoot] = None   # seeded by the first incremental execution

    def execute(self, t: Tables, expected_state_root: Optional[bytes] = None, threshold: Optional[int] = None,
                max_steps: Optional[int] = None) -> Optional[bytes]:
        """Rebuild leg (merkle.rs:210-310).  With `threshold` the build runs in ranges of at least that many hashed entries:
        every step's TrieUpdates are written to the trie tables and the inner checkpoint (`t.merkle_checkpoint`, the role of
        MerkleCheckpoint / save_execution_checkpoint, merkle.rs:118-148,255-275) is saved, exactly the loop the pipeline
        drives by calling execute() until it reports done.  `max_steps` stops after that many ranges and returns None
        (ExecOutput{done: false}); the next execute() continues from the saved checkpoint."""
        self.close()  # a rebuild invalidates the resident state
        state = HashedPostStateSorted(list(t.hashed_accounts),
                                      {k: HashedStorageSorted(list(v)) for k, v in t.hashed_storages.items()})
        if threshold is None:
            root, updates = StateRoot(self.engine, state).root_with_updates()
            if expected_state_root is not None and root != expected_state_root:
                # merkle.rs:437-453 -> StageError::Block{BodyStateRootDiff}
                raise StageError(f"state root mismatch: got {root.hex()}, expected {expected_state_root.hex()}")
            t.trie_updates = updates
            t.merkle_checkpoint = None
            return root
        inter = getattr(t, "merkle_inner_state", None)
        if inter is None:  # "Rebuilding trie": reset the checkpoint and clear the trie tables (merkle.rs:229-237)
            t.trie_updates = TrieUpdates()
            t.merkle_checkpoint = None
        steps = 0
        while True:
            progress = StateRoot(self.engine, state).with_threshold(threshold).with_intermediate_state(inter) \
                .root_with_progress()
            self.write_trie_updates(t.trie_updates, progress.updates)
            steps += 1
            if progress.complete:
                t.merkle_inner_state = None
                t.merkle_checkpoint = None
                if expected_state_root is not None and progress.root != expected_state_root:
                    raise StageError(f"state root mismatch: got {progress.root.hex()}, expected {expected_state_root.hex()}")
                return progress.root
            inter = progress.state
            t.merkle_inner_state = inter
            t.merkle_checkpoint = (inter.last_hashed_key, inter.checkpoint())   # what would be persisted
            if max_steps is not None and steps >= max_steps:
                return None

    # ------------------------------------------------------------------ incremental leg
    def execute_incremental(self, t: Tables, changed_accounts: Dict[bytes, Optional[Account]],
                            changed_storage: Dict[bytes, Dict[int, int]], wiped: Optional[set] = None,
                            expected_state_root: Optional[bytes] = None) -> bytes:
        """changed_accounts: plain address -> new Account (None = destroyed); changed_storage: plain address -> {slot:
        new value (0 = cleared)}; wiped: addresses whose storage is wiped before the changes apply.  Updates the plain and
        hashed tables, commits to the resident state and applies the TrieUpdates to `t.trie_updates`."""
        wiped = wiped or set()
        if t.trie_updates is None:
            raise StageError("no trie tables: run the rebuild (execute) once before incremental executions")
        if self.resident is None:  # seed the device state from the hashed tables as they stand BEFORE this range
            seed = HashedPostStateSorted(list(t.hashed_accounts),
                                         {k: HashedStorageSorted(list(v)) for k, v in t.hashed_storages.items()})
            self.resident = DynamicStateRoot(self.engine, seed)
        # 1. incremental hashing: every changed address and slot in one device batch each
        hasher = KeccakKeyHasher(self.engine)
        addrs = sorted(set(changed_accounts) | set(changed_storage) | set(wiped))
        ha = dict(zip(addrs, hasher.hash_keys(addrs)))
        slot_list = sorted({int(s) for st in changed_storage.values() for s in st})
        hs = dict(zip(slot_list, hasher.hash_keys([s.to_bytes(32, "big") for s in slot_list])))
        post = HashedPostState()
        for a, acc in changed_accounts.items():
            post.accounts[ha[a]] = acc
        for a in set(changed_storage) | set(wiped):
            post.storages[ha[a]] = HashedStorage(a in wiped, {hs[int(s)]: int(v) for s, v in changed_storage.get(a, {}).items()})
        # 2. plain + hashed tables, built aside: they replace the live ones only after the root has been validated
        # (merkle.rs:437-453 validates before the transaction commits; a mismatch drops it)
        plain_accounts = dict(t.plain_accounts)
        plain_storage = dict(t.plain_storage)
        hashed_storages = dict(t.hashed_storages)
        hashed_accounts = dict(t.hashed_accounts)
        for a, acc in changed_accounts.items():
            if acc is None:
                plain_accounts.pop(a, None)
                plain_storage.pop(a, None)
                hashed_accounts.pop(ha[a], None)
                hashed_storages.pop(ha[a], None)
            else:
                plain_accounts[a] = acc
                hashed_accounts[ha[a]] = acc
        for a in set(changed_storage) | set(wiped):
            if changed_accounts.get(a, 0) is None:
                continue
            plain = {} if a in wiped else dict(plain_storage.get(a, {}))
            hashed = {} if a in wiped else dict(hashed_storages.get(ha[a], []))
            for s, v in changed_storage.get(a, {}).items():
                if v == 0:
                    plain.pop(int(s), None)
                    hashed.pop(hs[int(s)], None)
                else:
                    plain[int(s)] = int(v)
                    hashed[hs[int(s)]] = int(v)
            plain_storage[a] = plain
            if hashed:
                hashed_storages[ha[a]] = sorted(hashed.items())
            else:
                hashed_storages.pop(ha[a], None)
        # 3. commit on the device, validate, then publish the tables and write the trie tables
        root, upd = self.resident.commit(post)
        if expected_state_root is not None and root != expected_state_root:
            # the device state has advanced past what the tables hold: drop it, the next execution re-seeds from the tables
            self.close()
            raise StageError(f"state root mismatch: got {root.hex()}, expected {expected_state_root.hex()}")
        t.plain_accounts, t.plain_storage = plain_accounts, plain_storage
        t.hashed_storages = hashed_storages
        t.hashed_accounts = sorted(hashed_accounts.items())
        self.write_trie_updates(t.trie_updates, upd)
        return root

    def execute_incremental_from_changesets(self, t: Tables, changed_addresses, changed_rows,
                                            expected_state_root: Optional[bytes] = None) -> bytes:
        """The incremental leg as the pipeline runs it (merkle.rs:255-300): nothing resident on the device.  The range's
        changesets become prefix sets in one device call (load_prefix_sets_with_provider, crates/trie/db/src/prefix_set.rs:
        22-60 -> b200_hash_changesets), the hashed tables — already brought up to date by the two hashing stages — supply the
        leaves, the trie tables the hashes of every subtree the prefix sets do not touch (StateRoot::incremental_root_with_
        updates -> IncrementalStateRoot over b200_root_from_items).  The root is validated before the trie tables are
        written."""
        from .walker import IncrementalStateRoot
        if t.trie_updates is None:
            raise StageError("no trie tables: run the rebuild (execute) once before incremental executions")
        self.close()
        addrs = [bytes(a) for a in changed_addresses]
        rows = [(bytes(a), int(sl)) for a, sl in changed_rows]
        cs = self.engine.hash_changesets(
            np.frombuffer(b"".join(addrs), np.uint8).reshape(-1, 20) if addrs else np.zeros((0, 20), np.uint8),
            np.frombuffer(b"".join(a for a, _ in rows), np.uint8).reshape(-1, 20) if rows else np.zeros((0, 20), np.uint8),
            np.frombuffer(b"".join(sl.to_bytes(32, "big") for _, sl in rows), np.uint8).reshape(-1, 32) if rows
            else np.zeros((0, 32), np.uint8))
        live = dict(t.hashed_accounts)
        offs = cs["storage_seg_offsets"]
        prefix_sets = TriePrefixSets(
            PrefixSet([unpack_nibbles(k.tobytes()) for k in cs["account_prefix_keys"]]),
            {hk.tobytes(): PrefixSet([unpack_nibbles(cs["slot_keys"][j].tobytes()) for j in range(int(offs[i]), int(offs[i + 1]))])
             for i, hk in enumerate(cs["storage_account_keys"])},
            # destroyed_accounts: changed addresses that have no HashedAccounts row any more (prefix_set.rs:44-51)
            {k.tobytes() for k in cs["account_keys"] if k.tobytes() not in live})
        state = HashedPostStateSorted(list(t.hashed_accounts),
                                      {k: HashedStorageSorted(list(v)) for k, v in t.hashed_storages.items()})
        root, upd = IncrementalStateRoot(self.engine, t.trie_updates, state, prefix_sets).root_with_updates()
        if expected_state_root is not None and root != expected_state_root:
            raise StageError(f"state root mismatch: got {root.hex()}, expected {expected_state_root.hex()}")
        self.write_trie_updates(t.trie_updates, upd)
        return root

    @staticmethod
    def write_trie_updates(tables: TrieUpdates, upd: TrieUpdates):
        """write_trie_updates_sorted (crates/storage/provider/src/providers/database/provider.rs:3125-3160,
        crates/trie/db/src/trie_cursor.rs:280-312): removed paths deleted, updated nodes upserted; a deleted storage trie
        loses all its rows first."""
        for p in upd.removed_nodes:
            tables.account_nodes.pop(p, None)
        tables.account_nodes.update(upd.account_nodes)
        for addr, st in upd.storage_tries.items():
            cur = tables.storage_tries.get(addr)
            if st.is_deleted or cur is None:
                cur = StorageTrieUpdates()
            for p in st.removed_nodes:
                cur.storage_nodes.pop(p, None)
            cur.storage_nodes.update(st.storage_nodes)
            if cur.storage_nodes:
                tables.storage_tries[addr] = cur
            else:
                tables.storage_tries.pop(addr, None)

    def close(self):
        if self.resident is not None:
            self.resident.close()
            self.resident = None
