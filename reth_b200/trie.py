"""Host-side mirror of reth's root calculators (crates/trie/trie/src/trie.rs, crates/trie/parallel/src/root.rs,
crates/trie/common/src/updates.rs) over the C ABI.

The calculators take the hashed state itself where reth takes cursor factories over it: with no stored trie nodes
underneath (from-scratch build: MerkleStage's rebuild path `merkle.rs:210-254`, `StateRootProvider::state_root`
on a full state, every test that uses `MockHashedCursorFactory` + `NoopTrieCursor`) the walk degenerates to
"stream all leaves in key order" (SURVEY.md §3.2), which is what the device consumes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .engine import EMPTY_ROOT_HASH, Engine
from .hashed_state import HashedPostStateSorted, HashedStorageSorted, TriePrefixSets

B256 = bytes
Nibbles = bytes


@dataclass(frozen=True)
class BranchNodeCompact:
    """alloy_trie::BranchNodeCompact as reth stores it (crates/storage/db-api/src/tables/mod.rs:484-494)."""
    state_mask: int
    tree_mask: int
    hash_mask: int
    hashes: Tuple[bytes, ...]
    root_hash: Optional[bytes] = None


@dataclass
class StorageTrieUpdates:
    """crates/trie/common/src/updates.rs:235-245."""
    is_deleted: bool = False
    storage_nodes: Dict[Nibbles, BranchNodeCompact] = field(default_factory=dict)
    removed_nodes: set = field(default_factory=set)

    @classmethod
    def deleted(cls):
        return cls(is_deleted=True)

    def is_empty(self) -> bool:
        return not self.is_deleted and not self.storage_nodes and not self.removed_nodes

    def __len__(self):
        return int(self.is_deleted) + len(self.storage_nodes) + len(self.removed_nodes)


@dataclass
class TrieUpdates:
    """crates/trie/common/src/updates.rs:17-26."""
    account_nodes: Dict[Nibbles, BranchNodeCompact] = field(default_factory=dict)
    removed_nodes: set = field(default_factory=set)
    storage_tries: Dict[B256, StorageTrieUpdates] = field(default_factory=dict)

    def insert_storage_updates(self, hashed_address: B256, updates: StorageTrieUpdates):
        if updates.is_empty():  # updates.rs:132-134
            return
        assert hashed_address not in self.storage_tries
        self.storage_tries[hashed_address] = updates

    def is_empty(self) -> bool:
        return not self.account_nodes and not self.removed_nodes and not self.storage_tries


@dataclass
class IntermediateStateRootState:
    """crates/trie/trie/src/progress.rs:24-30.  reth keeps the HashBuilder stack and the walker position; here the open
    right edge of the build is a b200_root_stream (frontier of the closed top-nibble buckets + the accounts of the open
    bucket in HBM) and the position is the index of the next account of the sorted state."""
    stream: object            # engine.RootStream
    next_account: int
    last_hashed_key: bytes

    def checkpoint(self) -> bytes:
        """What MerkleStage persists between runs (MerkleCheckpoint, merkle.rs:118-148): 1104 bytes."""
        return self.stream.checkpoint()


@dataclass
class StateRootProgress:
    """crates/trie/trie/src/progress.rs:12-21: `Complete(root, walked, updates)` (complete = True, state = None) or
    `Progress(state, walked, updates)` (complete = False, root = None; `updates` are the nodes finished so far by this call)."""
    root: Optional[bytes]
    hashed_entries_walked: int
    updates: TrieUpdates
    complete: bool = True
    state: Optional[IntermediateStateRootState] = None


class StateRootError(RuntimeError):
    """StateRootError::Database(DatabaseError::Other(msg)) in the Rust shim."""


def _records_to_nodes(records) -> Dict[int, Dict[Nibbles, BranchNodeCompact]]:
    out: Dict[int, Dict[Nibbles, BranchNodeCompact]] = {}
    for tid, path, sm, tm, hm, hashes in records:
        out.setdefault(tid, {})[bytes(path)] = BranchNodeCompact(sm, tm, hm, tuple(hashes))
    return out


class StorageRoot:
    """StorageRoot::new_hashed(..).root() / root_with_updates() — crates/trie/trie/src/trie.rs:479-615."""

    def __init__(self, engine: Engine, hashed_address: B256, storage: HashedStorageSorted):
        self.engine, self.hashed_address, self.storage = engine, hashed_address, storage
        self.prefix_set = None
        self.threshold = 100_000

    def with_prefix_set(self, prefix_set):
        """Kept for interface parity (trie.rs:512-516).  This calculator is given every slot of the trie and hashes all of
        them, so a changed-key set cannot alter its result; the skipping that prefix sets drive in reth happens in
        IncrementalStateRoot (stored nodes + prefix sets) and in the resident paths."""
        self.prefix_set = prefix_set
        return self

    def with_threshold(self, threshold: int):
        """trie.rs:523-527.  One storage trie is one device build (a 10M-slot trie takes milliseconds): `calculate`
        always returns the Complete variant, whatever the threshold — the thresholded, resumable build is StateRoot's
        (account ranges; every storage trie of a range is finished inside its push)."""
        self.threshold = threshold
        return self

    def with_no_threshold(self):
        self.threshold = 2**64 - 1
        return self

    def _flat(self):
        slots = [(k, v) for k, v in self.storage.storage_slots if v != 0]
        m = len(slots)
        keys = np.frombuffer(b"".join(k for k, _ in slots), np.uint8).reshape(m, 32) if m else np.zeros((0, 32), np.uint8)
        vals = np.frombuffer(b"".join(int(v).to_bytes(32, "big") for _, v in slots), np.uint8).reshape(m, 32) \
            if m else np.zeros((0, 32), np.uint8)
        return keys, vals, np.array([0, m], np.uint64)

    def root(self) -> bytes:
        return self.calculate(False)[0]

    def root_with_updates(self) -> Tuple[bytes, int, StorageTrieUpdates]:
        return self.calculate(True)

    def calculate(self, retain_updates: bool) -> Tuple[bytes, int, StorageTrieUpdates]:
        """-> (root, storage_slots_walked, updates) like StorageRootProgress::Complete (trie.rs:615-721)."""
        keys, vals, offs = self._flat()
        if len(keys) == 0:  # trie.rs:622-629
            return EMPTY_ROOT_HASH, 0, StorageTrieUpdates.deleted()
        try:
            if retain_updates:
                roots, recs = self.engine.storage_roots(keys, vals, offs, want_updates=True)
                upd = StorageTrieUpdates(storage_nodes=_records_to_nodes(recs).get(0, {}))
            else:
                roots, upd = self.engine.storage_roots(keys, vals, offs), StorageTrieUpdates()
        except Exception as e:  # noqa: BLE001 - mapped like the shim maps native errors
            raise StateRootError(str(e)) from e
        return roots[0].tobytes(), len(keys), upd


class StateRoot:
    """StateRoot::{root, root_with_updates, root_with_progress} — crates/trie/trie/src/trie.rs:54-158.

    with_threshold(t) + root_with_progress(): the build stops after a range of accounts holding at least `t` hashed
    entries (accounts + slots; reth counts retained trie updates, trie.rs:296-306 — a device build knows the entries of a
    range before it runs, the updates only afterwards) and returns StateRootProgress(complete=False, state=...); feeding the
    state back through with_intermediate_state() continues where it stopped.  root() / root_with_updates() ignore the
    threshold like the reference (trie.rs:126-140)."""

    def __init__(self, engine: Engine, hashed_state: HashedPostStateSorted):
        self.engine, self.state = engine, hashed_state
        self.prefix_sets = TriePrefixSets()
        self.threshold = 100_000  # DEFAULT_INTERMEDIATE_THRESHOLD, trie.rs:25
        self.previous_state: Optional[IntermediateStateRootState] = None

    def with_prefix_sets(self, prefix_sets: TriePrefixSets):
        """A from-scratch build hashes every leaf it is given, so the changed-key sets have nothing to skip; only
        `destroyed_accounts` is consumed (TrieUpdates::finalize).  The skip logic lives in IncrementalStateRoot (stored
        nodes + prefix sets) and in the resident paths."""
        self.prefix_sets = prefix_sets
        return self

    def with_threshold(self, threshold: int):
        self.threshold = threshold
        return self

    def with_no_threshold(self):
        self.threshold = 2**64 - 1
        return self

    def with_intermediate_state(self, state: Optional[IntermediateStateRootState]):
        self.previous_state = state
        return self

    def root(self) -> bytes:
        return self._calculate(False).root

    def root_with_updates(self) -> Tuple[bytes, TrieUpdates]:
        p = self._calculate(True)
        return p.root, p.updates

    def root_with_progress(self) -> StateRootProgress:
        if self.threshold >= 2**64 - 1 and self.previous_state is None:
            return self._calculate(True)
        return self._calculate_range()

    def _finalize(self, updates: TrieUpdates):
        # TrieUpdates::finalize (updates.rs:140-158): destroyed accounts -> is_deleted
        for destroyed in self.prefix_sets.destroyed_accounts:
            updates.storage_tries.setdefault(destroyed, StorageTrieUpdates()).is_deleted = True

    def _calculate(self, retain_updates: bool) -> StateRootProgress:
        keys, accts, skeys, svals, offs = self.state.to_flat()
        try:
            if retain_updates:
                root, acct_recs, stor_recs = self.engine.state_root_full(keys, accts, skeys, svals, offs,
                                                                         want_updates=True)
            else:
                root = self.engine.state_root_full(keys, accts, skeys, svals, offs)
                acct_recs = stor_recs = []
        except Exception as e:  # noqa: BLE001
            raise StateRootError(str(e)) from e
        updates = TrieUpdates()
        if retain_updates:
            updates.account_nodes = _records_to_nodes(acct_recs).get(0, {})
            per_trie = _records_to_nodes(stor_recs)
            for i in range(len(keys)):
                addr = keys[i].tobytes()
                if offs[i + 1] == offs[i]:
                    # StorageRoot::calculate returns StorageTrieUpdates::deleted() for empty storage (trie.rs:622-629)
                    updates.insert_storage_updates(addr, StorageTrieUpdates.deleted())
                else:
                    updates.insert_storage_updates(addr, StorageTrieUpdates(storage_nodes=per_trie.get(i, {})))
            self._finalize(updates)
        walked = int(len(keys) + len(skeys))
        return StateRootProgress(root, walked, updates)

    def _calculate_range(self) -> StateRootProgress:
        """One step of the thresholded build: push the next range of accounts into the stream; Complete when none is left."""
        from .engine import RootStream
        keys, accts, skeys, svals, offs = self.state.to_flat()
        n = len(keys)
        st = self.previous_state
        try:
            if st is None:
                st = IntermediateStateRootState(RootStream(self.engine, retain_updates=True), 0, b"")
            a0 = st.next_account
            # the range: accounts a0 .. a1 holding >= threshold hashed entries (at least one account)
            target = int(offs[a0]) + a0 + min(self.threshold, 2**62) if a0 < n else 0
            entries = offs[a0:n + 1].astype(np.int64) + np.arange(a0, n + 1)      # entries before account i
            a1 = min(n, max(a0 + 1, int(np.searchsorted(entries, target, side="left")))) if a0 < n else n
            updates = TrieUpdates()
            walked = 0
            if a1 > a0:
                s0, s1 = int(offs[a0]), int(offs[a1])
                _, acct_recs, stor_recs = st.stream.push(keys[a0:a1], accts[a0:a1], skeys[s0:s1], svals[s0:s1],
                                                         (offs[a0:a1 + 1] - offs[a0]).astype(np.uint64))
                for _, path, sm, tm, hm, hashes in acct_recs:
                    updates.account_nodes[bytes(path)] = BranchNodeCompact(sm, tm, hm, tuple(hashes))
                per_trie = _records_to_nodes(stor_recs)
                for i in range(a0, a1):
                    addr = keys[i].tobytes()
                    if offs[i + 1] == offs[i]:
                        updates.insert_storage_updates(addr, StorageTrieUpdates.deleted())
                    else:
                        updates.insert_storage_updates(addr, StorageTrieUpdates(storage_nodes=per_trie.get(i - a0, {})))
                walked = (a1 - a0) + (s1 - s0)
                st.next_account = a1
                st.last_hashed_key = keys[a1 - 1].tobytes()
            if a1 < n:
                return StateRootProgress(None, walked, updates, complete=False, state=st)
            root, acct_recs = st.stream.finish()
            st.stream.close()
            for _, path, sm, tm, hm, hashes in acct_recs:
                updates.account_nodes[bytes(path)] = BranchNodeCompact(sm, tm, hm, tuple(hashes))
            self._finalize(updates)
            return StateRootProgress(root, walked, updates)
        except StateRootError:
            raise
        except Exception as e:  # noqa: BLE001
            raise StateRootError(str(e)) from e


class ParallelStateRoot(StateRoot):
    """ParallelStateRoot::{incremental_root, incremental_root_with_updates} — crates/trie/parallel/src/root.rs:35-77.
    On the device the storage-root fan-out and the account fold are the same launches."""

    def incremental_root(self) -> bytes:
        return self.root()

    def incremental_root_with_updates(self) -> Tuple[bytes, TrieUpdates]:
        return self.root_with_updates()


class ResidentStateRoot:
    """Live-path commitment with the account trie resident in HBM (BASELINE config 5).

    Plays the part of `StateRoot::overlay_root_with_updates` / `ParallelStateRoot::incremental_root_with_updates`
    (crates/trie/db/src/state.rs:184-230, crates/trie/parallel/src/root.rs:35-77): `commit(HashedPostState)` folds one
    block's hashed post state into the committed state and returns the new root.  The hashed tables the reference
    reads through cursors (`HashedAccounts`, `HashedStorages`) are kept here as host dictionaries; the trie itself —
    every node hash of every level — lives on the device and only the dirty paths are re-hashed
    (`b200_trie_apply`).  Storage roots of the touched accounts are recomputed from their complete post-block storage
    in one `b200_storage_roots` call."""

    def __init__(self, engine: Engine, state: HashedPostStateSorted, dynamic: bool = False):
        """dynamic=True keeps the account trie in a `DynamicTrie` (b200_dtrie_*): new and destroyed accounts are applied in
        place instead of through merge + rebuild (`commit` then always reports rebuilt=False)."""
        from .engine import ACCOUNT_DTYPE, DynamicTrie, ResidentTrie
        self.engine = engine
        self.accounts = {k: a for k, a in state.accounts if a is not None}
        self.storages = {k: {s: v for s, v in st.storage_slots if v != 0} for k, st in state.storages.items()
                         if k in self.accounts}
        keys, accts, skeys, svals, offs = state.to_flat()
        sroots = engine.storage_roots(skeys, svals, offs) if len(keys) else np.zeros((0, 32), np.uint8)
        self.dynamic = dynamic
        self.trie = (DynamicTrie if dynamic else ResidentTrie).create(engine, keys, accts, sroots)
        self._dtype = ACCOUNT_DTYPE

    def root(self) -> bytes:
        return self.trie.root()

    def commit(self, post) -> Tuple[bytes, bool]:
        """post: HashedPostState.  -> (new root, rebuilt) where rebuilt tells whether the trie shape changed."""
        from .hashed_state import Account
        touched = set(post.accounts) | set(post.storages)
        # 1. storage overlay: wiped hides everything older, zero deletes (hashed_cursor/post_state.rs:185-195,260-297)
        for addr, hs in post.storages.items():
            cur = {} if hs.wiped else dict(self.storages.get(addr, {}))
            for slot, val in hs.storage.items():
                if val == 0:
                    cur.pop(slot, None)
                else:
                    cur[slot] = val
            self.storages[addr] = cur
        # 2. account overlay
        for addr, acc in post.accounts.items():
            if acc is None:
                self.accounts.pop(addr, None)
                self.storages.pop(addr, None)
            else:
                self.accounts[addr] = acc
        dirty = sorted(touched)
        if not dirty:
            return self.trie.root(), False
        live = [k for k in dirty if k in self.accounts]
        # 3. storage roots of the live touched accounts, all tries in one device call
        slot_keys, slot_vals, offs = [], [], [0]
        for k in live:
            st = sorted(self.storages.get(k, {}).items())
            slot_keys += [s for s, _ in st]
            slot_vals += [int(v).to_bytes(32, "big") for _, v in st]
            offs.append(offs[-1] + len(st))
        m = len(slot_keys)
        sk = np.frombuffer(b"".join(slot_keys), np.uint8).reshape(m, 32) if m else np.zeros((0, 32), np.uint8)
        sv = np.frombuffer(b"".join(slot_vals), np.uint8).reshape(m, 32) if m else np.zeros((0, 32), np.uint8)
        roots = self.engine.storage_roots(sk, sv, np.array(offs, np.uint64)) if live else np.zeros((0, 32), np.uint8)
        root_of = {k: roots[i] for i, k in enumerate(live)}
        # 4. one sorted dirty set for the device: upserts and deletes
        keys = np.frombuffer(b"".join(dirty), np.uint8).reshape(len(dirty), 32)
        accts = np.zeros(len(dirty), self._dtype)
        present = np.zeros(len(dirty), np.uint8)
        sroots = np.zeros((len(dirty), 32), np.uint8)
        for i, k in enumerate(dirty):
            a = self.accounts.get(k)
            if a is None:
                continue  # destroyed (or storage of an account that does not exist): delete / no-op
            present[i] = 1
            accts[i]["nonce"] = a.nonce
            accts[i]["balance"] = np.frombuffer(int(a.balance).to_bytes(32, "big"), np.uint8)
            accts[i]["code_hash"] = np.frombuffer(a.code_hash(), np.uint8)
            sroots[i] = root_of[k]
        try:
            if self.dynamic:
                return self.trie.apply(keys, accts, present, sroots), False
            return self.trie.apply(keys, accts, present, sroots)
        except Exception as e:  # noqa: BLE001
            raise StateRootError(str(e)) from e

    def close(self):
        self.trie.close()


class DynamicStateRoot:
    """Live-path commitment with the WHOLE hashed state resident in HBM (b200_dstate_*): accounts and every storage trie.

    The role of reth's `SparseStateTrie` fed by the state-root task (crates/trie/sparse/src/state.rs,
    crates/engine/tree/src/tree/payload_processor/sparse_trie.rs) and of `StateRoot::overlay_root_with_updates`
    (crates/trie/db/src/state.rs:184-230): `commit(HashedPostState)` applies one block in place — new / changed /
    destroyed accounts, slot writes, zeroed slots, wiped storages — and returns the new root with the block's
    `TrieUpdates` (account_nodes, removed_nodes, storage_tries with is_deleted).  Nothing of the state is kept on the host.
    Emulation-validated; first B200 run pending (include/b200trie.h)."""

    def __init__(self, engine: Engine, state: HashedPostStateSorted, sharded: bool = False):
        """sharded=True: this object is one rank's shard (see reth_b200.sharded.ShardedDynamicStateRoot); `commit`'s root is
        then the shard's own root and the state root comes from the gathered frontiers."""
        from .engine import ACCOUNT_DTYPE, DynamicState
        keys, accts, skeys, svals, offs = state.to_flat()
        self.ds = DynamicState.create(engine, keys, accts, skeys, svals, offs, sharded=sharded)
        self._dtype = ACCOUNT_DTYPE

    def root(self) -> bytes:
        return self.ds.root()

    def commit(self, post) -> Tuple[bytes, TrieUpdates]:
        """post: HashedPostState -> (root, TrieUpdates of the block)."""
        from .engine import DynamicState as DS
        touched = sorted(set(post.accounts) | set(post.storages))
        m = len(touched)
        keys = np.frombuffer(b"".join(touched), np.uint8).reshape(m, 32) if m else np.zeros((0, 32), np.uint8)
        accts = np.zeros(m, self._dtype)
        flags = np.zeros(m, np.uint8)
        sk, sv, offs = [], [], [0]
        for i, k in enumerate(touched):
            hs = post.storages.get(k)
            if k in post.accounts:
                a = post.accounts[k]
                if a is None:
                    offs.append(len(sk))  # destroyed: flags 0, its slots (if any) are irrelevant
                    continue
                flags[i] = DS.EXISTS
                accts[i]["nonce"] = a.nonce
                accts[i]["balance"] = np.frombuffer(int(a.balance).to_bytes(32, "big"), np.uint8)
                accts[i]["code_hash"] = np.frombuffer(a.code_hash(), np.uint8)
            else:
                flags[i] = DS.EXISTS | DS.UNCHANGED  # storage-only change
            if hs is not None:
                if hs.wiped:
                    flags[i] |= DS.WIPED
                for s, v in sorted(hs.storage.items()):
                    sk.append(s)
                    sv.append(int(v).to_bytes(32, "big"))
            offs.append(len(sk))
        skeys = np.frombuffer(b"".join(sk), np.uint8).reshape(-1, 32) if sk else np.zeros((0, 32), np.uint8)
        svals = np.frombuffer(b"".join(sv), np.uint8).reshape(-1, 32) if sv else np.zeros((0, 32), np.uint8)
        try:
            root, au, ar, su, sr, deleted = self.ds.apply(keys, accts, flags, skeys, svals, np.array(offs, np.uint64),
                                                          want_updates=True)
        except Exception as e:  # noqa: BLE001
            raise StateRootError(str(e)) from e
        upd = TrieUpdates()
        upd.account_nodes = _records_to_nodes(au).get(0, {})
        upd.removed_nodes = {bytes(p) for p in ar}
        per_entry = _records_to_nodes(su)
        removed_per_entry: Dict[int, set] = {}
        for entry, p in sr:
            removed_per_entry.setdefault(entry, set()).add(bytes(p))
        for i, k in enumerate(touched):
            st = StorageTrieUpdates(bool(deleted[i]), per_entry.get(i, {}), removed_per_entry.get(i, set()))
            upd.insert_storage_updates(k, st)
        return root, upd

    def close(self):
        self.ds.close()
