"""Host mirror of reth's incremental walk — TrieWalker + TrieNodeIter (crates/trie/trie/src/walker.rs:161-388,
node_iter.rs:200-304) — and of the incremental StateRoot / StorageRoot built on it (trie.rs:160-330,615-721), with the
HashBuilder fold replaced by one device call (b200_root_from_items).

The walk is cursor work over the trie tables: it descends the stored branch nodes only where the prefix set says a key
changed (PrefixSet::contains, crates/trie/common/src/prefix_set.rs:205-231; walker.rs:161-202), yields the stored hash of
every child it may skip (`TrieElement::Branch`) and leaves for everything else (`TrieElement::Leaf`), and remembers every
stored node it descended into as removed (walker.rs:336-344) — the nodes the fold re-creates take precedence
(updates.rs:160-167).  In a reth integration this is reth's own walker; this module is its restatement over plain
dictionaries for the host mirror and the tests.  Paths are bytes of nibbles.
"""
from __future__ import annotations

from bisect import bisect_left, bisect_right
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

from .engine import ACCOUNT_DTYPE, EMPTY_ROOT_HASH, Engine
from .hashed_state import HashedPostStateSorted, PrefixSet, TriePrefixSets

Nibbles = bytes


def unpack(key: bytes) -> Nibbles:
    return bytes(x for b in key for x in (b >> 4, b & 15))


def pack_padded(path: Nibbles, fill: int) -> bytes:
    nibs = list(path) + [fill] * (64 - len(path))
    return bytes((nibs[2 * i] << 4) | nibs[2 * i + 1] for i in range(32))


@dataclass
class TrieElement:
    """node_iter.rs:16-27: Branch(TrieBranchNode{key, value: hash, children_are_in_trie}) | Leaf(key, value index)."""
    path: Nibbles           # 64 nibbles for a leaf
    hash: Optional[bytes]   # stored hash of the skipped subtree (None: leaf)
    children_are_in_trie: bool = False
    leaf_index: int = -1    # position in the sorted leaf list

    @property
    def is_leaf(self) -> bool:
        return self.hash is None


def walk(stored: Dict[Nibbles, object], changes: PrefixSet, leaf_keys: Sequence[bytes]) -> Tuple[List[TrieElement], Set[Nibbles]]:
    """The element stream of one trie, in key order, and the stored node paths the walk descended into.

    stored: path -> BranchNodeCompact (the rows of AccountsTrie, or of one account's StoragesTrie; the root node is never
    stored); changes: the frozen prefix set; leaf_keys: the packed 32-byte keys of the trie's CURRENT leaves, ascending (what
    the hashed cursor over the post-state yields)."""
    paths = sorted(stored)
    skips: List[Tuple[Nibbles, bytes, bool]] = []
    removed: Set[Nibbles] = set()

    def visit(p: Nibbles):
        removed.add(p)                      # consume_node with deletions retained (walker.rs:336-344)
        node = stored[p]
        hi = 0
        for c in range(16):
            bit = 1 << c
            if not node.state_mask & bit:
                continue
            pc = p + bytes([c])
            has_hash = bool(node.hash_mask & bit)
            h = node.hashes[hi] if has_hash else None
            hi += has_hash
            if has_hash and not changes.contains(pc):          # update_skip_node (walker.rs:172-202)
                skips.append((pc, h, bool(node.tree_mask & bit)))
            elif node.tree_mask & bit:                          # children are in the trie: consume the next stored node
                j = bisect_left(paths, pc)
                if j < len(paths) and paths[j].startswith(pc):
                    visit(paths[j])

    i = 0
    while i < len(paths):                   # stored nodes without a stored ancestor (the root branch is never stored)
        q = paths[i]
        visit(q)
        i += 1
        while i < len(paths) and paths[i].startswith(q):
            i += 1
    out: List[TrieElement] = []
    pos = 0
    for path, h, in_trie in skips:          # ascending: children are visited in nibble order
        a = bisect_left(leaf_keys, pack_padded(path, 0), pos)
        for li in range(pos, a):
            out.append(TrieElement(unpack(leaf_keys[li]), None, False, li))
        out.append(TrieElement(path, h, in_trie))
        pos = bisect_right(leaf_keys, pack_padded(path, 15), a)
    for li in range(pos, len(leaf_keys)):
        out.append(TrieElement(unpack(leaf_keys[li]), None, False, li))
    return out, removed


def _items_arrays(elements: List[TrieElement], row: int):
    n = len(elements)
    keys = np.zeros((n, 32), np.uint8)
    nibs = np.zeros(n, np.uint8)
    flags = np.zeros(n, np.uint8)
    vals = np.zeros((n, row), np.uint8)
    for i, e in enumerate(elements):
        keys[i] = np.frombuffer(pack_padded(e.path, 0), np.uint8)
        nibs[i] = len(e.path)
        if not e.is_leaf:
            flags[i] = 1 if e.children_are_in_trie else 0
            vals[i, :32] = np.frombuffer(e.hash, np.uint8)
    return keys, nibs, flags, vals


class IncrementalStateRoot:
    """StateRoot::new(trie tables, post-state).with_prefix_sets(..).root_with_updates() when the trie is not resident on the
    device: `tables` holds the stored nodes (TrieUpdates-shaped: account_nodes, storage_tries[addr].storage_nodes), `state`
    the complete hashed state AFTER the change (the hashed cursor factory), `prefix_sets` what changed
    (HashedPostState::construct_prefix_sets / load_prefix_sets).  Only the paths the prefix sets touch are re-hashed; every
    other subtree enters through its stored hash."""

    def __init__(self, engine: Engine, tables, state: HashedPostStateSorted, prefix_sets: TriePrefixSets):
        self.engine, self.tables, self.state, self.prefix_sets = engine, tables, state, prefix_sets

    def root(self) -> bytes:
        return self.root_with_updates()[0]

    def root_with_updates(self):
        from .trie import BranchNodeCompact, StorageTrieUpdates, TrieUpdates
        eng = self.engine
        akeys, accts, skeys, svals, offs = self.state.to_flat()
        acct_key_list = [akeys[i].tobytes() for i in range(len(akeys))]
        elements, removed_acct = walk(self.tables.account_nodes, self.prefix_sets.account_prefix_set, acct_key_list)
        self.hashed_entries_walked = sum(e.is_leaf for e in elements)
        # ---- storage roots of every account leaf the walk yields (trie.rs:262-292): one forest call
        leaf_accounts = [e.leaf_index for e in elements if e.is_leaf]
        st_elems, st_offs, st_removed, st_owner = [], [0], {}, []
        for ai in leaf_accounts:
            addr = acct_key_list[ai]
            s0, s1 = int(offs[ai]), int(offs[ai + 1])
            slot_keys = [skeys[j].tobytes() for j in range(s0, s1)]
            stored = self.tables.storage_tries[addr].storage_nodes if addr in self.tables.storage_tries else {}
            changes = self.prefix_sets.storage_prefix_sets.get(addr, PrefixSet([]))
            els, rem = walk(stored, changes, slot_keys)
            for e in els:
                if e.is_leaf:
                    e.leaf_index += s0
            self.hashed_entries_walked += sum(e.is_leaf for e in els)
            st_elems.extend(els)
            st_offs.append(len(st_elems))
            st_removed[addr] = rem
            st_owner.append(addr)
        updates = TrieUpdates()
        sroot_of: Dict[int, bytes] = {}
        if st_owner:
            k, nb, fl, vals = _items_arrays(st_elems, 32)
            for i, e in enumerate(st_elems):
                if e.is_leaf:
                    vals[i] = svals[e.leaf_index]
            roots, recs = eng.root_from_items(k, nb, fl, vals, None, np.array(st_offs, np.uint64), account=False, want_updates=True)
            per_trie: Dict[int, Dict[Nibbles, BranchNodeCompact]] = {}
            for tid, path, sm, tm, hm, hashes in recs:
                per_trie.setdefault(tid, {})[bytes(path)] = BranchNodeCompact(sm, tm, hm, tuple(hashes))
            for t, addr in enumerate(st_owner):
                ai = leaf_accounts[t]
                sroot_of[ai] = roots[t].tobytes()
                nodes = per_trie.get(t, {})
                if offs[ai + 1] == offs[ai]:
                    # empty storage: StorageRoot::calculate short-circuits to deleted() (trie.rs:622-629)
                    if addr in self.tables.storage_tries or addr in self.prefix_sets.storage_prefix_sets:
                        updates.insert_storage_updates(addr, StorageTrieUpdates.deleted())
                    continue
                su = StorageTrieUpdates(storage_nodes=nodes, removed_nodes={p for p in st_removed[addr] if p not in nodes})
                updates.insert_storage_updates(addr, su)
        # ---- the account trie
        k, nb, fl, vals = _items_arrays(elements, 72)
        sroots = np.zeros((len(elements), 32), np.uint8)
        for i, e in enumerate(elements):
            if e.is_leaf:
                vals[i] = np.frombuffer(accts[e.leaf_index].tobytes(), np.uint8)
                sroots[i] = np.frombuffer(sroot_of.get(e.leaf_index, EMPTY_ROOT_HASH), np.uint8)
        roots, recs = eng.root_from_items(k, nb, fl, vals, sroots, None, account=True, want_updates=True)
        for _, path, sm, tm, hm, hashes in recs:
            updates.account_nodes[bytes(path)] = BranchNodeCompact(sm, tm, hm, tuple(hashes))
        updates.removed_nodes = {p for p in removed_acct if p not in updates.account_nodes}
        # TrieUpdates::finalize (updates.rs:140-158): destroyed accounts -> is_deleted
        for destroyed in self.prefix_sets.destroyed_accounts:
            updates.storage_tries.setdefault(destroyed, StorageTrieUpdates()).is_deleted = True
        return roots[0].tobytes(), updates
