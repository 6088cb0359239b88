"""Trie-table verifier: the host-side mirror of reth's `Verifier` (crates/trie/trie/src/verify.rs:150-186,
`reth db repair-trie`).  The reference recomputes the stored branch nodes from the hashed tables with a `StateRoot` walk
and streams the differences against `AccountsTrie` / `StoragesTrie`; here the recomputation is one `b200_state_root_full`
with updates retained, the comparison is a dictionary diff.  Output records carry reth's names."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Union

from .engine import Engine
from .hashed_state import HashedPostStateSorted
from .trie import BranchNodeCompact, StateRoot, TrieUpdates

B256 = bytes
Nibbles = bytes


@dataclass(frozen=True)
class Extra:
    """AccountExtra / StorageExtra: a stored node that the state does not produce."""
    account: Optional[B256]  # None: the accounts trie
    path: Nibbles
    node: BranchNodeCompact


@dataclass(frozen=True)
class Wrong:
    """AccountWrong / StorageWrong."""
    account: Optional[B256]
    path: Nibbles
    expected: BranchNodeCompact
    found: BranchNodeCompact


@dataclass(frozen=True)
class Missing:
    """AccountMissing / StorageMissing: a node the state produces that is not stored."""
    account: Optional[B256]
    path: Nibbles
    node: BranchNodeCompact


Output = Union[Extra, Wrong, Missing]


class Verifier:
    def __init__(self, engine: Engine, hashed_state: HashedPostStateSorted):
        self.engine, self.state = engine, hashed_state

    def verify(self, tables: TrieUpdates) -> List[Output]:
        """tables: the stored trie tables (account_nodes, storage_tries[addr].storage_nodes).  -> inconsistencies in
        reth's order: accounts trie first, then storage tries by hashed address, paths ascending inside each."""
        _, expected = StateRoot(self.engine, self.state).root_with_updates()
        out: List[Output] = []
        self._diff(None, expected.account_nodes, tables.account_nodes, out)
        exp_st = {k: v.storage_nodes for k, v in expected.storage_tries.items() if v.storage_nodes}
        got_st = {k: v.storage_nodes for k, v in tables.storage_tries.items() if v.storage_nodes}
        for addr in sorted(set(exp_st) | set(got_st)):
            self._diff(addr, exp_st.get(addr, {}), got_st.get(addr, {}), out)
        return out

    @staticmethod
    def _diff(account, expected, found, out):
        for path in sorted(set(expected) | set(found)):
            e, f = expected.get(path), found.get(path)
            if e is None:
                out.append(Extra(account, path, f))
            elif f is None:
                out.append(Missing(account, path, e))
            elif e != f:
                out.append(Wrong(account, path, e, f))
