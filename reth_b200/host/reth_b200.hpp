// reth_b200.hpp — C++17 host-side mirror of reth's commitment interface over the C ABI (include/b200trie.h).
//
// reth is Rust and this image has no Rust toolchain, so the layer a reth maintainer would write in Rust
//  is written here in C++ with the same type names, method names, argument meaning and error
// behaviour:  HashedPostState / HashedStorage / HashedPostStateSorted (crates/trie/common/src/hashed_state.rs),
// PrefixSetMut / PrefixSet (prefix_set.rs), TrieUpdates / StorageTrieUpdates / BranchNodeCompact (updates.rs),
// KeccakKeyHasher (key.rs), StateRoot / StorageRoot (crates/trie/trie/src/trie.rs), ParallelStateRoot
// (crates/trie/parallel/src/root.rs).  Header-only; link with -lb200trie.  There is no CPU path: Engine's
// constructor throws B200Error when no CUDA device is usable.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200trie.h"

namespace reth_b200 {

using B256 = std::array<uint8_t, 32>;
using Address = std::array<uint8_t, 20>;
using U256 = std::array<uint8_t, 32>;  // big-endian
using Nibbles = std::vector<uint8_t>;  // one nibble per byte

inline const B256 KECCAK_EMPTY = {0xc5, 0xd2, 0x46, 0x01, 0x86, 0xf7, 0x23, 0x3c, 0x92, 0x7e, 0x7d, 0xb2, 0xdc, 0xc7, 0x03, 0xc0,
                                  0xe5, 0x00, 0xb6, 0x53, 0xca, 0x82, 0x27, 0x3b, 0x7b, 0xfa, 0xd8, 0x04, 0x5d, 0x85, 0xa4, 0x70};
inline const B256 EMPTY_ROOT_HASH = {0x56, 0xe8, 0x1f, 0x17, 0x1b, 0xcc, 0x55, 0xa6, 0xff, 0x83, 0x45, 0xe6, 0x92, 0xc0, 0xf8, 0x6e,
                                     0x5b, 0x48, 0xe0, 0x1b, 0x99, 0x6c, 0xad, 0xc0, 0x01, 0x62, 0x2f, 0xb5, 0xe3, 0x63, 0xb4, 0x21};

inline U256 u256_from_u64(uint64_t v) {
    U256 r{};
    for (int i = 0; i < 8; i++) r[31 - i] = (uint8_t)(v >> (8 * i));
    return r;
}
inline bool is_zero(const U256 &v) {
    for (uint8_t b : v)
        if (b) return false;
    return true;
}
inline Nibbles unpack_nibbles(const B256 &k) {
    Nibbles n(64);
    for (int i = 0; i < 32; i++) {
        n[2 * i] = k[i] >> 4;
        n[2 * i + 1] = k[i] & 15;
    }
    return n;
}

/// StateRootError::Database(DatabaseError::Other(msg)) in the Rust shim.
struct B200Error : std::runtime_error {
    int status;
    B200Error(int s, const std::string &m) : std::runtime_error("b200 status " + std::to_string(s) + ": " + m), status(s) {}
};

/// reth_primitives_traits::Account
struct Account {
    uint64_t nonce = 0;
    U256 balance{};
    std::optional<B256> bytecode_hash;
};

/// One b200_ctx (one GPU).  Internally locked by the library; share it freely between threads.
class Engine {
  public:
    explicit Engine(int device = 0) : ctx_(b200_create(device)) {
        if (!ctx_) throw B200Error(b200_create_status(), "b200_create failed (no CUDA device? reth_b200 has no CPU path)");
    }
    ~Engine() { b200_destroy(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    b200_ctx *raw() const { return ctx_; }
    void check(int32_t rc) const {
        if (rc != B200_OK) throw B200Error(rc, b200_last_error(ctx_));
    }
    /// n fixed-length messages -> n digests
    std::vector<B256> keccak256_fixed(const uint8_t *msgs, uint32_t msg_len, uint64_t n) const {
        std::vector<B256> out(n);
        check(b200_keccak256_fixed(ctx_, msgs, msg_len, msg_len, n, n ? out[0].data() : nullptr));
        return out;
    }

  private:
    b200_ctx *ctx_;
};

/// crates/trie/common/src/key.rs:4-18, plus the batch form the device wants
struct KeccakKeyHasher {
    static B256 hash_key(const Engine &e, const uint8_t *bytes, uint32_t len) { return e.keccak256_fixed(bytes, len, 1)[0]; }
    template <size_t N>
    static std::vector<B256> hash_keys(const Engine &e, const std::vector<std::array<uint8_t, N>> &keys) {
        return e.keccak256_fixed(keys.empty() ? nullptr : keys[0].data(), (uint32_t)N, keys.size());
    }
};

// ---------------------------------------------------------------------------------------------- prefix sets
/// crates/trie/common/src/prefix_set.rs:182-231
class PrefixSet {
  public:
    PrefixSet() = default;
    PrefixSet(std::vector<Nibbles> keys, bool all) : all_(all), keys_(std::move(keys)) {}
    bool contains(const Nibbles &prefix) {
        if (all_) return true;
        while (index_ > 0 && keys_[index_] > prefix) index_--;
        for (size_t i = index_; i < keys_.size(); i++) {
            const Nibbles &k = keys_[i];
            if (k.size() >= prefix.size() && std::equal(prefix.begin(), prefix.end(), k.begin())) {
                index_ = i;
                return true;
            }
            if (k > prefix) {
                index_ = i;
                return false;
            }
        }
        return false;
    }
    size_t len() const { return keys_.size(); }
    bool is_all() const { return all_; }

  private:
    bool all_ = false;
    size_t index_ = 0;
    std::vector<Nibbles> keys_;
};

/// prefix_set.rs:100-177
class PrefixSetMut {
  public:
    static PrefixSetMut all() {
        PrefixSetMut p;
        p.all_ = true;
        return p;
    }
    void insert(Nibbles n) { keys_.push_back(std::move(n)); }
    void extend(const PrefixSetMut &o) {
        all_ |= o.all_;
        keys_.insert(keys_.end(), o.keys_.begin(), o.keys_.end());
    }
    bool is_empty() const { return !all_ && keys_.empty(); }
    PrefixSet freeze() const {
        if (all_) return PrefixSet({}, true);
        std::set<Nibbles> s(keys_.begin(), keys_.end());
        return PrefixSet(std::vector<Nibbles>(s.begin(), s.end()), false);
    }

  private:
    bool all_ = false;
    std::vector<Nibbles> keys_;
};

struct TriePrefixSets {
    PrefixSet account_prefix_set;
    std::map<B256, PrefixSet> storage_prefix_sets;
    std::set<B256> destroyed_accounts;
};
struct TriePrefixSetsMut {
    PrefixSetMut account_prefix_set;
    std::map<B256, PrefixSetMut> storage_prefix_sets;
    std::set<B256> destroyed_accounts;
    TriePrefixSets freeze() const {
        TriePrefixSets f;
        f.account_prefix_set = account_prefix_set.freeze();
        for (auto &kv : storage_prefix_sets) f.storage_prefix_sets.emplace(kv.first, kv.second.freeze());
        f.destroyed_accounts = destroyed_accounts;
        return f;
    }
};

// ---------------------------------------------------------------------------------------------- hashed state
/// hashed_state.rs:710-715
struct HashedStorageSorted {
    std::vector<std::pair<B256, U256>> storage_slots;  // sorted by hashed slot; zero value = deletion
    bool wiped = false;
};

/// hashed_state.rs:423-428
struct HashedStorage {
    bool wiped = false;
    std::map<B256, U256> storage;
    bool is_empty() const { return !wiped && storage.empty(); }
    PrefixSetMut construct_prefix_set() const {
        if (wiped) return PrefixSetMut::all();
        PrefixSetMut p;
        for (auto &kv : storage) p.insert(unpack_nibbles(kv.first));
        return p;
    }
    HashedStorageSorted into_sorted() const {
        return HashedStorageSorted{std::vector<std::pair<B256, U256>>(storage.begin(), storage.end()), wiped};
    }
};

/// The flat layout of include/b200trie.h.
struct FlatState {
    std::vector<uint8_t> acct_keys;  // n x 32
    std::vector<b200_account> accts;
    std::vector<uint8_t> slot_keys, slot_values;  // m x 32
    std::vector<uint64_t> seg_offsets{0};
    uint64_t n_accounts() const { return accts.size(); }
};

/// hashed_state.rs:519-524
struct HashedPostStateSorted {
    std::vector<std::pair<B256, std::optional<Account>>> accounts;
    std::map<B256, HashedStorageSorted> storages;

    /// Destroyed accounts (None) and zero-valued slots are dropped where the reference's cursors skip them
    /// (crates/trie/trie/src/hashed_cursor/post_state.rs:260-297); storage without an account entry is never
    /// visited by StateRoot::calculate.
    FlatState to_flat() const {
        FlatState f;
        for (auto &ka : accounts) {
            if (!ka.second) continue;
            const Account &a = *ka.second;
            f.acct_keys.insert(f.acct_keys.end(), ka.first.begin(), ka.first.end());
            b200_account ba;
            ba.nonce = a.nonce;
            std::memcpy(ba.balance_be, a.balance.data(), 32);
            const B256 &ch = a.bytecode_hash ? *a.bytecode_hash : KECCAK_EMPTY;  // account.rs:16-31
            std::memcpy(ba.code_hash, ch.data(), 32);
            f.accts.push_back(ba);
            uint64_t cnt = 0;
            auto it = storages.find(ka.first);
            if (it != storages.end())
                for (auto &sv : it->second.storage_slots)
                    if (!is_zero(sv.second)) {
                        f.slot_keys.insert(f.slot_keys.end(), sv.first.begin(), sv.first.end());
                        f.slot_values.insert(f.slot_values.end(), sv.second.begin(), sv.second.end());
                        cnt++;
                    }
            f.seg_offsets.push_back(f.seg_offsets.back() + cnt);
        }
        return f;
    }
};

/// One account of a BundleState as from_bundle_state sees it.
struct BundleAccount {
    std::optional<Account> info;
    bool was_destroyed = false;
    std::vector<std::pair<U256, U256>> storage;  // slot -> present value
};

/// hashed_state.rs:29-34
struct HashedPostState {
    std::map<B256, std::optional<Account>> accounts;
    std::map<B256, HashedStorage> storages;

    /// from_bundle_state (hashed_state.rs:49-69): all addresses in one device batch, all slots in another.
    static HashedPostState from_bundle_state(const Engine &e, const std::vector<std::pair<Address, BundleAccount>> &state) {
        std::vector<Address> addrs;
        std::vector<U256> slots;
        for (auto &kv : state) {
            addrs.push_back(kv.first);
            for (auto &sv : kv.second.storage) slots.push_back(sv.first);
        }
        auto ha = KeccakKeyHasher::hash_keys(e, addrs);
        auto hs = KeccakKeyHasher::hash_keys(e, slots);
        HashedPostState out;
        size_t si = 0;
        for (size_t i = 0; i < state.size(); i++) {
            const BundleAccount &b = state[i].second;
            out.accounts[ha[i]] = b.info;
            HashedStorage st;
            st.wiped = b.was_destroyed;
            for (auto &sv : b.storage) st.storage[hs[si++]] = sv.second;
            if (!st.is_empty()) out.storages[ha[i]] = std::move(st);
        }
        return out;
    }
    bool is_empty() const { return accounts.empty() && storages.empty(); }
    /// hashed_state.rs:105-126
    TriePrefixSetsMut construct_prefix_sets() const {
        TriePrefixSetsMut ps;
        for (auto &ka : accounts) {
            ps.account_prefix_set.insert(unpack_nibbles(ka.first));
            if (!ka.second) ps.destroyed_accounts.insert(ka.first);
        }
        for (auto &ks : storages) {
            ps.account_prefix_set.insert(unpack_nibbles(ks.first));
            ps.storage_prefix_sets.emplace(ks.first, ks.second.construct_prefix_set());
        }
        return ps;
    }
    /// hashed_state.rs:329-340
    HashedPostStateSorted into_sorted() const {
        HashedPostStateSorted s;
        s.accounts.assign(accounts.begin(), accounts.end());
        for (auto &ks : storages) s.storages.emplace(ks.first, ks.second.into_sorted());
        return s;
    }
};

// ---------------------------------------------------------------------------------------------- updates
struct BranchNodeCompact {
    uint16_t state_mask = 0, tree_mask = 0, hash_mask = 0;
    std::vector<B256> hashes;
    std::optional<B256> root_hash;
    bool operator==(const BranchNodeCompact &o) const {
        return state_mask == o.state_mask && tree_mask == o.tree_mask && hash_mask == o.hash_mask && hashes == o.hashes &&
               root_hash == o.root_hash;
    }
};
/// updates.rs:235-245
struct StorageTrieUpdates {
    bool is_deleted = false;
    std::map<Nibbles, BranchNodeCompact> storage_nodes;
    std::set<Nibbles> removed_nodes;
    static StorageTrieUpdates deleted() {
        StorageTrieUpdates u;
        u.is_deleted = true;
        return u;
    }
    bool is_empty() const { return !is_deleted && storage_nodes.empty() && removed_nodes.empty(); }
    size_t len() const { return (is_deleted ? 1 : 0) + storage_nodes.size() + removed_nodes.size(); }
};
/// updates.rs:17-26
struct TrieUpdates {
    std::map<Nibbles, BranchNodeCompact> account_nodes;
    std::set<Nibbles> removed_nodes;
    std::map<B256, StorageTrieUpdates> storage_tries;
    void insert_storage_updates(const B256 &addr, StorageTrieUpdates u) {
        if (u.is_empty()) return;  // updates.rs:132-134
        storage_tries.emplace(addr, std::move(u));
    }
};

namespace detail {
inline std::pair<Nibbles, BranchNodeCompact> branch_node(const b200_updates &u, uint64_t i) {
    Nibbles path(u.path_len[i]);
    const uint8_t *pp = u.path_packed + 32 * i;
    for (size_t j = 0; j < path.size(); j++) path[j] = (j & 1) ? (pp[j >> 1] & 15) : (pp[j >> 1] >> 4);
    BranchNodeCompact n;
    n.state_mask = u.state_mask[i];
    n.tree_mask = u.tree_mask[i];
    n.hash_mask = u.hash_mask[i];
    for (uint64_t h = u.hash_offset[i]; h < u.hash_offset[i + 1]; h++) {
        B256 x;
        std::memcpy(x.data(), u.hashes + 32 * h, 32);
        n.hashes.push_back(x);
    }
    return {std::move(path), std::move(n)};
}
}  // namespace detail

// ---------------------------------------------------------------------------------------------- table rows
/// One row of AccountsTrie / StoragesTrie (crates/storage/db-api/src/tables/mod.rs:484-494,542-572), byte-exact.
struct TableRow {
    std::vector<uint8_t> key, value;
    bool operator==(const TableRow &o) const { return key == o.key && value == o.value; }
};

namespace detail {
/// std::map<Nibbles, BranchNodeCompact> (+ trie ids) flattened into a b200_updates view.
struct FlatUpdates {
    std::vector<uint32_t> trie_id;
    std::vector<uint8_t> path_len, path_packed, hashes;
    std::vector<uint16_t> state, tree, hash;
    std::vector<uint64_t> hash_offset{0};
    void push(uint32_t id, const Nibbles &path, const BranchNodeCompact &n) {
        trie_id.push_back(id);
        path_len.push_back((uint8_t)path.size());
        size_t base = path_packed.size();
        path_packed.resize(base + 32, 0);
        for (size_t j = 0; j < path.size(); j++) path_packed[base + (j >> 1)] |= (j & 1) ? path[j] : (uint8_t)(path[j] << 4);
        state.push_back(n.state_mask);
        tree.push_back(n.tree_mask);
        hash.push_back(n.hash_mask);
        for (auto &h : n.hashes) hashes.insert(hashes.end(), h.begin(), h.end());
        hash_offset.push_back(hash_offset.back() + n.hashes.size());
    }
    b200_updates view() {
        b200_updates u{};
        u.n_nodes = trie_id.size();
        u.trie_id = trie_id.data();
        u.path_len = path_len.data();
        u.path_packed = path_packed.data();
        u.state_mask = state.data();
        u.tree_mask = tree.data();
        u.hash_mask = hash.data();
        u.hash_offset = hash_offset.data();
        u.hashes = hashes.data();
        return u;
    }
};
inline std::vector<TableRow> take_rows(b200_rows &r) {
    std::vector<TableRow> out(r.n_rows);
    for (uint64_t i = 0; i < r.n_rows; i++) {
        const uint8_t *p = r.bytes + r.row_offset[i], *e = r.bytes + r.row_offset[i + 1];
        out[i].key.assign(p, p + r.key_len[i]);
        out[i].value.assign(p + r.key_len[i], e);
    }
    b200_rows_release(&r);
    return out;
}
inline void check_rows(int32_t rc, const char *what) {
    if (rc != B200_OK) throw B200Error(rc, what);
}
}  // namespace detail

/// Rows `write_trie_updates_sorted` puts into AccountsTrie (provider.rs:3125-3160), in key order.
inline std::vector<TableRow> account_trie_rows(const TrieUpdates &u, b200_key_format fmt = B200_KEYS_LEGACY) {
    detail::FlatUpdates f;
    for (auto &kv : u.account_nodes) f.push(0, kv.first, kv.second);
    b200_updates view = f.view();
    b200_rows rows{};
    detail::check_rows(b200_account_trie_rows(&view, fmt, &rows), "b200_account_trie_rows");
    return detail::take_rows(rows);
}
/// Rows of StoragesTrie for every storage trie with nodes (trie_cursor.rs:280-312), by hashed address then subkey.
/// Tries with `is_deleted` carry no rows of their own: the writer clears their duplicates first (`:285-287`).
inline std::vector<TableRow> storage_trie_rows(const TrieUpdates &u, b200_key_format fmt = B200_KEYS_LEGACY) {
    detail::FlatUpdates f;
    std::vector<uint8_t> addrs;
    uint32_t id = 0;
    for (auto &kv : u.storage_tries) {
        addrs.insert(addrs.end(), kv.first.begin(), kv.first.end());
        for (auto &n : kv.second.storage_nodes) f.push(id, n.first, n.second);
        id++;
    }
    b200_updates view = f.view();
    b200_rows rows{};
    detail::check_rows(b200_storage_trie_rows(&view, addrs.data(), id, fmt, &rows), "b200_storage_trie_rows");
    return detail::take_rows(rows);
}

/// crates/trie/trie/src/progress.rs:24-30.  reth keeps the HashBuilder stack and the walker position; here the open right edge
/// of the build is a b200_root_stream (frontier of the closed top-nibble buckets + the accounts of the open bucket in HBM) and
/// the position is the index of the next account of the sorted state.
struct IntermediateStateRootState {
    b200_root_stream *stream = nullptr;
    uint64_t next_account = 0;
    B256 last_hashed_key{};
    ~IntermediateStateRootState() {
        if (stream) b200_root_stream_free(stream);
    }
    /// what MerkleStage persists between runs (MerkleCheckpoint, merkle.rs:118-148)
    b200_stream_checkpoint checkpoint() const {
        b200_stream_checkpoint cp{};
        if (b200_root_stream_checkpoint(stream, &cp) != B200_OK) throw B200Error(B200_ERR_INVALID_ARG, "stream checkpoint");
        return cp;
    }
};

/// crates/trie/trie/src/progress.rs:12-21: Complete(root, walked, updates) (complete, no state) or
/// Progress(state, walked, updates) (not complete; `updates` = the nodes this call finished).
struct StateRootProgress {
    B256 root{};
    size_t hashed_entries_walked = 0;
    TrieUpdates updates;
    bool complete = true;
    std::shared_ptr<IntermediateStateRootState> state;
};

/// StorageRoot::{root, root_with_updates, calculate} — trie.rs:479-721
class StorageRoot {
  public:
    StorageRoot(const Engine &e, B256 hashed_address, HashedStorageSorted storage)
        : e_(e), hashed_address_(hashed_address), storage_(std::move(storage)) {}
    B256 root() const { return std::get<0>(calculate(false)); }
    std::tuple<B256, size_t, StorageTrieUpdates> root_with_updates() const { return calculate(true); }
    std::tuple<B256, size_t, StorageTrieUpdates> calculate(bool retain_updates) const {
        std::vector<uint8_t> keys, vals;
        for (auto &sv : storage_.storage_slots)
            if (!is_zero(sv.second)) {
                keys.insert(keys.end(), sv.first.begin(), sv.first.end());
                vals.insert(vals.end(), sv.second.begin(), sv.second.end());
            }
        uint64_t m = keys.size() / 32;
        if (m == 0) return {EMPTY_ROOT_HASH, 0, StorageTrieUpdates::deleted()};  // trie.rs:622-629
        uint64_t offs[2] = {0, m};
        B256 root;
        b200_updates u{};
        e_.check(b200_storage_roots(e_.raw(), keys.data(), vals.data(), offs, 1, root.data(), retain_updates ? &u : nullptr,
                                    nullptr));
        StorageTrieUpdates upd;
        if (retain_updates) {
            for (uint64_t i = 0; i < u.n_nodes; i++) upd.storage_nodes.insert(detail::branch_node(u, i));
            b200_updates_release(&u);
        }
        return {root, (size_t)m, std::move(upd)};
    }

  private:
    const Engine &e_;
    B256 hashed_address_;
    HashedStorageSorted storage_;
};

/// StateRoot::{root, root_with_updates, root_with_progress, with_prefix_sets, with_threshold} — trie.rs:54-330.
/// Takes the hashed state where reth takes cursor factories over it (no stored trie nodes underneath:
/// MerkleStage's rebuild path, StateRootProvider::state_root on a full state, MockHashedCursorFactory tests).
class StateRoot {
  public:
    StateRoot(const Engine &e, HashedPostStateSorted state) : e_(e), state_(std::move(state)) {}
    StateRoot &with_prefix_sets(TriePrefixSets ps) {
        prefix_sets_ = std::move(ps);
        return *this;
    }
    StateRoot &with_threshold(uint64_t t) {
        threshold_ = t;
        return *this;
    }
    StateRoot &with_no_threshold() {
        threshold_ = UINT64_MAX;
        return *this;
    }
    /// trie.rs:85-88: continue a thresholded build where root_with_progress stopped
    StateRoot &with_intermediate_state(std::shared_ptr<IntermediateStateRootState> st) {
        previous_state_ = std::move(st);
        return *this;
    }
    /// root() / root_with_updates() ignore the threshold like the reference (trie.rs:126-140)
    B256 root() const { return calculate(false).root; }
    std::pair<B256, TrieUpdates> root_with_updates() const {
        auto p = calculate(true);
        return {p.root, std::move(p.updates)};
    }
    /// With a threshold the build stops after a range of accounts holding at least that many hashed entries (accounts +
    /// slots) and returns Progress; feeding `state` back through with_intermediate_state continues (b200_root_stream_*).
    StateRootProgress root_with_progress() const {
        if (threshold_ == UINT64_MAX && !previous_state_) return calculate(true);
        return calculate_range();
    }
    /// The rebuild leg of MerkleStage in one call (merkle.rs:216-253 → write_trie_updates): the root and the stored nodes
    /// as AccountsTrie / StoragesTrie rows in table order, laid out on the device (b200_state_root_full_rows).
    struct RootWithTables {
        B256 root;
        std::vector<TableRow> accounts_trie, storages_trie;
    };
    RootWithTables root_with_table_rows(b200_key_format fmt = B200_KEYS_LEGACY) const {
        FlatState f = state_.to_flat();
        RootWithTables out;
        b200_rows ar{}, sr{};
        e_.check(b200_state_root_full_rows(e_.raw(), f.acct_keys.data(), f.accts.data(), f.n_accounts(), f.slot_keys.data(),
                                           f.slot_values.data(), f.seg_offsets.data(), fmt, out.root.data(), &ar, &sr, nullptr));
        out.accounts_trie = detail::take_rows(ar);
        out.storages_trie = detail::take_rows(sr);
        return out;
    }

  protected:
    StateRootProgress calculate_range() const {
        FlatState f = state_.to_flat();
        const uint64_t n = f.n_accounts();
        auto st = previous_state_;
        if (!st) {
            st = std::make_shared<IntermediateStateRootState>();
            e_.check(b200_root_stream_begin(e_.raw(), 1, &st->stream));
        }
        StateRootProgress out;
        const uint64_t a0 = st->next_account;
        uint64_t a1 = a0;
        // the range: accounts a0 .. a1 holding >= threshold hashed entries (at least one account)
        while (a1 < n && (a1 == a0 || (a1 - a0) + (f.seg_offsets[a1] - f.seg_offsets[a0]) < threshold_)) a1++;
        auto take = [&](b200_updates &u, bool storage) {
            std::map<uint32_t, StorageTrieUpdates> per_trie;
            for (uint64_t i = 0; i < u.n_nodes; i++) {
                if (storage) per_trie[u.trie_id[i]].storage_nodes.insert(detail::branch_node(u, i));
                else out.updates.account_nodes.insert(detail::branch_node(u, i));
            }
            b200_updates_release(&u);
            return per_trie;
        };
        if (a1 > a0) {
            const uint64_t s0 = f.seg_offsets[a0], s1 = f.seg_offsets[a1];
            std::vector<uint64_t> rel(a1 - a0 + 1);
            for (uint64_t a = a0; a <= a1; a++) rel[a - a0] = f.seg_offsets[a] - s0;
            b200_updates au{}, su{};
            e_.check(b200_root_stream_push(st->stream, f.acct_keys.data() + 32 * a0, f.accts.data() + a0, a1 - a0,
                                           f.slot_keys.data() + 32 * s0, f.slot_values.data() + 32 * s0, rel.data(), &au, &su, nullptr));
            take(au, false);
            auto per_trie = take(su, true);
            for (uint64_t a = a0; a < a1; a++) {
                B256 addr;
                std::memcpy(addr.data(), f.acct_keys.data() + 32 * a, 32);
                if (f.seg_offsets[a + 1] == f.seg_offsets[a]) out.updates.insert_storage_updates(addr, StorageTrieUpdates::deleted());
                else out.updates.insert_storage_updates(addr, per_trie[(uint32_t)(a - a0)]);
            }
            out.hashed_entries_walked = (a1 - a0) + (s1 - s0);
            st->next_account = a1;
            std::memcpy(st->last_hashed_key.data(), f.acct_keys.data() + 32 * (a1 - 1), 32);
        }
        if (a1 < n) {
            out.complete = false;
            out.state = st;
            return out;
        }
        b200_updates au{};
        e_.check(b200_root_stream_finish(st->stream, out.root.data(), &au));
        take(au, false);
        for (auto &d : prefix_sets_.destroyed_accounts) out.updates.storage_tries[d].is_deleted = true;  // updates.rs:153-157
        return out;
    }
    StateRootProgress calculate(bool retain_updates) const {
        FlatState f = state_.to_flat();
        StateRootProgress out;
        b200_updates au{}, su{};
        e_.check(b200_state_root_full(e_.raw(), f.acct_keys.data(), f.accts.data(), f.n_accounts(), f.slot_keys.data(),
                                      f.slot_values.data(), f.seg_offsets.data(), out.root.data(),
                                      retain_updates ? &au : nullptr, retain_updates ? &su : nullptr, nullptr));
        out.hashed_entries_walked = f.n_accounts() + f.slot_keys.size() / 32;
        if (retain_updates) {
            for (uint64_t i = 0; i < au.n_nodes; i++) out.updates.account_nodes.insert(detail::branch_node(au, i));
            std::map<uint32_t, StorageTrieUpdates> per_trie;
            for (uint64_t i = 0; i < su.n_nodes; i++) per_trie[su.trie_id[i]].storage_nodes.insert(detail::branch_node(su, i));
            for (uint64_t a = 0; a < f.n_accounts(); a++) {
                B256 addr;
                std::memcpy(addr.data(), f.acct_keys.data() + 32 * a, 32);
                if (f.seg_offsets[a + 1] == f.seg_offsets[a])
                    out.updates.insert_storage_updates(addr, StorageTrieUpdates::deleted());  // trie.rs:622-629
                else
                    out.updates.insert_storage_updates(addr, per_trie[(uint32_t)a]);
            }
            for (auto &d : prefix_sets_.destroyed_accounts) out.updates.storage_tries[d].is_deleted = true;  // updates.rs:153-157
            b200_updates_release(&au);
            b200_updates_release(&su);
        }
        return out;
    }
    const Engine &e_;
    HashedPostStateSorted state_;
    TriePrefixSets prefix_sets_;
    uint64_t threshold_ = 100000;  // DEFAULT_INTERMEDIATE_THRESHOLD, trie.rs:25
    std::shared_ptr<IntermediateStateRootState> previous_state_;
};

/// The account trie resident in HBM as an arena of 16-slot branch nodes (b200_dtrie_*): a block's upserts and deletes
/// are applied in place and only the touched paths are re-hashed — the part reth's sparse trie plays on the live path
/// (crates/trie/sparse/src/parallel.rs: update_leaf / remove_leaf / root).  `apply` returns the block's
/// TrieUpdates{account_nodes, removed_nodes} (crates/trie/common/src/updates.rs:17-26).
class DynamicTrie {
  public:
    DynamicTrie(const Engine &e, const FlatState &f, const std::vector<uint8_t> *storage_roots = nullptr) : e_(e) {
        e_.check(b200_dtrie_create(e_.raw(), f.acct_keys.data(), f.accts.data(), storage_roots ? storage_roots->data() : nullptr,
                                   f.n_accounts(), &t_, root_.data()));
    }
    DynamicTrie(const DynamicTrie &) = delete;
    DynamicTrie &operator=(const DynamicTrie &) = delete;
    ~DynamicTrie() { b200_dtrie_destroy(t_); }
    const B256 &root() const { return root_; }
    uint64_t leaves() const { return b200_dtrie_leaves(t_); }
    /// keys strictly ascending; present[i] == 0 deletes (nullptr: all upserts)
    std::pair<B256, TrieUpdates> apply(const std::vector<uint8_t> &keys32, const std::vector<b200_account> &accts,
                                       const std::vector<uint8_t> *present = nullptr,
                                       const std::vector<uint8_t> *storage_roots = nullptr) {
        b200_updates up{}, rm{};
        e_.check(b200_dtrie_apply(t_, keys32.data(), accts.data(), present ? present->data() : nullptr,
                                  storage_roots ? storage_roots->data() : nullptr, accts.size(), root_.data(), &up, &rm, nullptr));
        TrieUpdates out;
        for (uint64_t i = 0; i < up.n_nodes; i++) out.account_nodes.insert(detail::branch_node(up, i));
        for (uint64_t i = 0; i < rm.n_nodes; i++) out.removed_nodes.insert(detail::branch_node(rm, i).first);
        b200_updates_release(&up);
        b200_updates_release(&rm);
        return {root_, std::move(out)};
    }

  private:
    const Engine &e_;
    b200_dtrie *t_ = nullptr;
    B256 root_{};
};

/// The whole hashed state resident in HBM (b200_dstate_*): accounts and every storage trie.  `commit` applies one block's
/// HashedPostState in place and returns the new state root with the block's TrieUpdates — the role of reth's
/// SparseStateTrie on the live path (crates/trie/sparse/src/state.rs) and of StateRoot::overlay_root_with_updates
/// (crates/trie/db/src/state.rs:184-230); `account_proof` / `storage_proof` serve Proof::account_proof
/// (crates/trie/trie/src/proof/mod.rs) from the same state.
class DynamicStateRoot {
  public:
    DynamicStateRoot(const Engine &e, const HashedPostStateSorted &state) : e_(e) {
        FlatState f = state.to_flat();
        e_.check(b200_dstate_create(e_.raw(), f.acct_keys.data(), f.accts.data(), f.n_accounts(), f.slot_keys.data(),
                                    f.slot_values.data(), f.seg_offsets.data(), &s_, root_.data()));
    }
    DynamicStateRoot(const DynamicStateRoot &) = delete;
    DynamicStateRoot &operator=(const DynamicStateRoot &) = delete;
    ~DynamicStateRoot() { b200_dstate_destroy(s_); }
    const B256 &root() const { return root_; }
    uint64_t accounts() const { return b200_dstate_accounts(s_); }
    uint64_t slots() const { return b200_dstate_slots(s_); }

    std::pair<B256, TrieUpdates> commit(const HashedPostState &post) {
        std::set<B256> touched;
        for (auto &ka : post.accounts) touched.insert(ka.first);
        for (auto &ks : post.storages) touched.insert(ks.first);
        std::vector<B256> order(touched.begin(), touched.end());
        std::vector<uint8_t> keys, flags, slot_keys, slot_vals;
        std::vector<b200_account> accts(order.size());
        std::vector<uint64_t> offs{0};
        for (size_t i = 0; i < order.size(); i++) {
            const B256 &k = order[i];
            keys.insert(keys.end(), k.begin(), k.end());
            auto ia = post.accounts.find(k);
            auto is = post.storages.find(k);
            uint8_t fl = 0;
            if (ia != post.accounts.end()) {
                if (ia->second) {
                    fl = 1;  // exists
                    accts[i].nonce = ia->second->nonce;
                    std::memcpy(accts[i].balance_be, ia->second->balance.data(), 32);
                    const B256 &ch = ia->second->bytecode_hash ? *ia->second->bytecode_hash : KECCAK_EMPTY;
                    std::memcpy(accts[i].code_hash, ch.data(), 32);
                }  // else destroyed: flags 0
            } else {
                fl = 1 | 2;  // storage-only entry: account data unchanged
            }
            if (fl && is != post.storages.end()) {
                if (is->second.wiped) fl |= 4;
                for (auto &sv : is->second.storage) {
                    slot_keys.insert(slot_keys.end(), sv.first.begin(), sv.first.end());
                    slot_vals.insert(slot_vals.end(), sv.second.begin(), sv.second.end());
                }
            }
            flags.push_back(fl);
            offs.push_back(slot_keys.size() / 32);
        }
        b200_updates au{}, ar{}, su{}, sr{};
        std::vector<uint8_t> deleted(order.size() + 1, 0);
        e_.check(b200_dstate_apply(s_, keys.data(), accts.data(), flags.data(), order.size(), slot_keys.data(), slot_vals.data(),
                                   offs.data(), root_.data(), &au, &ar, &su, &sr, deleted.data(), nullptr));
        TrieUpdates out;
        for (uint64_t i = 0; i < au.n_nodes; i++) out.account_nodes.insert(detail::branch_node(au, i));
        for (uint64_t i = 0; i < ar.n_nodes; i++) out.removed_nodes.insert(detail::branch_node(ar, i).first);
        std::map<uint32_t, StorageTrieUpdates> per_entry;
        for (uint64_t i = 0; i < su.n_nodes; i++) per_entry[su.trie_id[i]].storage_nodes.insert(detail::branch_node(su, i));
        for (uint64_t i = 0; i < sr.n_nodes; i++) per_entry[sr.trie_id[i]].removed_nodes.insert(detail::branch_node(sr, i).first);
        for (size_t i = 0; i < order.size(); i++) {
            StorageTrieUpdates st = per_entry.count((uint32_t)i) ? per_entry[(uint32_t)i] : StorageTrieUpdates{};
            st.is_deleted = deleted[i] != 0;
            out.insert_storage_updates(order[i], std::move(st));
        }
        for (b200_updates *u : {&au, &ar, &su, &sr}) b200_updates_release(u);
        return {root_, std::move(out)};
    }

    /// node RLPs from the state root down to the account (or to where the trie shows it does not exist)
    std::vector<std::vector<uint8_t>> account_proof(const B256 &hashed_address) const {
        b200_proofs p{};
        e_.check(b200_dstate_account_proofs(s_, hashed_address.data(), 1, &p));
        return take(p, 0);
    }
    /// (storage root, one proof per hashed slot)
    std::pair<B256, std::vector<std::vector<std::vector<uint8_t>>>> storage_proofs(const B256 &hashed_address,
                                                                                   const std::vector<B256> &hashed_slots) const {
        std::vector<uint8_t> sk;
        for (auto &s : hashed_slots) sk.insert(sk.end(), s.begin(), s.end());
        b200_proofs p{};
        B256 sroot{};
        e_.check(b200_dstate_storage_proofs(s_, hashed_address.data(), sk.data(), hashed_slots.size(), sroot.data(), &p));
        std::vector<std::vector<std::vector<uint8_t>>> out;
        for (uint64_t t = 0; t < hashed_slots.size(); t++) out.push_back(nodes_of(p, t));
        b200_proofs_release(&p);
        return {sroot, std::move(out)};
    }

    /// MultiProof / StorageMultiProof of crates/trie/common/src/proofs.rs:180-188,594-602: node path -> RLP, with the hash / tree
    /// masks of the stored branch nodes among them (branch_node_masks, what Proof::with_branch_node_masks(true) collects).
    struct StorageMultiProof {
        B256 root{};
        std::map<Nibbles, std::vector<uint8_t>> subtree;
        std::map<Nibbles, std::pair<uint16_t, uint16_t>> branch_node_masks;  // path -> (hash_mask, tree_mask)
    };
    struct MultiProof {
        std::map<Nibbles, std::vector<uint8_t>> account_subtree;
        std::map<Nibbles, std::pair<uint16_t, uint16_t>> branch_node_masks;
        std::map<B256, StorageMultiProof> storages;
    };
    /// Proof::multiproof(MultiProofTargets) in one device call (crates/trie/trie/src/proof/mod.rs:143-193): hashed address ->
    /// hashed slot targets.
    MultiProof multiproof(const std::map<B256, std::vector<B256>> &targets) const {
        std::vector<uint8_t> ak, sk;
        std::vector<uint64_t> offs{0};
        std::vector<std::vector<B256>> slots_of;
        for (auto &kv : targets) {
            ak.insert(ak.end(), kv.first.begin(), kv.first.end());
            std::set<B256> uniq(kv.second.begin(), kv.second.end());
            slots_of.emplace_back(uniq.begin(), uniq.end());
            for (auto &s : slots_of.back()) sk.insert(sk.end(), s.begin(), s.end());
            offs.push_back(sk.size() / 32);
        }
        const uint64_t n = targets.size();
        std::vector<uint8_t> sroots(32 * (n ? n : 1));
        b200_proofs pa{}, ps{};
        e_.check(b200_dstate_multiproof(s_, ak.data(), n, offs.data(), sk.data(), &pa, sroots.data(), &ps));
        MultiProof out;
        auto fold = [](const b200_proofs &p, uint64_t t, const B256 &key, std::map<Nibbles, std::vector<uint8_t>> &subtree,
                       std::map<Nibbles, std::pair<uint16_t, uint16_t>> &masks) {
            Nibbles nib = unpack_nibbles(key);
            for (uint64_t k = p.node_offset[t]; k < p.node_offset[t + 1]; k++) {
                Nibbles path(nib.begin(), nib.begin() + p.node_depth[k]);
                subtree[path] = std::vector<uint8_t>(p.rlp + p.rlp_offset[k], p.rlp + p.rlp_offset[k + 1]);
                if (p.node_masks[k]) masks[path] = {(uint16_t)(p.node_masks[k] >> 16), (uint16_t)(p.node_masks[k] & 0xFFFF)};
            }
        };
        uint64_t i = 0;
        for (auto &kv : targets) {
            fold(pa, i, kv.first, out.account_subtree, out.branch_node_masks);
            StorageMultiProof &st = out.storages[kv.first];
            std::memcpy(st.root.data(), sroots.data() + 32 * i, 32);
            for (uint64_t j = offs[i]; j < offs[i + 1]; j++) fold(ps, j, slots_of[i][j - offs[i]], st.subtree, st.branch_node_masks);
            i++;
        }
        b200_proofs_release(&pa);
        b200_proofs_release(&ps);
        return out;
    }

  private:
    static std::vector<std::vector<uint8_t>> nodes_of(const b200_proofs &p, uint64_t t) {
        std::vector<std::vector<uint8_t>> out;
        for (uint64_t k = p.node_offset[t]; k < p.node_offset[t + 1]; k++)
            out.emplace_back(p.rlp + p.rlp_offset[k], p.rlp + p.rlp_offset[k + 1]);
        return out;
    }
    static std::vector<std::vector<uint8_t>> take(b200_proofs &p, uint64_t t) {
        auto out = nodes_of(p, t);
        b200_proofs_release(&p);
        return out;
    }
    const Engine &e_;
    b200_dstate *s_ = nullptr;
    B256 root_{};
};

/// ParallelStateRoot::{incremental_root, incremental_root_with_updates} — crates/trie/parallel/src/root.rs:35-77.
/// The storage-root fan-out and the account fold are the same device launches.
class ParallelStateRoot : public StateRoot {
  public:
    using StateRoot::StateRoot;
    B256 incremental_root() const { return root(); }
    std::pair<B256, TrieUpdates> incremental_root_with_updates() const { return root_with_updates(); }
};

// ------------------------------------------------------------------------------------------------ ordered roots
/// OrderedRootError — crates/trie/common/src/ordered_root.rs:9-80.
struct OrderedRootError : std::runtime_error {
    enum Kind { Incomplete, IndexOutOfBounds, DuplicateIndex } kind;
    size_t expected = 0, received = 0, idx = 0, len = 0;
    OrderedRootError(Kind k, std::string msg) : std::runtime_error(std::move(msg)), kind(k) {}
    static OrderedRootError incomplete(size_t expected, size_t received) {
        OrderedRootError e(Incomplete, "incomplete: expected " + std::to_string(expected) + " items, received " + std::to_string(received));
        e.expected = expected;
        e.received = received;
        return e;
    }
    static OrderedRootError out_of_bounds(size_t index, size_t len) {
        OrderedRootError e(IndexOutOfBounds, "index " + std::to_string(index) + " out of bounds for length " + std::to_string(len));
        e.idx = index;
        e.len = len;
        return e;
    }
    static OrderedRootError duplicate(size_t index) {
        OrderedRootError e(DuplicateIndex, "duplicate item at index " + std::to_string(index));
        e.idx = index;
        return e;
    }
    bool is_incomplete() const { return kind == Incomplete; }
    bool is_index_out_of_bounds() const { return kind == IndexOutOfBounds; }
    bool is_duplicate_index() const { return kind == DuplicateIndex; }
    std::optional<size_t> index() const { return kind == Incomplete ? std::nullopt : std::optional<size_t>(idx); }
};

/// Roots of many lists of pre-encoded items in one device call — alloy_trie::root::ordered_trie_root_encoded per list
/// (the transactions / receipts / withdrawals roots of a batch of blocks; b200_ordered_roots).
inline std::vector<B256> ordered_trie_roots(const Engine &e, const std::vector<std::vector<std::vector<uint8_t>>> &lists) {
    std::vector<uint64_t> seg{0}, off{0};
    std::vector<uint8_t> blob;
    for (const auto &l : lists) {
        for (const auto &it : l) {
            blob.insert(blob.end(), it.begin(), it.end());
            off.push_back(blob.size());
        }
        seg.push_back(off.size() - 1);
    }
    std::vector<B256> roots(lists.size());
    if (!lists.empty())
        e.check(b200_ordered_roots(e.raw(), blob.data(), off.data(), seg.data(), lists.size(), roots[0].data(), nullptr));
    return roots;
}
inline B256 ordered_trie_root_encoded(const Engine &e, const std::vector<std::vector<uint8_t>> &items) {
    return ordered_trie_roots(e, {items})[0];
}

/// OrderedTrieRootEncodedBuilder — ordered_root.rs:131-257: same names, argument meaning and errors.  Items are buffered
/// (the reference flushes into a HashBuilder as the key order allows); the trie is built on the device at finalize().
class OrderedTrieRootEncodedBuilder {
  public:
    OrderedTrieRootEncodedBuilder(const Engine &e, size_t len) : e_(e), len_(len), pending_(len) {}
    void push(size_t index, const std::vector<uint8_t> &bytes) {
        if (index >= len_) throw OrderedRootError::out_of_bounds(index, len_);
        if (pending_[index].has_value()) throw OrderedRootError::duplicate(index);
        push_unchecked(index, bytes);
    }
    void push_unchecked(size_t index, const std::vector<uint8_t> &bytes) {
        pending_[index] = bytes;
        received_++;
    }
    bool is_complete() const { return received_ == len_; }
    size_t pushed_count() const { return received_; }
    size_t expected_count() const { return len_; }
    B256 finalize() {
        if (len_ == 0) return EMPTY_ROOT_HASH;
        if (received_ != len_) throw OrderedRootError::incomplete(len_, received_);
        std::vector<std::vector<uint8_t>> items;
        items.reserve(len_);
        for (auto &p : pending_) items.push_back(std::move(*p));
        return ordered_trie_root_encoded(e_, items);
    }

  private:
    const Engine &e_;
    size_t len_, received_ = 0;
    std::vector<std::optional<std::vector<uint8_t>>> pending_;
};

}  // namespace reth_b200
