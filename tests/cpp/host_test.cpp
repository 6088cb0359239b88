// host_test.cpp — the reference's own trie tests, restated against the C++ host mirror (reth_b200/host/reth_b200.hpp)
// and checked against the CPU oracle.  Built by __graft_entry__.build(), run by tests/test_gpu_cpp_host.py.
//   account_and_storage_trie          crates/trie/db/tests/trie.rs:357-477
//   from_bundle_state_with_rayon      crates/trie/db/src/state.rs:408-437
//   storage_trie_around_extension_node crates/trie/db/tests/trie.rs:719-805
//   prefix set semantics              crates/trie/common/src/prefix_set.rs:292-305
//   ordered root builder              crates/trie/common/src/ordered_root.rs:263-353
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../oracle/oracle.h"
#include "../../reth_b200/host/reth_b200.hpp"

using namespace reth_b200;

static int failures = 0;
#define CHECK(cond)                                                       \
    do {                                                                  \
        if (!(cond)) {                                                    \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            failures++;                                                   \
        }                                                                 \
    } while (0)

static B256 b256(const char *hex) {
    B256 r{};
    for (int i = 0; i < 32; i++) {
        unsigned v;
        std::sscanf(hex + 2 * i, "%2x", &v);
        r[i] = (uint8_t)v;
    }
    return r;
}
static Address addr(const char *hex) {
    Address r{};
    for (int i = 0; i < 20; i++) {
        unsigned v;
        std::sscanf(hex + 2 * i, "%2x", &v);
        r[i] = (uint8_t)v;
    }
    return r;
}
static std::string hex(const B256 &b) {
    char s[65];
    for (int i = 0; i < 32; i++) std::snprintf(s + 2 * i, 3, "%02x", b[i]);
    return s;
}
static U256 ether(uint64_t n) {  // n * 1e18
    unsigned __int128 v = (unsigned __int128)n * 1000000000000000000ULL;
    U256 r{};
    for (int i = 0; i < 16; i++) r[31 - i] = (uint8_t)(v >> (8 * i));
    return r;
}

static void account_and_storage_trie(const Engine &e) {
    HashedPostState st;
    Address a2 = addr("7db3e81b72d2695e19764583f6d219dbee0f35ca"), a3 = addr("16b07afd1c635f77172e842a000ead9a2a222459");
    B256 key2 = KeccakKeyHasher::hash_key(e, a2.data(), 20), key3 = KeccakKeyHasher::hash_key(e, a3.data(), 20);
    CHECK(key2[0] == 0xB0 && key2[1] == 0x40 && key3[0] == 0xB0 && key3[1] == 0x41);
    st.accounts[b256("b000000000000000000000000000000000000000000000000000000000000000")] = Account{0, ether(3), std::nullopt};
    st.accounts[key2] = Account{0, ether(1), std::nullopt};
    st.accounts[key3] = Account{0, ether(2), b256("5be74cad16203c4905c068b012a2e9fb6d19d036c410f16fd177f337541440dd")};
    st.accounts[b256("B1A0000000000000000000000000000000000000000000000000000000000000")] = Account{0, ether(4), std::nullopt};
    st.accounts[b256("B310000000000000000000000000000000000000000000000000000000000000")] = Account{0, ether(8), std::nullopt};
    st.accounts[b256("B340000000000000000000000000000000000000000000000000000000000000")] = Account{0, ether(1), std::nullopt};
    HashedStorage hs;
    hs.storage[b256("1200000000000000000000000000000000000000000000000000000000000000")] = u256_from_u64(0x42);
    hs.storage[b256("1400000000000000000000000000000000000000000000000000000000000000")] = u256_from_u64(0x01);
    hs.storage[b256("3000000000000000000000000000000000000000000000000000000000E00000")] = u256_from_u64(0x127a89);
    hs.storage[b256("3000000000000000000000000000000000000000000000000000000000E00001")] = u256_from_u64(0x05);
    st.storages[key3] = hs;
    auto [root, updates] = StateRoot(e, st.into_sorted()).root_with_updates();
    CHECK(hex(root) == "72861041bc90cd2f93777956f058a545412b56de79af5eb6b8075fe2eabbe015");
    CHECK(updates.account_nodes.size() == 2);
    auto it = updates.account_nodes.begin();
    CHECK(it->first == Nibbles({0xB}));
    CHECK(it->second.state_mask == 0b1011 && it->second.tree_mask == 0b0001 && it->second.hash_mask == 0b1001);
    CHECK(it->second.hashes.size() == 2 && !it->second.root_hash);
    ++it;
    CHECK(it->first == Nibbles({0xB, 0x0}));
    CHECK(it->second.state_mask == 0b10001 && it->second.tree_mask == 0 && it->second.hash_mask == 0b10000);
    CHECK(it->second.hashes.size() == 1);
    size_t deleted = 0;
    for (auto &kv : updates.storage_tries) deleted += kv.second.is_deleted;
    CHECK(deleted == 5);  // StorageTrieUpdates::deleted() for every account without storage
    CHECK(ParallelStateRoot(e, st.into_sorted()).incremental_root() == root);
    CHECK(StateRoot(e, st.into_sorted()).root() == root);
    // add the account at keccak(0x4f61..5c91): crates/trie/db/tests/trie.rs:479-491
    Address a4b = addr("4f61f2d5ebd991b85aa1677db97307caf5215c91");
    st.accounts[KeccakKeyHasher::hash_key(e, a4b.data(), 20)] = Account{0, ether(5), std::nullopt};
    CHECK(hex(StateRoot(e, st.into_sorted()).root()) == "8e263cd4eefb0c3cbbb14e5541a66a755cad25bcfab1e10dd9d706263e811b28");
}

static void from_bundle_state(const Engine &e) {
    Address a1{}, a2{};
    a1[19] = 1;
    a2[19] = 2;
    BundleAccount b1, b2;
    b1.info = Account{1, {}, std::nullopt};
    b1.storage.push_back({u256_from_u64(1015), u256_from_u64(10)});
    b2.info = Account{2, {}, std::nullopt};
    b2.storage.push_back({u256_from_u64(2015), u256_from_u64(20)});
    auto post = HashedPostState::from_bundle_state(e, {{a1, b1}, {a2, b2}});
    CHECK(post.accounts.size() == 2 && post.storages.size() == 2);
    CHECK(hex(StateRoot(e, post.into_sorted()).root()) == "b464525710cafcf5d4044ac85b72c08b1e76231b8d91f288fe438cc41d8eaafd");
}

static void extension_node_storage_trie(const Engine &e) {
    const char *keys[6] = {"30af561000000000000000000000000000000000000000000000000000000000",
                           "30af569000000000000000000000000000000000000000000000000000000000",
                           "30af650000000000000000000000000000000000000000000000000000000000",
                           "30af6f0000000000000000000000000000000000000000000000000000000000",
                           "30af8f0000000000000000000000000000000000000000000000000000000000",
                           "3100000000000000000000000000000000000000000000000000000000000000"};
    HashedStorage hs;
    for (auto k : keys) hs.storage[b256(k)] = u256_from_u64(1);
    auto [root, walked, upd] = StorageRoot(e, B256{}, hs.into_sorted()).root_with_updates();
    CHECK(walked == 6 && upd.storage_nodes.size() == 2 && !upd.is_deleted);
    auto &n3 = upd.storage_nodes.at(Nibbles({3}));
    CHECK(n3.state_mask == 0b0011 && n3.tree_mask == 0b0001 && n3.hash_mask == 0 && n3.hashes.empty());
    auto &n30af = upd.storage_nodes.at(Nibbles({3, 0, 0xA, 0xF}));
    CHECK(n30af.state_mask == 0b101100000 && n30af.tree_mask == 0 && n30af.hash_mask == 0b001000000 && n30af.hashes.size() == 1);
    // oracle agrees on the root
    std::vector<uint8_t> k, v;
    for (auto &sv : hs.storage) {
        k.insert(k.end(), sv.first.begin(), sv.first.end());
        v.insert(v.end(), sv.second.begin(), sv.second.end());
    }
    uint64_t offs[2] = {0, 6};
    B256 oroot;
    CHECK(orc_storage_roots(k.data(), v.data(), offs, 1, oroot.data(), nullptr, 1) == 0 && oroot == root);
    // empty storage and all-zero storage
    auto [r0, w0, u0] = StorageRoot(e, B256{}, HashedStorageSorted{}).root_with_updates();
    CHECK(r0 == EMPTY_ROOT_HASH && w0 == 0 && u0.is_deleted);
    HashedStorage z;
    z.storage[b256(keys[0])] = U256{};
    CHECK(StorageRoot(e, B256{}, z.into_sorted()).root() == EMPTY_ROOT_HASH);
}

static void prefix_sets_and_destroyed(const Engine &e) {
    PrefixSetMut m;
    m.insert({1, 2, 3});
    m.insert({1, 2, 4});
    m.insert({4, 5, 6});
    m.insert({1, 2, 3});
    PrefixSet ps = m.freeze();
    CHECK(ps.contains({1, 2}) && ps.contains({4, 5}) && !ps.contains({7, 8}) && ps.len() == 3);
    B256 k1, k2;
    k1.fill(0x11);
    k2.fill(0x22);
    HashedPostState st;
    st.accounts[k1] = Account{1, u256_from_u64(1), std::nullopt};
    st.accounts[k2] = std::nullopt;  // destroyed
    st.storages[k2].wiped = true;
    auto sets = st.construct_prefix_sets().freeze();
    CHECK(sets.destroyed_accounts.count(k2) == 1 && sets.storage_prefix_sets.at(k2).is_all());
    auto [root, upd] = StateRoot(e, st.into_sorted()).with_prefix_sets(sets).root_with_updates();
    HashedPostState only;
    only.accounts[k1] = st.accounts[k1];
    CHECK(root == StateRoot(e, only.into_sorted()).root());
    CHECK(upd.storage_tries.at(k2).is_deleted);
    CHECK(StateRoot(e, HashedPostStateSorted{}).root() == EMPTY_ROOT_HASH);
}

static void random_state_vs_oracle(const Engine &e) {
    std::mt19937_64 rng(42);
    HashedPostState st;
    for (int i = 0; i < 3000; i++) {
        B256 k;
        for (auto &b : k) b = (uint8_t)rng();
        Account a{rng() & 0xffff, u256_from_u64(rng()), std::nullopt};
        if (i % 3 == 0) {
            B256 ch;
            for (auto &b : ch) b = (uint8_t)rng();
            a.bytecode_hash = ch;
        }
        st.accounts[k] = a;
        if (i % 5 == 0)
            for (int s = 0; s < 1 + (int)(rng() % 40); s++) {
                B256 sk;
                for (auto &b : sk) b = (uint8_t)rng();
                st.storages[k].storage[sk] = u256_from_u64(rng() | 1);
            }
    }
    auto sorted = st.into_sorted();
    FlatState f = sorted.to_flat();
    B256 oroot;
    orc_updates oa{}, os{};
    CHECK(orc_state_root_full(f.acct_keys.data(), reinterpret_cast<const orc_account *>(f.accts.data()), f.n_accounts(),
                              f.slot_keys.data(), f.slot_values.data(), f.seg_offsets.data(), oroot.data(), &oa, &os, 4) == 0);
    auto [root, upd] = StateRoot(e, sorted).root_with_updates();
    CHECK(root == oroot);
    CHECK(upd.account_nodes.size() == oa.n_nodes);
    size_t storage_nodes = 0;
    for (auto &kv : upd.storage_tries) storage_nodes += kv.second.storage_nodes.size();
    CHECK(storage_nodes == os.n_nodes);
    orc_updates_free(&oa);
    orc_updates_free(&os);
    // thresholded build (StateRoot::with_threshold / with_intermediate_state / root_with_progress, trie.rs:73-88,156):
    // looped until Complete == the one-shot root and updates (reth: arbitrary_state_root_with_progress)
    for (uint64_t threshold : {uint64_t(1), uint64_t(700), uint64_t(1) << 40}) {
        std::shared_ptr<IntermediateStateRootState> inter;
        TrieUpdates acc;
        size_t steps = 0, walked = 0;
        for (;;) {
            auto p = StateRoot(e, sorted).with_threshold(threshold).with_intermediate_state(inter).root_with_progress();
            steps++;
            walked += p.hashed_entries_walked;
            for (auto &kv : p.updates.account_nodes) acc.account_nodes[kv.first] = kv.second;
            for (auto &kv : p.updates.storage_tries) acc.storage_tries[kv.first] = kv.second;
            if (p.complete) {
                CHECK(!p.state && p.root == oroot);
                break;
            }
            CHECK(p.state != nullptr);
            CHECK(p.state->checkpoint().resume_nibble <= 16);
            inter = p.state;
        }
        CHECK(walked == f.n_accounts() + f.slot_keys.size() / 32);
        CHECK(acc.account_nodes == upd.account_nodes);
        CHECK(acc.storage_tries.size() == upd.storage_tries.size());
        for (auto &kv : upd.storage_tries) CHECK(acc.storage_tries.at(kv.first).storage_nodes == kv.second.storage_nodes);
        CHECK((steps == 1) == (threshold >= walked));
    }
    // rows laid out on the device == rows the host encoder makes of the same build's TrieUpdates
    {
        for (b200_key_format fmt : {B200_KEYS_LEGACY, B200_KEYS_PACKED}) {
            auto t = StateRoot(e, sorted).root_with_table_rows(fmt);
            CHECK(t.root == oroot);
            CHECK(t.accounts_trie == account_trie_rows(upd, fmt));
            CHECK(t.storages_trie == storage_trie_rows(upd, fmt));
            CHECK(!t.accounts_trie.empty() && !t.storages_trie.empty());
        }
    }
}

// DynamicTrie: blocks of inserts / deletes / updates applied in place == oracle root over the merged state, and the
// node set {previous - removed + updated} == the oracle's stored nodes (the incremental == full criterion of
// crates/trie/db/tests/trie.rs:680-717).
static void dynamic_trie_blocks(const Engine &e) {
    std::mt19937_64 rng(7);
    std::map<B256, b200_account> state;
    auto rand_key = [&] {
        B256 k;
        for (auto &b : k) b = (uint8_t)rng();
        return k;
    };
    auto rand_acct = [&] {
        b200_account a{};
        a.nonce = rng() & 0xffff;
        for (int i = 24; i < 32; i++) a.balance_be[i] = (uint8_t)rng();
        std::memcpy(a.code_hash, KECCAK_EMPTY.data(), 32);
        return a;
    };
    for (int i = 0; i < 1500; i++) state[rand_key()] = rand_acct();
    auto flat_of = [&](const std::map<B256, b200_account> &st) {
        FlatState f;
        for (auto &kv : st) {
            f.acct_keys.insert(f.acct_keys.end(), kv.first.begin(), kv.first.end());
            f.accts.push_back(kv.second);
        }
        f.seg_offsets.assign(st.size() + 1, 0);
        return f;
    };
    auto oracle_nodes = [&](const std::map<B256, b200_account> &st, B256 &root) {
        FlatState f = flat_of(st);
        orc_updates ou{};
        CHECK(orc_state_root(f.acct_keys.data(), reinterpret_cast<const orc_account *>(f.accts.data()), nullptr, st.size(),
                             root.data(), &ou) == 0);
        std::set<std::pair<std::vector<uint8_t>, uint8_t>> paths;
        for (uint64_t i = 0; i < ou.n_nodes; i++)
            paths.insert({std::vector<uint8_t>(ou.path_packed + 32 * i, ou.path_packed + 32 * i + 32), ou.path_len[i]});
        orc_updates_free(&ou);
        return paths;
    };
    B256 oroot;
    auto db = oracle_nodes(state, oroot);
    DynamicTrie trie(e, flat_of(state));
    CHECK(trie.root() == oroot);
    auto pack = [](const Nibbles &p) {
        std::vector<uint8_t> out(32, 0);
        for (size_t j = 0; j < p.size(); j++) out[j >> 1] |= (j & 1) ? p[j] : (uint8_t)(p[j] << 4);
        return std::make_pair(out, (uint8_t)p.size());
    };
    for (int block = 0; block < 6; block++) {
        std::map<B256, std::pair<uint8_t, b200_account>> dirty;
        std::vector<B256> existing;
        for (auto &kv : state) existing.push_back(kv.first);
        for (int i = 0; i < 120; i++) {
            int r = (int)(rng() % 3);
            if (r == 0) dirty[rand_key()] = {1, rand_acct()};
            else if (r == 1) dirty[existing[rng() % existing.size()]] = {0, b200_account{}};
            else dirty[existing[rng() % existing.size()]] = {1, rand_acct()};
        }
        std::vector<uint8_t> keys, present;
        std::vector<b200_account> accts;
        for (auto &kv : dirty) {
            keys.insert(keys.end(), kv.first.begin(), kv.first.end());
            present.push_back(kv.second.first);
            accts.push_back(kv.second.second);
            if (kv.second.first) state[kv.first] = kv.second.second;
            else state.erase(kv.first);
        }
        auto [root, upd] = trie.apply(keys, accts, &present);
        auto expect = oracle_nodes(state, oroot);
        CHECK(root == oroot);
        CHECK(trie.leaves() == state.size());
        for (auto &p : upd.removed_nodes) db.erase(pack(p));
        for (auto &kv : upd.account_nodes) db.insert(pack(kv.first));
        CHECK(db == expect);
    }
}

// DynamicStateRoot: blocks with new / changed / destroyed accounts and slot writes / zeroing / wipes; root == oracle over
// the merged state after every block, and an account proof hashes up to that root.  Opt-in on a GPU like dynamic_trie_blocks.
static void dynamic_state_blocks(const Engine &e) {
    std::mt19937_64 rng(11);
    auto rand_key = [&] {
        B256 k;
        for (auto &b : k) b = (uint8_t)rng();
        return k;
    };
    HashedPostState merged;
    for (int i = 0; i < 400; i++) {
        B256 k = rand_key();
        merged.accounts[k] = Account{rng() & 0xff, u256_from_u64(rng()), std::nullopt};
        if (i % 3 == 0)
            for (int s = 0; s < 1 + (int)(rng() % 20); s++) merged.storages[k].storage[rand_key()] = u256_from_u64(rng() | 1);
    }
    auto oracle_root = [&](const HashedPostState &st) {
        FlatState f = st.into_sorted().to_flat();
        B256 r;
        CHECK(orc_state_root_full(f.acct_keys.data(), reinterpret_cast<const orc_account *>(f.accts.data()), f.n_accounts(),
                                  f.slot_keys.data(), f.slot_values.data(), f.seg_offsets.data(), r.data(), nullptr, nullptr, 2) == 0);
        return r;
    };
    DynamicStateRoot ds(e, merged.into_sorted());
    CHECK(ds.root() == oracle_root(merged));
    for (int block = 0; block < 4; block++) {
        HashedPostState post;
        std::vector<B256> live;
        for (auto &ka : merged.accounts) live.push_back(ka.first);
        for (int i = 0; i < 20; i++) {  // balance changes
            const B256 &k = live[rng() % live.size()];
            Account a = *merged.accounts[k];
            a.nonce++;
            post.accounts[k] = a;
        }
        for (int i = 0; i < 5; i++) {  // new accounts with storage
            B256 k = rand_key();
            post.accounts[k] = Account{0, u256_from_u64(7), std::nullopt};
            post.storages[k].storage[rand_key()] = u256_from_u64(9);
        }
        for (int i = 0; i < 3; i++) {  // destroyed
            const B256 &k = live[rng() % live.size()];
            post.accounts[k] = std::nullopt;
            post.storages[k].wiped = true;
            post.storages[k].storage.clear();
        }
        int touched = 0;
        for (auto &ks : merged.storages) {  // storage-only writes, one zeroed slot each
            if (post.accounts.count(ks.first) || ks.second.storage.empty() || ++touched > 6) continue;
            post.storages[ks.first].storage[ks.second.storage.begin()->first] = U256{};
            post.storages[ks.first].storage[rand_key()] = u256_from_u64(rng() | 1);
        }
        auto [root, upd] = ds.commit(post);
        for (auto &ks : post.storages) {  // the model: HashedPostState::extend + deletion rules
            HashedStorage &cur = merged.storages[ks.first];
            if (ks.second.wiped) cur.storage.clear();
            for (auto &sv : ks.second.storage) {
                if (is_zero(sv.second)) cur.storage.erase(sv.first);
                else cur.storage[sv.first] = sv.second;
            }
        }
        for (auto &ka : post.accounts) {
            if (!ka.second) {
                merged.accounts.erase(ka.first);
                merged.storages.erase(ka.first);
            } else {
                merged.accounts[ka.first] = ka.second;
            }
        }
        CHECK(root == oracle_root(merged));
        CHECK(ds.accounts() == merged.accounts.size());
        CHECK(!upd.account_nodes.empty());
        // an inclusion proof: its first node hashes to the root
        auto proof = ds.account_proof(merged.accounts.begin()->first);
        CHECK(!proof.empty());
        B256 h;
        orc_keccak256(proof[0].data(), proof[0].size(), h.data());
        CHECK(h == root);
        // the multiproof of a few accounts with slot targets == the union of their single proofs; storage roots included
        std::map<B256, std::vector<B256>> targets;
        int picked = 0;
        for (auto &ks : merged.storages) {
            if (!merged.accounts.count(ks.first) || ks.second.storage.empty() || ++picked > 4) continue;
            targets[ks.first] = {ks.second.storage.begin()->first, rand_key()};
        }
        targets[rand_key()] = {rand_key()};  // an account that does not exist
        auto mp = ds.multiproof(targets);
        CHECK(mp.storages.size() == targets.size());
        CHECK(mp.account_subtree.count(Nibbles{}) == 1);
        size_t union_nodes = 0;
        std::set<std::vector<uint8_t>> seen;
        for (auto &kv : targets) {
            for (auto &node : ds.account_proof(kv.first))
                if (seen.insert(node).second) union_nodes++;
            auto sp = ds.storage_proofs(kv.first, kv.second);
            CHECK(sp.first == mp.storages[kv.first].root);
            std::set<std::vector<uint8_t>> snodes;
            for (auto &pr : sp.second) snodes.insert(pr.begin(), pr.end());
            std::set<std::vector<uint8_t>> got;
            for (auto &pn : mp.storages[kv.first].subtree) got.insert(pn.second);
            CHECK(got == snodes);
        }
        CHECK(mp.account_subtree.size() == union_nodes);
        for (auto &bm : mp.branch_node_masks) CHECK(mp.account_subtree.count(bm.first) == 1 && (bm.second.first | bm.second.second));
    }
}

// Table rows — host-only, runs before a device is needed.  Key vectors: crates/trie/common/src/nibbles.rs
// :321-346 (StoredNibbles [2,4] -> 02 04; subkey = 64 nibble bytes + count), :443-450 (packed 0xAB 0xC0 .. 03).
static void table_rows_host_only() {
    TrieUpdates u;
    BranchNodeCompact n;
    n.state_mask = 0xf607, n.tree_mask = 0x0005, n.hash_mask = 0x4004;
    n.hashes = {b256("90d53cd810cc5d4243766cd4451e7b9d14b736a1148b26b3baac7617f617d321"),
                b256("cc35c964dda53ba6c0b87798073a9628dbc9cd26b5cce88eb69655a9c609caf1")};
    u.account_nodes[Nibbles{0xA, 0xB, 0xC}] = n;
    u.account_nodes[Nibbles{2, 4}] = n;
    auto rows = account_trie_rows(u, B200_KEYS_LEGACY);
    CHECK(rows.size() == 2);
    CHECK((rows[0].key == std::vector<uint8_t>{2, 4}));
    CHECK((rows[1].key == std::vector<uint8_t>{0xA, 0xB, 0xC}));
    CHECK(rows[0].value.size() == 6 + 64);
    const uint8_t masks[6] = {0xf6, 0x07, 0x00, 0x05, 0x40, 0x04};
    CHECK(std::memcmp(rows[0].value.data(), masks, 6) == 0);
    CHECK(std::memcmp(rows[0].value.data() + 6, n.hashes[0].data(), 32) == 0);
    CHECK(std::memcmp(rows[0].value.data() + 38, n.hashes[1].data(), 32) == 0);
    auto packed = account_trie_rows(u, B200_KEYS_PACKED);
    CHECK(packed.size() == 2 && packed[1].key.size() == 33);
    CHECK(packed[1].key[0] == 0xAB && packed[1].key[1] == 0xC0 && packed[1].key[2] == 0 && packed[1].key[32] == 3);
    CHECK(packed[0].key[0] == 0x24 && packed[0].key[32] == 2);

    B256 a1 = b256("1100000000000000000000000000000000000000000000000000000000000000");
    B256 a0 = b256("0100000000000000000000000000000000000000000000000000000000000000");
    StorageTrieUpdates s1, s0;
    s1.storage_nodes[Nibbles{2, 4}] = n;
    s0.storage_nodes[Nibbles{5}] = n;
    s0.storage_nodes[Nibbles{5, 0}] = n;
    u.insert_storage_updates(a1, s1);
    u.insert_storage_updates(a0, s0);
    u.insert_storage_updates(b256("2200000000000000000000000000000000000000000000000000000000000000"), StorageTrieUpdates::deleted());
    auto srows = storage_trie_rows(u, B200_KEYS_LEGACY);
    CHECK(srows.size() == 3);
    CHECK(std::memcmp(srows[0].key.data(), a0.data(), 32) == 0 && std::memcmp(srows[2].key.data(), a1.data(), 32) == 0);
    CHECK(srows[0].value.size() == 65 + 70 && srows[0].value[0] == 5 && srows[0].value[64] == 1);
    CHECK(srows[1].value[0] == 5 && srows[1].value[1] == 0 && srows[1].value[64] == 2);
    CHECK(srows[2].value[0] == 2 && srows[2].value[1] == 4 && srows[2].value[2] == 0 && srows[2].value[64] == 2);
    CHECK(std::memcmp(srows[2].value.data() + 65, masks, 6) == 0);
    auto sp = storage_trie_rows(u, B200_KEYS_PACKED);
    CHECK(sp.size() == 3 && sp[2].value.size() == 33 + 70 && sp[2].value[0] == 0x24 && sp[2].value[32] == 2);
}

// crates/trie/common/src/ordered_root.rs:263-353 (equivalence, out of order, empty, incomplete, index errors)
static void ordered_root_builder(const Engine &e) {
    auto item = [](size_t i) {
        std::string s = "item_" + std::to_string(i) + "_data";
        return std::vector<uint8_t>(s.begin(), s.end());
    };
    auto oracle_root = [](const std::vector<std::vector<uint8_t>> &items) {
        std::vector<uint8_t> blob(1);
        std::vector<uint64_t> off{0};
        blob.clear();
        for (auto &it : items) {
            blob.insert(blob.end(), it.begin(), it.end());
            off.push_back(blob.size());
        }
        uint64_t seg[2] = {0, items.size()};
        B256 r{};
        blob.push_back(0);
        CHECK(orc_ordered_roots(blob.data(), off.data(), seg, 1, r.data()) == 0);
        return r;
    };
    for (size_t len : {0, 1, 2, 3, 10, 127, 128, 129, 130, 200}) {
        std::vector<std::vector<uint8_t>> items;
        for (size_t i = 0; i < len; i++) items.push_back(item(i));
        B256 expected = oracle_root(items);
        OrderedTrieRootEncodedBuilder b(e, len);
        for (size_t i = 0; i < len; i++) b.push(i, items[i]);
        CHECK(b.finalize() == expected);
        OrderedTrieRootEncodedBuilder rev(e, len);
        for (size_t i = len; i-- > 0;) rev.push(i, items[i]);
        CHECK(rev.finalize() == expected);
    }
    {
        OrderedTrieRootEncodedBuilder b(e, 0);
        CHECK(b.is_complete() && b.finalize() == EMPTY_ROOT_HASH);
    }
    {
        OrderedTrieRootEncodedBuilder b(e, 3);
        b.push(0, item(0));
        b.push(1, item(1));
        CHECK(!b.is_complete());
        bool threw = false;
        try {
            b.finalize();
        } catch (const OrderedRootError &err) {
            threw = err.is_incomplete() && err.expected == 3 && err.received == 2 && !err.index().has_value();
        }
        CHECK(threw);
    }
    {
        OrderedTrieRootEncodedBuilder b(e, 2);
        bool oob = false, dup = false;
        try {
            b.push(5, item(5));
        } catch (const OrderedRootError &err) {
            oob = err.is_index_out_of_bounds() && err.idx == 5 && err.len == 2;
        }
        b.push(0, item(0));
        try {
            b.push(0, item(9));
        } catch (const OrderedRootError &err) {
            dup = err.is_duplicate_index() && err.index() == std::optional<size_t>(0);
        }
        b.push(1, item(1));
        CHECK(oob && dup && b.is_complete() && b.pushed_count() == 2 && b.expected_count() == 2);
    }
    // a batch with empty lists
    std::vector<std::vector<std::vector<uint8_t>>> lists(5);
    for (size_t i = 0; i < 140; i++) lists[1].push_back(std::vector<uint8_t>(1 + i * 3, (uint8_t)i));
    lists[3].push_back(item(7));
    auto roots = ordered_trie_roots(e, lists);
    for (size_t l = 0; l < lists.size(); l++) CHECK(roots[l] == oracle_root(lists[l]));
}

int main() {
    table_rows_host_only();
    if (failures) {
        std::printf("host_test: %d FAILURES (host-only part)\n", failures);
        return 1;
    }
    try {
        Engine e(0);
        account_and_storage_trie(e);
        from_bundle_state(e);
        extension_node_storage_trie(e);
        prefix_sets_and_destroyed(e);
        random_state_vs_oracle(e);
        {
            dynamic_trie_blocks(e);
            dynamic_state_blocks(e);
            ordered_root_builder(e);
        }
    } catch (const B200Error &err) {
        std::printf("B200Error: %s\n", err.what());
        return err.status == B200_ERR_NO_DEVICE ? 77 : 2;
    }
    std::printf(failures ? "host_test: %d FAILURES\n" : "host_test: all checks passed\n", failures);
    return failures ? 1 : 0;
}
