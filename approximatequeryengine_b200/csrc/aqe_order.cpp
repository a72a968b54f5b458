// aqe_order.cpp -- the row order of the reference's B+ tree for tables that hold DUPLICATE ids (host only, no CUDA).
//
// Every scan, sampler and file of the reference sees the table in leaf-chain order (collect_all_records
// custom_bplus_db.cpp:660, collect_leaf_records :715).  For distinct ids that is ascending id, which is all the
// device columns need.  Equal ids are another matter: a leaf insert goes in FRONT of the equal keys already in that
// leaf (std::lower_bound, :32-37), the descent sends an id equal to a separator to the RIGHT child (:222-225), a
// full leaf keeps its first 127 keys and hands the other 128 to a new right sibling (:43-58), and insert_batch first
// runs std::sort -- which is not stable -- over the batch (:196-200).  So where duplicates end up depends on the
// history of inserts and on where the leaves split.  This file replays that history on (id, arrival number) pairs
// with the same tree mechanics and returns the resulting order; the engine calls it only for tables in which
// duplicate ids were seen (aqe_engine.cu order_like_reference), every other table is sorted by id.
//
// The replay keeps to the reference's mechanics literally where they decide the order -- including the separator an
// internal split hands to its parent (the first key of the NEW node, :222-231 + :59-68, i.e. old key mid + 1; old key
// mid is dropped) -- because with equal keys around a split the position of a later duplicate follows from them.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "aqe_order.hpp"

namespace aqe {
namespace {

constexpr int kMaxKeys = 255;   // BPlusTreeNode::MAX_KEYS (custom_bplus_db.hpp)

struct Tree {
    struct Node {
        bool leaf = true;
        std::vector<int64_t> keys;
        std::vector<uint32_t> seq;     // leaf: arrival numbers, parallel to keys
        std::vector<uint32_t> child;   // internal: node indices, keys.size() + 1 of them
        int64_t next = -1;             // leaf chain
    };
    std::vector<Node> nodes;
    uint32_t root = 0;
    uint32_t last_leaf = 0;            // rightmost leaf (restore() appends there)

    Tree() { nodes.emplace_back(); nodes[0].keys.reserve(kMaxKeys); nodes[0].seq.reserve(kMaxKeys); }

    uint32_t split(uint32_t ni) {
        const uint32_t nn = (uint32_t)nodes.size();
        nodes.emplace_back();
        Node& a = nodes[ni];
        Node& b = nodes[nn];
        const int mid = kMaxKeys / 2;
        b.leaf = a.leaf;
        if (a.leaf) {
            b.keys.reserve(kMaxKeys); b.seq.reserve(kMaxKeys);
            b.keys.assign(a.keys.begin() + mid, a.keys.end());
            b.seq.assign(a.seq.begin() + mid, a.seq.end());
            a.keys.resize(mid); a.seq.resize(mid);
            b.next = a.next; a.next = nn;
            if (last_leaf == ni) last_leaf = nn;
        } else {
            b.keys.assign(a.keys.begin() + mid + 1, a.keys.end());
            b.child.assign(a.child.begin() + mid + 1, a.child.end());
            a.keys.resize(mid); a.child.resize(mid + 1);
        }
        return nn;
    }
    // true: the node is full and its parent must split it
    bool insert_into(uint32_t ni, int64_t id, uint32_t s) {
        if (nodes[ni].leaf) {
            Node& n = nodes[ni];
            const size_t pos = (size_t)(std::lower_bound(n.keys.begin(), n.keys.end(), id) - n.keys.begin());
            n.keys.insert(n.keys.begin() + pos, id);
            n.seq.insert(n.seq.begin() + pos, s);
            return (int)n.keys.size() >= kMaxKeys;
        }
        size_t i = 0;
        while (i < nodes[ni].keys.size() && id >= nodes[ni].keys[i]) ++i;
        const uint32_t c = nodes[ni].child[i];
        if (!insert_into(c, id, s)) return false;
        const uint32_t nc = split(c);              // (may reallocate `nodes`: no references held across it)
        const int64_t sep = nodes[nc].keys[0];
        Node& n = nodes[ni];
        n.keys.insert(n.keys.begin() + i, sep);
        n.child.insert(n.child.begin() + i + 1, nc);
        return (int)n.keys.size() >= kMaxKeys;
    }
    void grow_root() {
        const uint32_t nn = split(root);
        const int64_t sep = nodes[nn].keys[0];
        const uint32_t nr = (uint32_t)nodes.size();
        nodes.emplace_back();
        Node& r = nodes[nr];
        r.leaf = false;
        r.keys.push_back(sep);
        r.child.push_back(root); r.child.push_back(nn);
        root = nr;
    }
    void insert(int64_t id, uint32_t s) { if (insert_into(root, id, s)) grow_root(); }

    // A table that already IS in tree order (pulled back from the device before an append): its rows go to the end of
    // the rightmost leaf one by one -- the shape sorted inserts build -- without the re-ordering of equal ids.
    bool append_into(uint32_t ni, int64_t id, uint32_t s) {
        if (nodes[ni].leaf) {
            nodes[ni].keys.push_back(id); nodes[ni].seq.push_back(s);
            return (int)nodes[ni].keys.size() >= kMaxKeys;
        }
        const size_t i = nodes[ni].keys.size();
        const uint32_t c = nodes[ni].child[i];
        if (!append_into(c, id, s)) return false;
        const uint32_t nc = split(c);
        const int64_t sep = nodes[nc].keys[0];
        nodes[ni].keys.push_back(sep);
        nodes[ni].child.push_back(nc);
        return (int)nodes[ni].keys.size() >= kMaxKeys;
    }
    void append(int64_t id, uint32_t s) { if (append_into(root, id, s)) grow_root(); }
};

}  // namespace

bool reference_order(const int64_t* ids, uint64_t n, const OrderOp* ops, size_t n_ops, uint64_t* perm) {
    if (n > kReferenceOrderMaxRows) return false;
    uint64_t covered = 0;
    for (size_t o = 0; o < n_ops; ++o) covered += ops[o].rows;
    if (covered != n) return false;
    Tree t;
    struct Item { int64_t id; uint32_t s; };
    std::vector<Item> batch;
    uint64_t at = 0;
    for (size_t o = 0; o < n_ops; ++o) {
        const uint64_t cnt = ops[o].rows;
        if (ops[o].kind == ORDER_OP_RESTORE) {
            for (uint64_t k = 0; k < cnt; ++k) t.append(ids[at + k], (uint32_t)(at + k));
        } else {
            batch.resize(cnt);
            for (uint64_t k = 0; k < cnt; ++k) batch[k] = Item{ids[at + k], (uint32_t)(at + k)};
            // insert_batch: std::sort with `a.id < b.id` (:198-200); unstable, and which equal element lands where depends only on the
            // comparisons, not on the element type -- the same library algorithm over (id, arrival) permutes like the reference's
            if (cnt > 1) std::sort(batch.begin(), batch.end(), [](const Item& a, const Item& b) { return a.id < b.id; });
            for (uint64_t k = 0; k < cnt; ++k) t.insert(batch[k].id, batch[k].s);
        }
        at += cnt;
    }
    uint32_t leaf = t.root;
    while (!t.nodes[leaf].leaf) leaf = t.nodes[leaf].child[0];
    uint64_t out = 0;
    for (int64_t cur = leaf; cur >= 0; cur = t.nodes[(size_t)cur].next)
        for (uint32_t s : t.nodes[(size_t)cur].seq) perm[out++] = s;
    return out == n;
}

}  // namespace aqe
