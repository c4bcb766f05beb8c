// Host-side forest specialiser + packer.
//
// Takes one tree ensemble in the SoA form of fmc_forest_desc and produces the node table and the
// root stream the sm_100a kernels walk.  Two things happen here, both exact (they never change
// which leaf a row reaches, and they keep the reference's tree order, so margins are bit-identical
// to walking the original trees -- the oracle walks the ORIGINAL trees, which is what makes the
// parity tests meaningful):
//
//  1. Constant folding.  Columns whose value is the same for every row of a launch are resolved
//     at pack time: all one-hot columns (the OneHotEncoder half of ColumnTransformer.transform,
//     FMC:744/756/784-809, with the hot columns given by fmc_set_active_columns) and, in "sim"
//     mode, the six numerics that are constant per (offense, defense) orientation: both timeouts
//     (never spent, FMC:911-912) and the four SP+ ratings (FMC:1001-1004).  For CSR-fed boosters an
//     exact zero is a MISSING value and takes the node's default branch.
//     A tree that folds to a single leaf becomes a CONSTANT: it is not walked, its value is added
//     at its place in the tree order.
//
//  2. Re-layout for a branch-free, lock-step walk.  Surviving trees are taken kIlp at a time, in
//     tree order ("group").  All trees of a group are walked for exactly D levels, D = the deepest
//     tree of the group, with no per-lane leaf test:
//        node slot (8 bytes): lo32 = float threshold
//                             hi32 = (feature byte offset) << 20 | byte offset of the LEFT child
//                                    inside the table's 1 MiB window (8-byte aligned; the right
//                                    child is the next slot)
//        xgboost leaf        : lo32 = the float32 leaf value, hi32 = a node that tests the "-inf"
//                              feature column and points at ITSELF -> a lane that reaches a leaf
//                              early stays there (-inf < value is always true => left => self)
//        sklearn leaf        : the float64 value itself (pre-multiplied by the learning rate);
//                              a leaf above level D is reached through a chain of pass-through
//                              nodes (again testing the "-inf" column), so that after D levels
//                              every lane holds leaf bits.
//     "feature byte offset": the kernels keep the feature rows of 32 requests FEATURE-MAJOR in
//     shared memory ([feature][lane], 128 bytes per feature), so a lane's load of any feature hits
//     its own bank -- conflict-free whatever node each lane is at.
//     The walk starts from the ROOT STREAM: the root slots of the groups, kIlp x 8 bytes per group,
//     in tree order, read with warp-uniform 16-byte loads.  The low three bits of the four child
//     fields of a group carry its metadata (D, "has constants", window of the group inside a table
//     larger than 1 MiB).  Constants live in a side stream that is only touched by groups that have
//     them.
//
// "Feature rows" are the columns of the per-request feature record.  For CSR-fed boosters a
// non-flag numeric that may be exactly zero gets two rows: row A holds -inf when the value is 0
// (used by default-left nodes: -inf < thr is always true) and row B holds +inf (default-right
// nodes).  0/1 flags need no second row: their node is rewritten as `flag < 0.5` with the children
// placed so that "absent" follows the default branch.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fmc.h"

namespace fmc {

constexpr int kNumMax = 17;
#ifndef FMC_ILP
#define FMC_ILP 4
#endif
constexpr int kIlp = FMC_ILP;              // trees walked together per lane (one group)
static_assert(kIlp == 4, "the group metadata layout (3 spare bits x 4 roots) assumes 4 trees per group");
constexpr int kFeatShift = 20;             // hi32 >> 20 = byte offset of the feature inside a 32-request chunk
constexpr int kFeatBytes = 128;            // one feature of a chunk = 32 lanes x 4 bytes
constexpr uint32_t kChildMask = 0x000FFFF8u;   // byte offset of the left child inside the window
constexpr uint32_t kMetaMask = 0x7u;
constexpr size_t kWindowBytes = 1u << 20;  // a table must sit inside one 1 MiB-aligned window
constexpr int kMaxGroupDepth = 15;

struct PackSpec {
    int32_t active[2] = {-1, -1};   // hot one-hot columns
    uint32_t fold_mask = 0;         // bit k: numeric k is constant
    double fold_value[kNumMax] = {0};
    int8_t row[kNumMax];            // feature row of numeric k (A view)
    int8_t row_b[kNumMax];          // B view row, or -1
    uint8_t is_flag[kNumMax];       // 0/1 valued numeric
    int ninf_row = 0;               // feature row that always holds -inf
    bool nan_views = false;         // predict preset: every numeric of an xgboost forest has a B view, so that a NaN (= missing
                                    // for xgboost, whatever the booster was fed) follows the node's default branch
    int tree_begin = 0, tree_end = -1;
    // player mode: one-hot columns whose 0/1 value is a per-request feature row (the sampled passer /
    // target / rusher of the play) instead of a pack-time constant
    int n_dyn = 0;
    int32_t dyn_col[FMC_MAX_PASSER_ROWS + FMC_MAX_NAME_ROWS];
    int8_t dyn_row[FMC_MAX_PASSER_ROWS + FMC_MAX_NAME_ROWS];
    int dyn_row_of(int col) const {
        for (int i = 0; i < n_dyn; ++i)
            if (dyn_col[i] == col) return dyn_row[i];
        return -1;
    }
    // play_model: fold values are standardised first
    int n_scaled = 0;
    int32_t scaler_cols[16];
    double scaler_mean[16], scaler_scale[16];
};

struct HostForest {
    bool loaded = false;
    int kind = 0, n_outputs = 0, n_features = 0, num_base = 0, n_num = 0, zero_is_missing = 0;
    double base[8] = {0};
    double scale = 1.0;
    std::vector<int32_t> feat, left, right, root, out;
    std::vector<float> thr;
    std::vector<uint8_t> dl;
    std::vector<double> value;
    int n_scaled = 0;
    int32_t scaler_cols[16];
    double scaler_mean[16], scaler_scale[16];
    int32_t active[2] = {-1, -1};

    void assign(const fmc_forest_desc &d) {
        kind = d.kind; n_outputs = d.n_outputs; n_features = d.n_features; num_base = d.num_base;
        n_num = d.n_num; zero_is_missing = d.zero_is_missing; scale = d.scale;
        for (int k = 0; k < 8; ++k) base[k] = d.base[k];
        feat.assign(d.feat, d.feat + d.n_nodes);
        left.assign(d.left, d.left + d.n_nodes);
        right.assign(d.right, d.right + d.n_nodes);
        thr.assign(d.thr, d.thr + d.n_nodes);
        dl.assign(d.default_left, d.default_left + d.n_nodes);
        value.assign(d.value, d.value + d.n_nodes);
        root.assign(d.tree_root, d.tree_root + d.n_trees);
        out.assign(d.tree_out, d.tree_out + d.n_trees);
        loaded = true;
    }
};

struct PackedForest {
    // split thresholds of the WALKED nodes per feature row (exact-memo keys, fmc_memo.hpp): what a request's
    // feature values are compared against, hence all that its outputs depend on
    std::vector<std::vector<float>> row_thr;
    std::vector<uint64_t> slots;        // node table; child offsets are relative to the table start until relocate()
    std::vector<uint8_t> slot_is_node;  // 1: hi32 carries a child offset (relocatable), 0: raw float64 leaf
    std::vector<uint64_t> stream;       // root slots, kIlp per group, outputs concatenated
    std::vector<uint8_t> stream_is_node;
    std::vector<uint64_t> consts;       // side stream: per group with constants, 1 count word + the values
    uint32_t stream_off[8] = {0};       // first stream word (8 bytes) of each output
    uint32_t n_groups[8] = {0};
    uint32_t consts_off[8] = {0};       // first side-stream word of each output
    int n_outputs = 0;
    int rounds = 0;                     // real boosting rounds per output in range
    int max_depth = 0;
    uint64_t internal = 0, leaves = 0, constants = 0, pass_through = 0;

    size_t table_bytes() const { return slots.size() * 8; }
    // Add the table's byte offset inside its window to every child field (tables of at most one window;
    // larger tables are placed window-aligned and need no relocation).
    void relocate(uint32_t base) {
        if (base == 0) return;
        for (size_t i = 0; i < slots.size(); ++i)
            if (slot_is_node[i]) slots[i] += (uint64_t)base << 32;
        for (size_t i = 0; i < stream.size(); ++i)
            if (stream_is_node[i]) stream[i] += (uint64_t)base << 32;
    }
};

// Feature-row presets -------------------------------------------------------------------------
// NUM order (FMC:676-682): 0 down 1 distance 2 yardsToGoal 3 is_red_zone 4 score_diff
// 5 seconds_remaining 6 offenseTimeouts 7 defenseTimeouts 8..11 SP+ 12 goal_to_go
// 13 fourth_and_short 14 fg_range 15 half 16 two_minute
constexpr int kSimNinfRow = 14;   // 11 varying numerics + B views of distance, yardsToGoal, score_diff, then -inf
constexpr int kSimRows = 15;
constexpr int kPredNinfRow = 29;  // 17 numerics + the B views of the 12 that are not flags, then -inf (the slot format
constexpr int kPredRows = 30;     // addresses at most 32 feature rows)

inline void preset_sim(PackSpec &s) {
    const int8_t row[kNumMax] = {0, 1, 2, 3, 4, 5, -1, -1, -1, -1, -1, -1, 6, 7, 8, 9, 10};
    const int8_t rb[kNumMax] = {-1, 11, 12, -1, 13, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
    const uint8_t fl[kNumMax] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 1};
    std::memcpy(s.row, row, sizeof(row)); std::memcpy(s.row_b, rb, sizeof(rb)); std::memcpy(s.is_flag, fl, sizeof(fl));
    s.fold_mask = 0xFC0;  // numerics 6..11
    s.ninf_row = kSimNinfRow;
}
inline void preset_predict(PackSpec &s) {
    const uint8_t fl[kNumMax] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 1};
    int nb = kNumMax;
    for (int k = 0; k < kNumMax; ++k) {
        s.row[k] = (int8_t)k;
        s.is_flag[k] = fl[k];
        s.row_b[k] = fl[k] ? -1 : (int8_t)nb++;
    }
    s.fold_mask = 0;
    s.ninf_row = kPredNinfRow;
    s.nan_views = true;
}

namespace detail {
struct PNode {
    bool leaf;
    double value;
    int row;
    float thr;
    int l, r;
};

struct Builder {
    const HostForest &f;
    const PackSpec &s;
    std::vector<PNode> nodes;
    float cst[kNumMax];
    Builder(const HostForest &f_, const PackSpec &s_) : f(f_), s(s_) {
        double v[kNumMax];
        for (int k = 0; k < kNumMax; ++k) v[k] = s.fold_value[k];
        for (int i = 0; i < s.n_scaled; ++i) {
            int k = s.scaler_cols[i];
            v[k] = (v[k] - s.scaler_mean[i]) / s.scaler_scale[i];
        }
        for (int k = 0; k < kNumMax; ++k) cst[k] = (float)v[k];
    }
    int leaf(double v) { nodes.push_back({true, v, 0, 0.f, -1, -1}); return (int)nodes.size() - 1; }
    static bool same_leaf(const PNode &a, const PNode &b) {
        return a.leaf && b.leaf && std::memcmp(&a.value, &b.value, sizeof(double)) == 0;
    }
    // direction a constant takes at original node i
    bool const_left(int i, float v) const {
        if (f.kind == FMC_KIND_XGB) {
            if (f.zero_is_missing && v == 0.0f) return f.dl[i] != 0;
            return v < f.thr[i];
        }
        return v <= f.thr[i];
    }
    int build(int i) {
        int dyn = -1;      // feature row of a dynamic one-hot column
        for (;;) {
            if (f.left[i] < 0) return leaf(f.value[i]);
            int col = f.feat[i];
            if (col >= f.num_base && col < f.num_base + f.n_num) {
                int k = col - f.num_base;
                if (!((s.fold_mask >> k) & 1u)) break;
                i = const_left(i, cst[k]) ? f.left[i] : f.right[i];
            } else {
                if (s.n_dyn && (dyn = s.dyn_row_of(col)) >= 0) break;
                float v = (col == s.active[0] || col == s.active[1]) ? 1.0f : 0.0f;
                i = const_left(i, v) ? f.left[i] : f.right[i];
            }
        }
        bool zm = f.kind == FMC_KIND_XGB && f.zero_is_missing;
        if (dyn >= 0) {
            // 0/1 column read per request: row holds 1.0 when the sampled name lights this column
            const bool l0 = const_left(i, 0.0f), l1 = const_left(i, 1.0f);
            if (l0 == l1) return build(l0 ? f.left[i] : f.right[i]);
            int absent = build(l0 ? f.left[i] : f.right[i]);
            int present = build(l1 ? f.left[i] : f.right[i]);
            if (same_leaf(nodes[absent], nodes[present])) return absent;
            // `row < 0.5` (xgboost) / `row <= 0.5` (sklearn): 0 goes left, 1 goes right
            nodes.push_back({false, 0.0, dyn, 0.5f, absent, present});
            return (int)nodes.size() - 1;
        }
        int k = f.feat[i] - f.num_base;
        if (zm && s.is_flag[k]) {
            bool present_left = 1.0f < f.thr[i];
            bool missing_left = f.dl[i] != 0;
            if (present_left == missing_left) return build(present_left ? f.left[i] : f.right[i]);
            int m = build(missing_left ? f.left[i] : f.right[i]);
            int p = build(present_left ? f.left[i] : f.right[i]);
            if (same_leaf(nodes[m], nodes[p])) return m;
            nodes.push_back({false, 0.0, s.row[k], 0.5f, m, p});
            return (int)nodes.size() - 1;
        }
        int l = build(f.left[i]);
        int r = build(f.right[i]);
        if (same_leaf(nodes[l], nodes[r])) return l;
        int row = s.row[k];
        // default-right nodes read the B view: it holds +inf where the value is missing (an exact zero of a CSR-fed
        // booster; a NaN of any xgboost forest in the predict preset), the A view holds -inf there
        if ((zm || (s.nan_views && f.kind == FMC_KIND_XGB)) && !f.dl[i] && s.row_b[k] >= 0) row = s.row_b[k];
        nodes.push_back({false, 0.0, row, f.thr[i], l, r});
        return (int)nodes.size() - 1;
    }
    int depth_of(int i) const {
        const PNode &n = nodes[i];
        if (n.leaf) return 0;
        int a = depth_of(n.l), b = depth_of(n.r);
        return 1 + (a > b ? a : b);
    }
};

inline uint64_t f64_bits(double v) { uint64_t u; std::memcpy(&u, &v, 8); return u; }
inline uint32_t f32_bits(float v) { uint32_t u; std::memcpy(&u, &v, 4); return u; }
}  // namespace detail

// Value word of a constant / leaf as the kernels add it: sklearn the float64, xgboost float32 in lo32.
inline uint64_t value_word(int kind, double v) {
    if (kind == FMC_KIND_SKL) return detail::f64_bits(v);
    return (uint64_t)detail::f32_bits((float)v);
}

// Returns "" on success, otherwise an error message.
inline std::string pack_forest(const HostForest &f, const PackSpec &s, PackedForest &out) {
    using namespace detail;
    if (!f.loaded) return "model not loaded";
    const int n_trees = (int)f.root.size();
    const int tb = s.tree_begin, te = (s.tree_end < 0 || s.tree_end > n_trees) ? n_trees : s.tree_end;
    if (tb < 0 || tb > te) return "bad tree range";
    if (f.n_outputs > 8) return "too many outputs";
    out = PackedForest();
    out.n_outputs = f.n_outputs;
    const bool skl = f.kind == FMC_KIND_SKL;
    if (s.ninf_row < 0 || s.ninf_row * kFeatBytes >= (1 << (32 - kFeatShift))) return "-inf feature row out of range";
    const uint32_t ninf_hi = (uint32_t)(s.ninf_row * kFeatBytes) << kFeatShift;

    // slots of the group being laid out; child fields hold GROUP-relative slot indices until placement
    std::vector<uint64_t> gs;
    std::vector<uint8_t> gs_node;
    auto push_slot = [&](uint64_t w, bool node) -> uint32_t {
        gs.push_back(w);
        gs_node.push_back(node ? 1 : 0);
        return (uint32_t)(gs.size() - 1);
    };
    auto node_word = [&](uint32_t lo, uint32_t feat_hi, uint32_t child_slot) -> uint64_t {
        return ((uint64_t)(feat_hi | (child_slot * 8u)) << 32) | lo;
    };

    std::vector<std::vector<int>> per_out(f.n_outputs);
    for (int t = tb; t < te; ++t) {
        if (f.out[t] < 0 || f.out[t] >= f.n_outputs) return "tree output index out of range";
        per_out[f.out[t]].push_back(t);
    }
    size_t rounds = 0;
    for (auto &v : per_out) rounds = v.size() > rounds ? v.size() : rounds;
    out.rounds = (int)rounds;

    Builder b(f, s);
    b.nodes.reserve(f.left.size() / (size_t)(f.n_outputs > 0 ? f.n_outputs : 1) + 64);
    out.slots.reserve(f.left.size() + 1024);
    out.slot_is_node.reserve(f.left.size() + 1024);
    out.stream.reserve((size_t)(te - tb) + 64);
    out.stream_is_node.reserve((size_t)(te - tb) + 64);
    // a tree of the output being laid out: its nodes live in the builder's pool (one allocation per output, not per tree)
    struct Entry { int root = -1; int depth = 0; std::vector<uint64_t> pre; };   // root < 0: padding tree
    std::vector<int> order, depth_q;
    std::vector<uint32_t> slot_of;

    for (int k = 0; k < f.n_outputs; ++k) {
        out.stream_off[k] = (uint32_t)out.stream.size();
        out.consts_off[k] = (uint32_t)out.consts.size();
        // ---- trees of this output in order: constants attach to the next walked tree
        std::vector<Entry> entries;
        std::vector<uint64_t> pending;
        auto flush_padding_for_pending = [&]() {   // a run of more than 255 constants needs a padding tree to carry it
            Entry e;
            e.pre.swap(pending);
            entries.push_back(std::move(e));
        };
        b.nodes.clear();
        for (size_t j = 0; j < per_out[k].size(); ++j) {
            const size_t mark = b.nodes.size();
            const int root = b.build(f.root[per_out[k][j]]);
            if (b.nodes[root].leaf) {
                if (!(std::fabs(b.nodes[root].value) < 1e60)) return "leaf value out of range";
                if (pending.size() == 255) flush_padding_for_pending();
                pending.push_back(value_word(f.kind, b.nodes[root].value));
                out.constants++;
                b.nodes.resize(mark);       // a folded tree leaves nothing in the pool
                continue;
            }
            Entry e;
            e.root = root;
            e.depth = b.depth_of(root);
            if (e.depth > kMaxGroupDepth) return "tree deeper than the walk supports (15 levels)";
            e.pre.swap(pending);
            entries.push_back(std::move(e));
        }
        if (!pending.empty()) flush_padding_for_pending();
        while (entries.size() % kIlp) entries.push_back(Entry());
        out.n_groups[k] = (uint32_t)(entries.size() / kIlp);

        // ---- groups
        for (size_t g = 0; g < entries.size(); g += kIlp) {
            int D = 0;
            size_t n_pre = 0;
            for (int q = 0; q < kIlp; ++q) {
                if (entries[g + q].depth > D) D = entries[g + q].depth;
                n_pre += entries[g + q].pre.size();
            }
            if (skl && D == 0) D = 1;    // a float64 root cannot carry the group metadata: an all-padding group walks one pass-through level
            if (D > out.max_depth) out.max_depth = D;
            gs.clear();
            gs_node.clear();
            uint64_t rootw[kIlp];
            for (int q = 0; q < kIlp; ++q) {
                const Entry &e = entries[g + q];
                if (e.root < 0) {
                    // padding tree: D levels down to a +0.0 leaf.  xgboost: a self-pointing 0.0f leaf;
                    // sklearn: the float64 0.0 behind a chain of D pass-through nodes (D >= 1 here)
                    if (skl) {
                        uint32_t next = push_slot(f64_bits(0.0), false);
                        for (int x = 1; x < D; ++x) next = push_slot(node_word(0, ninf_hi, next), true);
                        rootw[q] = node_word(0, ninf_hi, next);
                    } else {
                        const uint32_t me = push_slot(0, true);
                        gs[me] = node_word(f32_bits(0.0f), ninf_hi, me);
                        rootw[q] = gs[me];
                    }
                    continue;
                }
                // breadth-first layout, children adjacent; the root lives only in the stream
                order.assign(1, e.root);
                slot_of.assign(1, 0xFFFFFFFFu);
                depth_q.assign(1, 0);
                for (size_t qi = 0; qi < order.size(); ++qi) {
                    const PNode &n = b.nodes[order[qi]];
                    uint64_t w;
                    bool is_node = true;
                    if (n.leaf) {
                        if (!(std::fabs(n.value) < 1e60)) return "leaf value out of range";
                        out.leaves++;
                        if (skl) {
                            // pass-through chain so that the float64 is reached after exactly D levels
                            const int extra = D - depth_q[qi];
                            if (extra == 0) { w = f64_bits(n.value); is_node = false; }
                            else {
                                uint32_t next = push_slot(f64_bits(n.value), false);
                                for (int x = 1; x < extra; ++x) next = push_slot(node_word(0, ninf_hi, next), true);
                                w = node_word(0, ninf_hi, next);
                                out.pass_through += (uint64_t)extra;
                            }
                        } else {
                            w = node_word(f32_bits((float)n.value), ninf_hi, slot_of[qi]);   // self-pointing leaf
                        }
                    } else {
                        if (n.row < 0 || n.row * kFeatBytes >= (1 << (32 - kFeatShift))) return "feature row out of range for the table format";
                        const uint32_t c = push_slot(0, true);
                        push_slot(0, true);
                        w = node_word(f32_bits(n.thr), (uint32_t)(n.row * kFeatBytes) << kFeatShift, c);
                        if ((size_t)n.row >= out.row_thr.size()) out.row_thr.resize((size_t)n.row + 1);
                        out.row_thr[(size_t)n.row].push_back(n.thr);
                        out.internal++;
                        order.push_back(n.l); slot_of.push_back(c); depth_q.push_back(depth_q[qi] + 1);
                        order.push_back(n.r); slot_of.push_back(c + 1); depth_q.push_back(depth_q[qi] + 1);
                    }
                    if (qi == 0) rootw[q] = w;      // a walked tree's root is always an internal node
                    else { gs[slot_of[qi]] = w; gs_node[slot_of[qi]] = is_node ? 1 : 0; }
                }
            }
            // ---- placement: a group never straddles a 1 MiB boundary of the table (tables above 1 MiB are
            // placed window-aligned by the arena, smaller ones are relocated as a whole)
            const size_t gbytes = gs.size() * 8;
            if (gbytes > kWindowBytes) return "one tree group exceeds the 1 MiB window of the node format";
            size_t pos = out.slots.size() * 8;
            if (pos % kWindowBytes + gbytes > kWindowBytes) {
                const size_t pad = (kWindowBytes - pos % kWindowBytes) / 8;
                out.slots.insert(out.slots.end(), pad, 0);
                out.slot_is_node.insert(out.slot_is_node.end(), pad, 0);
                pos = out.slots.size() * 8;
            }
            const uint32_t window = (uint32_t)(pos / kWindowBytes);
            if (window > 63) return "packed table exceeds 64 windows of 1 MiB";
            const uint64_t shift = (uint64_t)(pos % kWindowBytes) << 32;      // group-relative -> window-relative offsets
            for (size_t i = 0; i < gs.size(); ++i) {
                out.slots.push_back(gs_node[i] ? gs[i] + shift : gs[i]);
                out.slot_is_node.push_back(gs_node[i]);
            }
            // ---- group metadata in the three spare low bits of the four child fields:
            // depth (4 bits), "has constants" (1 bit), window of the group inside the table (6 bits)
            const uint32_t meta[kIlp] = {(uint32_t)D & 7u, (((uint32_t)D >> 3) & 1u) | (n_pre ? 2u : 0u), window & 7u, (window >> 3) & 7u};
            for (int q = 0; q < kIlp; ++q) {
                out.stream.push_back((rootw[q] + shift) | ((uint64_t)meta[q] << 32));
                out.stream_is_node.push_back(1);
            }
            if (n_pre) {
                uint64_t cw = 0;
                for (int q = 0; q < kIlp; ++q) cw |= (uint64_t)entries[g + q].pre.size() << (8 * q);
                out.consts.push_back(cw);
                for (int q = 0; q < kIlp; ++q)
                    for (uint64_t v : entries[g + q].pre) out.consts.push_back(v);
            }
        }
    }
    return "";
}

}  // namespace fmc
