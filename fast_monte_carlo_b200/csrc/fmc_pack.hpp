// Host-side forest specialiser + packer.
//
// Takes one tree ensemble in the SoA form of fmc_forest_desc and produces the 8-byte-slot node
// table the sm_100a kernels walk.  Two things happen here, both exact (they never change which
// leaf a row reaches, so margins are bit-identical to walking the original trees in the same
// tree order -- the oracle walks the ORIGINAL trees, which is what makes the parity tests
// meaningful):
//
//  1. Constant folding.  Columns whose value is the same for every row of a launch are resolved
//     at pack time: all one-hot columns (the OneHotEncoder half of ColumnTransformer.transform,
//     FMC:744/756/784-809, with the hot columns given by fmc_set_active_columns) and, in "sim"
//     mode, the six numerics that are constant per (offense, defense) orientation: both timeouts
//     (never spent, FMC:911-912) and the four SP+ ratings (FMC:1001-1004).  For CSR-fed boosters an
//     exact zero is a MISSING value and takes the node's default branch.
//  2. Re-layout.  Each surviving tree is written breadth-first into 8-byte slots with the two
//     children of a node adjacent (left at c, right at c+1):
//        internal slot : lo32 = float threshold, hi32 = 0x50000000 | (4 * feature row) << 20 | child slot
//        leaf slot     : the value itself (sklearn: the float64; xgboost: float32 in lo32, hi32 = 0)
//     A slot is internal iff (int32)hi32 >= 0x50000000: as the high word of a float64 that range
//     means a positive value >= 2^257, which no leaf holds (checked at pack time).  20 child bits
//     (1 M slots per table) and 8 bits of byte offset into the feature row (64 rows).
//
// "Feature rows" are the columns of the per-lane feature record the kernels build in shared
// memory.  For CSR-fed boosters a non-flag numeric that may be exactly zero gets two rows: row A
// holds -inf when the value is 0 (used by default-left nodes: -inf < thr is always true) and
// row B holds +inf (default-right nodes).  0/1 flags need no second row: their node is rewritten
// as `flag < 0.5` with the children placed so that "absent" follows the default branch.
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
constexpr int kIlp = FMC_ILP;             // trees walked together per lane; rounds are padded to a multiple
constexpr int kRootWords = kIlp * 2;      // one group = the kIlp root SLOTS themselves (8 bytes each), copied inline
constexpr int kChildBits = 20;
constexpr uint32_t kInternalTag = 0x50000000u;

struct PackSpec {
    int32_t active[2] = {-1, -1};   // hot one-hot columns
    uint32_t fold_mask = 0;         // bit k: numeric k is constant
    double fold_value[kNumMax] = {0};
    int8_t row[kNumMax];            // feature row of numeric k (A view)
    int8_t row_b[kNumMax];          // B view row, or -1
    uint8_t is_flag[kNumMax];       // 0/1 valued numeric
    int tree_begin = 0, tree_end = -1;
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
    std::vector<uint32_t> roots;   // [n_outputs][rounds_padded][2]: a COPY of every tree's root slot (lo, hi), in tree order
    std::vector<uint64_t> slots;
    int n_outputs = 0;
    int rounds = 0;                // real boosting rounds per output in range
    int rounds_padded = 0;
    int max_depth = 0;
    uint64_t internal = 0, leaves = 0;
};

// Feature-row presets -------------------------------------------------------------------------
// NUM order (FMC:676-682): 0 down 1 distance 2 yardsToGoal 3 is_red_zone 4 score_diff
// 5 seconds_remaining 6 offenseTimeouts 7 defenseTimeouts 8..11 SP+ 12 goal_to_go
// 13 fourth_and_short 14 fg_range 15 half 16 two_minute
constexpr int kSimRows = 14;      // 11 varying numerics + B views of distance, yardsToGoal, score_diff
constexpr int kSimStride = 15;    // odd => conflict-free shared-memory rows
constexpr int kPredRows = 29;     // 17 numerics + 12 B views
constexpr int kPredStride = 29;

inline void preset_sim(PackSpec &s) {
    const int8_t row[kNumMax] = {0, 1, 2, 3, 4, 5, -1, -1, -1, -1, -1, -1, 6, 7, 8, 9, 10};
    const int8_t rb[kNumMax] = {-1, 11, 12, -1, 13, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
    const uint8_t fl[kNumMax] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 1};
    std::memcpy(s.row, row, sizeof(row)); std::memcpy(s.row_b, rb, sizeof(rb)); std::memcpy(s.is_flag, fl, sizeof(fl));
    s.fold_mask = 0xFC0;  // numerics 6..11
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
        for (;;) {
            if (f.left[i] < 0) return leaf(f.value[i]);
            int col = f.feat[i];
            if (col >= f.num_base && col < f.num_base + f.n_num) {
                int k = col - f.num_base;
                if (!((s.fold_mask >> k) & 1u)) break;
                i = const_left(i, cst[k]) ? f.left[i] : f.right[i];
            } else {
                float v = (col == s.active[0] || col == s.active[1]) ? 1.0f : 0.0f;
                i = const_left(i, v) ? f.left[i] : f.right[i];
            }
        }
        int k = f.feat[i] - f.num_base;
        bool zm = f.kind == FMC_KIND_XGB && f.zero_is_missing;
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
        if (zm && !f.dl[i] && s.row_b[k] >= 0) row = s.row_b[k];
        nodes.push_back({false, 0.0, row, f.thr[i], l, r});
        return (int)nodes.size() - 1;
    }
};
}  // namespace detail

inline uint64_t leaf_slot(int kind, double v) {
    uint64_t u;
    if (kind == FMC_KIND_SKL) {
        std::memcpy(&u, &v, 8);
    } else {
        float fv = (float)v;
        uint32_t lo;
        std::memcpy(&lo, &fv, 4);
        u = lo;
    }
    return u;
}

// Returns "" on success, otherwise an error message.
inline std::string pack_forest(const HostForest &f, const PackSpec &s, PackedForest &out) {
    if (!f.loaded) return "model not loaded";
    int n_trees = (int)f.root.size();
    int tb = s.tree_begin, te = (s.tree_end < 0 || s.tree_end > n_trees) ? n_trees : s.tree_end;
    if (tb < 0 || tb > te) return "bad tree range";
    out = PackedForest();
    out.n_outputs = f.n_outputs;
    std::vector<std::vector<int>> per_out(f.n_outputs);
    for (int t = tb; t < te; ++t) per_out[f.out[t]].push_back(t);
    size_t rounds = 0;
    for (auto &v : per_out) rounds = v.size() > rounds ? v.size() : rounds;
    out.rounds = (int)rounds;
    out.rounds_padded = (int)((rounds + kIlp - 1) / kIlp * kIlp);
    if (out.rounds_padded == 0) out.rounds_padded = kIlp;
    out.roots.assign((size_t)f.n_outputs * out.rounds_padded * 2, 0);   // padding trees: the +0.0 leaf

    const uint32_t child_cap = 1u << kChildBits;
    const int feat_cap = 64;

    out.slots.clear();
    out.slots.push_back(leaf_slot(f.kind, 0.0));  // slot 0: the zero leaf used by padding trees
    for (auto &r : out.roots) r = 0;

    detail::Builder b(f, s);
    std::vector<int> order, depth;  // BFS queue of PNode ids, slot of each queued node
    std::vector<uint32_t> slot_of;
    for (int k = 0; k < f.n_outputs; ++k) {
        for (size_t j = 0; j < per_out[k].size(); ++j) {
            b.nodes.clear();
            int root = b.build(f.root[per_out[k][j]]);
            // breadth-first layout
            uint32_t root_slot = (uint32_t)out.slots.size();
            out.slots.push_back(0);
            order.assign(1, root);
            slot_of.assign(1, root_slot);
            depth.assign(1, 0);
            for (size_t q = 0; q < order.size(); ++q) {
                const detail::PNode &n = b.nodes[order[q]];
                uint32_t me = slot_of[q];
                if (n.leaf) {
                    if (!(std::fabs(n.value) < 1e60)) return "leaf value out of range for the slot format";
                    out.slots[me] = leaf_slot(f.kind, n.value);
                    out.leaves++;
                    if (depth[q] > out.max_depth) out.max_depth = depth[q];
                    continue;
                }
                uint32_t c = (uint32_t)out.slots.size();
                if (c + 1 >= child_cap) return "packed table exceeds the child-index range";
                if (n.row < 0 || n.row >= feat_cap) return "feature row out of range for the table format";
                out.slots.push_back(0);
                out.slots.push_back(0);
                uint32_t lo, hi;
                std::memcpy(&lo, &n.thr, 4);
                hi = kInternalTag | ((uint32_t)(n.row * 4) << kChildBits) | c;
                out.slots[me] = ((uint64_t)hi << 32) | lo;
                out.internal++;
                order.push_back(n.l); slot_of.push_back(c); depth.push_back(depth[q] + 1);
                order.push_back(n.r); slot_of.push_back(c + 1); depth.push_back(depth[q] + 1);
            }
            out.roots[((size_t)k * out.rounds_padded + j) * 2] = (uint32_t)out.slots[root_slot];
            out.roots[((size_t)k * out.rounds_padded + j) * 2 + 1] = (uint32_t)(out.slots[root_slot] >> 32);
        }
    }
    return "";
}

}  // namespace fmc
