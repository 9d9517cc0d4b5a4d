// Internal representation of a compiled topology ("plan") and goal program.
// Host side only builds index arrays; the device mirrors are plain int32 / double arrays.
#pragma once

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "dfdm_b200.h"

namespace dfdm {

// Device-resident view of a plan (pointers into one device allocation).
struct PlanDev {
    int n = 0, E = 0, nf = 0, nx = 0, nnz = 0, w = 0;
    const int32_t* edges = nullptr;       // [E][2]
    const int32_t* node_src = nullptr;    // [n]: free index f >= 0, or -1 - k for fixed index k
    const int32_t* free_nodes = nullptr;  // [nf] node id of each free index
    // CSR of K (free order, sorted columns)
    const int32_t* rowptr = nullptr;      // [nf+1]
    const int32_t* colidx = nullptr;      // [nnz]
    const int32_t* diagslot = nullptr;    // [nf]
    // the same pattern padded to ell_w entries per row: column (own row for padding); slots are rowptr[f] + u
    int ell_w = 0;
    const int32_t* ell_col = nullptr;     // [nf][ell_w]
    // free-row incidence in free order: entry -> (edge, code) with code = CSR slot of the off-diagonal
    // (>= 0) or -1 - k when the neighbour is fixed node k (contributes to the right-hand side)
    const int32_t* finc_ptr = nullptr;    // [nf+1]
    const int32_t* finc_edge = nullptr;
    const int32_t* finc_code = nullptr;
    // the same incidence padded to inc_w entries per row (PCG row set only; inc_w = 0 when a node has too many edges):
    // edge (-1 for padding) and the free row of its other end (-1: fixed).  K p straight from q: no stored matrix.
    int inc_w = 0;
    const int32_t* inc_edge = nullptr;    // [nf][inc_w]
    const int32_t* inc_col = nullptr;     // [nf][inc_w]
    // node incidence (all nodes): C[e, node]
    const int32_t* ninc_ptr = nullptr;    // [n+1]
    const int32_t* ninc_edge = nullptr;
    const int32_t* ninc_sign = nullptr;
    // direct solver ordering: position p <-> free index
    const int32_t* perm = nullptr;        // [nf] free index at position p
    const int32_t* pos = nullptr;         // [nf] position of free index
    // incidence by solver position: code = band offset p - p_nbr (> 0, lower band), 0 = upper (skip),
    // -1 - k = fixed neighbour k
    const int32_t* sinc_ptr = nullptr;    // [nf+1]
    const int32_t* sinc_edge = nullptr;
    const int32_t* sinc_code = nullptr;
};

// One coarse level of the aggregation multigrid hierarchy (built from the level above it).
// Coarse operator = Galerkin product with piecewise-constant prolongation: K_c(I,J) = sum of the
// fine entries K(i,j), i in I, j in J - again a weighted graph Laplacian with grounding, so every
// level has the same kernels.
struct MgLevelDev {
    int n_fine = 0, n = 0, nnz = 0, ell_w = 0;
    const int32_t* agg = nullptr;        // [n_fine] fine row -> coarse row
    const int32_t* mem_ptr = nullptr;    // [n+1]    coarse row -> its fine rows
    const int32_t* mem_idx = nullptr;    // [n_fine]
    const int32_t* rowptr = nullptr;     // [n+1]    coarse CSR (sorted columns)
    const int32_t* ell_col = nullptr;    // [n][ell_w]
    const int32_t* diagslot = nullptr;   // [n]
    const int32_t* gal_ptr = nullptr;    // [nnz+1]  coarse slot -> fine slots (Galerkin sum)
    const int32_t* gal_src = nullptr;    // [nnz_fine]
    // the same operator as a SLICED ELL table for the one-CTA-per-design-pair tail kernel (pcg.cu k_mg_tail): rows in
    // slices of 32, every slice padded to its own widest row; entry (slice s, position u, lane) lives at
    // (tl_off[s] + u) * 32 + lane, so the 32 rows of a slice read consecutive words
    int tl_slices = 0;
    int64_t tl_ent = 0;                  // 32 * sum of the slice widths
    const int32_t* tl_off = nullptr;     // [tl_slices + 1] prefix sums of the slice widths
    const int32_t* tl_col = nullptr;     // [tl_ent] column (the row itself where padded)
    const int32_t* tl_src = nullptr;     // [tl_ent] CSR slot of the entry, -1 where padded
    const int32_t* tl_agg = nullptr;     // [tl_ent] aggregate (row of the next coarser level) of that column; null on the last level
};
constexpr int kMaxMgLevels = 12;
struct MgDev {
    int n_levels = 0;                    // number of COARSE levels
    MgLevelDev lv[kMaxMgLevels];
};

struct MgLevel {
    int n_fine = 0, n = 0, nnz = 0, ell_w = 0;
    std::vector<int32_t> agg, mem_ptr, mem_idx, rowptr, colidx, ell_col, diagslot, gal_ptr, gal_src;
};

// entries of a level's operator in the sliced ELL layout of the tail kernel (slices of 32 rows padded to their widest row)
inline int64_t tail_entries(const std::vector<int32_t>& rowptr, int n) {
    int64_t total = 0;
    for (int s0 = 0; s0 < n; s0 += 32) {
        int wmax = 0;
        for (int f = s0; f < n && f < s0 + 32; ++f) wmax = wmax > rowptr[f + 1] - rowptr[f] ? wmax : rowptr[f + 1] - rowptr[f];
        total += wmax;
    }
    return total * 32;
}

// The pattern of K and the assembly gather lists for one ordering of the free nodes.
struct RowSet {
    int nf = 0, nnz = 0, ell_w = 0;
    std::vector<int32_t> free_nodes;     // [nf] node id of each row
    std::vector<int32_t> node_src;       // [n]  row of a free node, or -1 - k for fixed node k
    std::vector<int32_t> rowptr, colidx, diagslot, ell_col;
    std::vector<int32_t> finc_ptr, finc_edge, finc_code;
    int inc_w = 0;
    std::vector<int32_t> inc_edge, inc_col;
};

struct Plan {
    int n = 0, E = 0, nf = 0, nx = 0, nnz = 0, w_nat = 0, w = 0;
    bool use_rcm = false;
    std::vector<int32_t> edges, free_nodes, fixed_nodes, freefixed, node_src;
    std::vector<int32_t> rowptr, colidx, diagslot, edge_slots;
    int ell_w = 0;
    std::vector<int32_t> ell_col;
    std::vector<int32_t> finc_ptr, finc_edge, finc_code;
    std::vector<int32_t> ninc_ptr, ninc_edge, ninc_sign;
    std::vector<int32_t> perm, pos;
    std::vector<int32_t> sinc_ptr, sinc_edge, sinc_code;
    RowSet pg;                            // rows in the PCG path's own (hierarchical, locality preserving) order
    std::vector<int32_t> pg_perm;         // [nf] free index of each PCG row
    std::vector<MgLevel> mg;              // coarse levels, finest first (built on the PCG row order)
    std::vector<int32_t> mg_rows, mg_nnz; // sizes of level 0 .. coarsest, for reporting (DFDM_ARR_MG_ROWS / _NNZ)
    mutable std::map<int, MgDev> mg_dev;

    // lazily created device mirrors and cached arenas for the *_host entry points
    mutable std::mutex mu;
    mutable std::map<int, PlanDev> dev;
    mutable std::map<int, PlanDev> dev_pcg;
    mutable std::map<int, void*> dev_block;
    // For the *_host entry points, kArenaSlots per device: device memory, three private streams (copy in, compute, copy out)
    // and the events that chain them.  `busy` is held for a whole host call; a second thread calling while the first slot
    // is busy takes the second, so that its copies run under the first call's kernels (api.cu: host_eval).
    static constexpr int kHostChunksMax = 8;
    static constexpr int kArenaSlots = 2;
    struct Arena {
        void* ptr = nullptr;
        int64_t bytes = 0;
        std::mutex busy;
        void* stream[3] = {nullptr, nullptr, nullptr};
        void* ev_in[kHostChunksMax] = {nullptr};
        void* ev_done[kHostChunksMax] = {nullptr};
        void* ev_end[3] = {nullptr, nullptr, nullptr};   // blocking-sync events: the end-of-call wait sleeps instead of spinning
    };
    struct ArenaSet { Arena slot[kArenaSlots]; };
    mutable std::map<int, ArenaSet> arena;
};

// One goal, flattened.  mode 0: contributes scale * weight * (p - t)^2 per component;
// mode 1 (PredictionError): contributes scale * p per component.
struct GoalRec {
    int32_t kind;
    int32_t elem;
    int32_t mode;
    int32_t pad;
    double scale;    // alpha of the term (divided by the entry count for MeanSquaredError)
    double weight;
    double p[DFDM_GOAL_NPARAMS];
};

struct GoalProgDev {
    int n_goals = 0, n_net = 0;
    const GoalRec* recs = nullptr;      // sorted: node goals by node, edge goals by edge, network goals
    const int32_t* ngoal_ptr = nullptr; // [n+1] into recs
    const int32_t* egoal_ptr = nullptr; // [E+1] into recs (absolute indices)
    int net_begin = 0;                  // network goals: recs[net_begin .. net_begin + n_net)
    double l2_alpha = 0.0;
};

struct GoalProg {
    int n = 0, E = 0;
    std::vector<GoalRec> recs;
    std::vector<int32_t> ngoal_ptr, egoal_ptr;
    int net_begin = 0, n_net = 0;
    double l2_alpha = 0.0;
    mutable std::mutex mu;
    mutable std::map<int, GoalProgDev> dev;
    mutable std::map<int, void*> dev_block;
};

void set_error(const std::string& msg);
int build_plan(int32_t n, int32_t E, const int32_t* edges, const uint8_t* is_fixed, Plan& plan);

// shared-memory footprint of the fused direct kernel for half-bandwidth w
inline int64_t direct_smem_bytes(int nf, int w) {
    int64_t ld = (int64_t)nf | 1;                      // odd leading dimension: conflict-free diagonals
    return 8 * ((int64_t)(w + 1) * ld + 3 * (int64_t)nf + 2 * (int64_t)(w + 1) + 32);
}
constexpr int64_t kDirectSmemLimit = 200 * 1024;

}  // namespace dfdm

struct dfdm_plan : dfdm::Plan {};
struct dfdm_goalprog : dfdm::GoalProg {};
