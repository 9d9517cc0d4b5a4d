// Topology compiler: integer preprocessing done once per network.
//
// Replaces what the reference recomputes on every evaluation (src/dfdm/equilibrium/model.py:44-47:
// column slices of the dense branch-node matrix) and what it builds per model
// (src/dfdm/equilibrium/structure.py:63-106).  Everything here must be bit-exact with respect to the
// reference's indexing: free / fixed lists ascending, freefixed = position of each node in
// [free; fixed], C[e,u] = -1, C[e,v] = +1.
#include "plan.hpp"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <queue>

namespace dfdm {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

// Reverse Cuthill-McKee over the free-node graph given as CSR (diagonal entries ignored).
static void rcm_order(int nf, const std::vector<int32_t>& rowptr, const std::vector<int32_t>& colidx,
                      std::vector<int32_t>& perm) {
    perm.clear();
    perm.reserve(nf);
    std::vector<int32_t> degree(nf), level(nf, -1);
    std::vector<char> visited(nf, 0);
    for (int i = 0; i < nf; ++i) degree[i] = rowptr[i + 1] - rowptr[i] - 1;

    auto bfs_far = [&](int start, std::vector<int32_t>& order) {
        // plain BFS; returns the visiting order, fills level[] for the component
        order.clear();
        order.push_back(start);
        level[start] = 0;
        for (size_t head = 0; head < order.size(); ++head) {
            int i = order[head];
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                int j = colidx[k];
                if (j != i && level[j] < 0) {
                    level[j] = level[i] + 1;
                    order.push_back(j);
                }
            }
        }
    };

    std::vector<int32_t> comp, nbrs;
    for (int seed = 0; seed < nf; ++seed) {
        if (visited[seed]) continue;
        // pseudo-peripheral start: iterate BFS from the farthest, lowest-degree node
        int start = seed;
        bfs_far(start, comp);
        for (int it = 0; it < 8; ++it) {
            int depth = level[comp.back()];
            int best = comp.back();
            for (int idx = (int)comp.size() - 1; idx >= 0 && level[comp[idx]] == depth; --idx)
                if (degree[comp[idx]] < degree[best] || (degree[comp[idx]] == degree[best] && comp[idx] < best)) best = comp[idx];
            for (int i : comp) level[i] = -1;
            std::vector<int32_t> trial;
            bfs_far(best, trial);
            int new_depth = level[trial.back()];
            if (new_depth > depth) {
                start = best;
                comp.swap(trial);
            } else {
                for (int i : trial) level[i] = -1;
                bfs_far(start, comp);
                break;
            }
        }
        for (int i : comp) level[i] = -1;
        // Cuthill-McKee from `start`: neighbours by ascending degree, ties by index
        size_t begin = perm.size();
        perm.push_back(start);
        visited[start] = 1;
        for (size_t head = begin; head < perm.size(); ++head) {
            int i = perm[head];
            nbrs.clear();
            for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                int j = colidx[k];
                if (j != i && !visited[j]) {
                    visited[j] = 1;
                    nbrs.push_back(j);
                }
            }
            std::sort(nbrs.begin(), nbrs.end(), [&](int a, int b) { return degree[a] != degree[b] ? degree[a] < degree[b] : a < b; });
            perm.insert(perm.end(), nbrs.begin(), nbrs.end());
        }
        std::reverse(perm.begin() + begin, perm.end());
    }
}

static int bandwidth(int nf, const std::vector<int32_t>& rowptr, const std::vector<int32_t>& colidx,
                     const std::vector<int32_t>& pos) {
    int w = 0;
    for (int i = 0; i < nf; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) w = std::max(w, std::abs(pos[i] - pos[colidx[k]]));
    return w;
}

// ---- aggregation multigrid hierarchy (topology only) ----------------------------------------------
// Pairwise matching applied twice per level (aggregates of up to 4 rows): visit rows by ascending
// degree, match each unmatched row with its unmatched neighbour of largest connection weight
// (weight = number of fine couplings the connection stands for).  Everything is integer and
// deterministic; the numerical coarse operators are summed per design on the device.
namespace {

struct Graph {
    int n = 0;
    std::vector<int32_t> ptr, idx, wgt;   // adjacency without self loops
};

Graph graph_from_csr(int n, const std::vector<int32_t>& rowptr, const std::vector<int32_t>& colidx) {
    Graph g;
    g.n = n;
    g.ptr.assign(n + 1, 0);
    for (int i = 0; i < n; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
            if (colidx[k] != i) {
                g.idx.push_back(colidx[k]);
                g.wgt.push_back(1);
                g.ptr[i + 1]++;
            }
    for (int i = 0; i < n; ++i) g.ptr[i + 1] += g.ptr[i];
    return g;
}

// one matching pass: returns the coarse graph and fills match[n] (coarse index of every node)
Graph match_pass(const Graph& g, std::vector<int32_t>& match) {
    const int n = g.n;
    match.assign(n, -1);
    std::vector<int32_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return g.ptr[a + 1] - g.ptr[a] < g.ptr[b + 1] - g.ptr[b]; });
    int nc = 0;
    for (int i : order) {
        if (match[i] >= 0) continue;
        int best = -1, bw = -1;
        for (int k = g.ptr[i]; k < g.ptr[i + 1]; ++k) {
            int j = g.idx[k];
            if (match[j] < 0 && j != i && g.wgt[k] > bw) {
                bw = g.wgt[k];
                best = j;
            }
        }
        match[i] = nc;
        if (best >= 0) match[best] = nc;
        ++nc;
    }
    // coarse graph: sum the weights of parallel connections
    std::vector<std::vector<std::pair<int32_t, int32_t>>> rows(nc);
    for (int i = 0; i < n; ++i)
        for (int k = g.ptr[i]; k < g.ptr[i + 1]; ++k) {
            int a = match[i], b = match[g.idx[k]];
            if (a != b) rows[a].emplace_back(b, g.wgt[k]);
        }
    Graph c;
    c.n = nc;
    c.ptr.assign(nc + 1, 0);
    for (int a = 0; a < nc; ++a) {
        auto& r = rows[a];
        std::sort(r.begin(), r.end());
        for (size_t k = 0; k < r.size();) {
            size_t e = k;
            int w = 0;
            while (e < r.size() && r[e].first == r[k].first) w += r[e++].second;
            c.idx.push_back(r[k].first);
            c.wgt.push_back(w);
            k = e;
        }
        c.ptr[a + 1] = (int)c.idx.size();
    }
    return c;
}

}  // namespace

// aggregation decisions only: agg[l] maps the rows of level l to the rows of level l+1
static std::vector<std::vector<int32_t>> aggregate_levels(int n, const std::vector<int32_t>& rowptr,
                                                          const std::vector<int32_t>& colidx, std::vector<int>& sizes) {
    constexpr int kCoarsest = 96;      // the coarsest operator is inverted densely in shared memory (pcg.cu, kCoarsestDense)
    std::vector<std::vector<int32_t>> aggs;
    Graph g = graph_from_csr(n, rowptr, colidx);
    sizes.assign(1, n);
    while (n > kCoarsest && (int)aggs.size() < kMaxMgLevels) {
        std::vector<int32_t> m1, m2;
        Graph g1 = match_pass(g, m1);
        Graph g2 = match_pass(g1, m2);
        if (g2.n >= n) break;                                  // no edges left to coarsen along
        std::vector<int32_t> agg(n);
        for (int i = 0; i < n; ++i) agg[i] = m2[m1[i]];
        aggs.push_back(std::move(agg));
        n = g2.n;
        sizes.push_back(n);
        g = std::move(g2);
        // the next level starts from the bare pattern of the coarse operator: carrying the coupling counts over
        // makes the first matching pass chase heavy connections and costs 50 % more CG iterations on grids
        std::fill(g.wgt.begin(), g.wgt.end(), 1);
    }
    return aggs;
}

// structures of one coarse level from the fine pattern and a prescribed aggregate map
static MgLevel build_level(int n, const std::vector<int32_t>& rowptr, const std::vector<int32_t>& colidx,
                           const std::vector<int32_t>& agg, int nc) {
    MgLevel L;
    L.n_fine = n;
    L.n = nc;
    L.agg = agg;
    L.mem_ptr.assign(L.n + 1, 0);
    for (int i = 0; i < n; ++i) L.mem_ptr[L.agg[i] + 1]++;
    for (int a = 0; a < L.n; ++a) L.mem_ptr[a + 1] += L.mem_ptr[a];
    L.mem_idx.resize(n);
    {
        std::vector<int32_t> cur(L.mem_ptr.begin(), L.mem_ptr.end() - 1);
        for (int i = 0; i < n; ++i) L.mem_idx[cur[L.agg[i]]++] = i;
    }
    // coarse pattern and the Galerkin gather lists from the fine pattern
    const int nnz_f = rowptr[n];
    std::vector<std::pair<int64_t, int32_t>> keyed(nnz_f);   // (coarse row * nc + coarse col, fine slot)
    for (int i = 0; i < n; ++i)
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) keyed[k] = {(int64_t)L.agg[i] * L.n + L.agg[colidx[k]], k};
    std::sort(keyed.begin(), keyed.end());
    L.rowptr.assign(L.n + 1, 0);
    L.gal_ptr.push_back(0);
    L.gal_src.resize(nnz_f);
    L.diagslot.assign(L.n, -1);
    for (int k = 0; k < nnz_f;) {
        int e = k;
        while (e < nnz_f && keyed[e].first == keyed[k].first) {
            L.gal_src[e] = keyed[e].second;
            ++e;
        }
        int row = (int)(keyed[k].first / L.n), col = (int)(keyed[k].first % L.n);
        if (row == col) L.diagslot[row] = (int)L.colidx.size();
        L.colidx.push_back(col);
        L.rowptr[row + 1]++;
        L.gal_ptr.push_back(e);
        k = e;
    }
    for (int a = 0; a < L.n; ++a) L.rowptr[a + 1] += L.rowptr[a];
    L.nnz = (int)L.colidx.size();
    for (int a = 0; a < L.n; ++a) L.ell_w = std::max(L.ell_w, L.rowptr[a + 1] - L.rowptr[a]);
    L.ell_col.assign((size_t)L.n * L.ell_w, 0);
    for (int a = 0; a < L.n; ++a)
        for (int u = 0; u < L.ell_w; ++u) {
            int k = L.rowptr[a] + u;
            L.ell_col[(size_t)a * L.ell_w + u] = k < L.rowptr[a + 1] ? L.colidx[k] : a;
        }
    return L;
}

// Pattern of K, its padded copy and the gather form of the scatter-add for a given order of the free nodes.
static void build_rowset(const Plan& P, const std::vector<int32_t>& free_nodes, RowSet& R) {
    const int n = P.n, nf = (int)free_nodes.size();
    const int32_t* edges = P.edges.data();
    R.nf = nf;
    R.free_nodes = free_nodes;
    R.node_src = P.node_src;                                   // fixed nodes keep -1 - k
    for (int r = 0; r < nf; ++r) R.node_src[free_nodes[r]] = r;
    (void)n;
    R.rowptr.assign(nf + 1, 0);
    std::vector<std::vector<int32_t>> cols(nf);
    for (int f = 0; f < nf; ++f) {
        int i = free_nodes[f];
        auto& c = cols[f];
        c.push_back(f);
        for (int k = P.ninc_ptr[i]; k < P.ninc_ptr[i + 1]; ++k) {
            int e = P.ninc_edge[k];
            int other = edges[2 * e] == i ? edges[2 * e + 1] : edges[2 * e];
            if (R.node_src[other] >= 0) c.push_back(R.node_src[other]);
        }
        std::sort(c.begin(), c.end());
        c.erase(std::unique(c.begin(), c.end()), c.end());
        R.rowptr[f + 1] = R.rowptr[f] + (int)c.size();
    }
    R.nnz = R.rowptr[nf];
    R.colidx.resize(R.nnz);
    R.diagslot.resize(nf);
    for (int f = 0; f < nf; ++f) {
        std::copy(cols[f].begin(), cols[f].end(), R.colidx.begin() + R.rowptr[f]);
        R.diagslot[f] = R.rowptr[f] + (int)(std::lower_bound(cols[f].begin(), cols[f].end(), f) - cols[f].begin());
    }
    auto slot_of = [&](int row, int col) {
        auto b = R.colidx.begin() + R.rowptr[row], e = R.colidx.begin() + R.rowptr[row + 1];
        return (int)(std::lower_bound(b, e, col) - R.colidx.begin());
    };
    // padded copy of the pattern for the SpMV kernels (indices depend on the row number only)
    R.ell_w = 0;
    for (int f = 0; f < nf; ++f) R.ell_w = std::max(R.ell_w, R.rowptr[f + 1] - R.rowptr[f]);
    R.ell_col.assign((size_t)nf * R.ell_w, 0);
    for (int f = 0; f < nf; ++f)
        for (int u = 0; u < R.ell_w; ++u) {
            int k = R.rowptr[f] + u;
            R.ell_col[(size_t)f * R.ell_w + u] = k < R.rowptr[f + 1] ? R.colidx[k] : f;
        }
    // free-row incidence (gather form of the scatter-add: one segment per row, fixed order)
    R.finc_ptr.assign(nf + 1, 0);
    for (int f = 0; f < nf; ++f) {
        int i = free_nodes[f];
        R.finc_ptr[f + 1] = R.finc_ptr[f] + (P.ninc_ptr[i + 1] - P.ninc_ptr[i]);
    }
    R.finc_edge.resize(R.finc_ptr[nf]);
    R.finc_code.resize(R.finc_ptr[nf]);
    for (int f = 0; f < nf; ++f) {
        int i = free_nodes[f];
        int o = R.finc_ptr[f];
        for (int k = P.ninc_ptr[i]; k < P.ninc_ptr[i + 1]; ++k, ++o) {
            int e = P.ninc_edge[k];
            int other = edges[2 * e] == i ? edges[2 * e + 1] : edges[2 * e];
            R.finc_edge[o] = e;
            R.finc_code[o] = R.node_src[other] >= 0 ? slot_of(f, R.node_src[other]) : R.node_src[other];
        }
        // parallel edges share a slot: keep them adjacent (code, then edge index) so the row sums them in a fixed order
        std::vector<std::pair<int32_t, int32_t>> seg;
        for (int k = R.finc_ptr[f]; k < R.finc_ptr[f + 1]; ++k) seg.emplace_back(R.finc_code[k], R.finc_edge[k]);
        std::sort(seg.begin(), seg.end());
        for (int k = R.finc_ptr[f], t = 0; k < R.finc_ptr[f + 1]; ++k, ++t) {
            R.finc_code[k] = seg[t].first;
            R.finc_edge[k] = seg[t].second;
        }
    }
    // padded incidence for the matrix-free product (K p)_f = sum_e q_e (p_f - p_other(e))
    R.inc_w = 0;
    for (int f = 0; f < nf; ++f) R.inc_w = std::max(R.inc_w, R.finc_ptr[f + 1] - R.finc_ptr[f]);
    if (R.inc_w > 16) R.inc_w = 0;                       // hubs: keep the stored matrix
    R.inc_edge.assign((size_t)nf * R.inc_w, -1);
    R.inc_col.assign((size_t)nf * R.inc_w, -1);
    for (int f = 0; f < nf && R.inc_w > 0; ++f)
        for (int k = R.finc_ptr[f], u = 0; k < R.finc_ptr[f + 1]; ++k, ++u) {
            R.inc_edge[(size_t)f * R.inc_w + u] = R.finc_edge[k];
            R.inc_col[(size_t)f * R.inc_w + u] = R.finc_code[k] >= 0 ? R.colidx[R.finc_code[k]] : -1;
        }
}

// The PCG path numbers its rows hierarchically: rows of one aggregate are contiguous, aggregates of one
// parent aggregate are contiguous, and so on up the hierarchy, so that a run of consecutive rows is a
// compact patch of the network at every level and neighbour gathers hit L1 instead of L2.
static void build_pcg_order(Plan& P) {
    std::vector<int> sizes;
    std::vector<std::vector<int32_t>> aggs = aggregate_levels(P.nf, P.rowptr, P.colidx, sizes);
    const int nl = (int)aggs.size();
    // newid[l][old row of level l] = position in the hierarchical order of level l (top-down)
    std::vector<std::vector<int32_t>> newid(nl + 1);
    newid[nl].resize(sizes[nl]);
    std::iota(newid[nl].begin(), newid[nl].end(), 0);
    for (int l = nl - 1; l >= 0; --l) {
        std::vector<int32_t> order(sizes[l]);
        std::iota(order.begin(), order.end(), 0);
        const auto& parent = aggs[l];
        const auto& ppos = newid[l + 1];
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ppos[parent[a]] < ppos[parent[b]]; });
        newid[l].resize(sizes[l]);
        for (int k = 0; k < sizes[l]; ++k) newid[l][order[k]] = k;
    }
    P.pg_perm.resize(P.nf);
    for (int f = 0; f < P.nf; ++f) P.pg_perm[newid[0][f]] = f;
    std::vector<int32_t> free_nodes(P.nf);
    for (int r = 0; r < P.nf; ++r) free_nodes[r] = P.free_nodes[P.pg_perm[r]];
    build_rowset(P, free_nodes, P.pg);
    // coarse levels on the relabelled rows with the relabelled aggregate maps
    std::vector<int32_t> rowptr = P.pg.rowptr, colidx = P.pg.colidx;
    for (int l = 0; l < nl; ++l) {
        std::vector<int32_t> agg(sizes[l]);
        for (int i = 0; i < sizes[l]; ++i) agg[newid[l][i]] = newid[l + 1][aggs[l][i]];
        MgLevel L = build_level(sizes[l], rowptr, colidx, agg, sizes[l + 1]);
        // hierarchical order: the members of an aggregate are consecutive rows (k_mg_down relies on it)
        for (int m = 0; m < sizes[l]; ++m)
            if (L.mem_idx[m] != m) set_error("internal: aggregate members are not contiguous");
        rowptr = L.rowptr;
        colidx = L.colidx;
        P.mg.push_back(std::move(L));
    }
    if (!P.mg.empty()) {
        P.mg_rows.push_back(P.nf);
        P.mg_nnz.push_back(P.nnz);
        for (const MgLevel& L : P.mg) {
            P.mg_rows.push_back(L.n);
            P.mg_nnz.push_back(L.nnz);
        }
    }
}

int build_plan(int32_t n, int32_t E, const int32_t* edges, const uint8_t* is_fixed, Plan& P) {
    if (n <= 0 || E < 0 || (E > 0 && !edges) || !is_fixed) {
        set_error("dfdm_plan_create: n must be positive and edges / is_fixed non-NULL");
        return DFDM_EINVAL;
    }
    for (int e = 0; e < E; ++e) {
        int u = edges[2 * e], v = edges[2 * e + 1];
        if (u < 0 || u >= n || v < 0 || v >= n) {
            set_error("dfdm_plan_create: edge " + std::to_string(e) + " references a node outside [0, n)");
            return DFDM_EINVAL;
        }
    }
    P.n = n;
    P.E = E;
    P.edges.assign(edges, edges + 2 * (size_t)E);

    // structure.py:74-106
    P.node_src.assign(n, 0);
    for (int i = 0; i < n; ++i) {
        if (is_fixed[i]) {
            P.node_src[i] = -1 - (int)P.fixed_nodes.size();
            P.fixed_nodes.push_back(i);
        } else {
            P.node_src[i] = (int)P.free_nodes.size();
            P.free_nodes.push_back(i);
        }
    }
    P.nf = (int)P.free_nodes.size();
    P.nx = (int)P.fixed_nodes.size();
    if (P.nf == 0) {
        set_error("dfdm_plan_create: the network has no free node");
        return DFDM_EINVAL;
    }
    P.freefixed.resize(n);
    for (int i = 0; i < n; ++i) P.freefixed[i] = P.node_src[i] >= 0 ? P.node_src[i] : P.nf + (-1 - P.node_src[i]);

    // node -> incident edges with C[e, node] (self-loops cancel in C and are skipped)
    P.ninc_ptr.assign(n + 1, 0);
    for (int e = 0; e < E; ++e) {
        int u = edges[2 * e], v = edges[2 * e + 1];
        if (u == v) continue;
        P.ninc_ptr[u + 1]++;
        P.ninc_ptr[v + 1]++;
    }
    for (int i = 0; i < n; ++i) P.ninc_ptr[i + 1] += P.ninc_ptr[i];
    P.ninc_edge.resize(P.ninc_ptr[n]);
    P.ninc_sign.resize(P.ninc_ptr[n]);
    {
        std::vector<int32_t> cursor(P.ninc_ptr.begin(), P.ninc_ptr.end() - 1);
        for (int e = 0; e < E; ++e) {
            int u = edges[2 * e], v = edges[2 * e + 1];
            if (u == v) continue;
            P.ninc_edge[cursor[u]] = e;
            P.ninc_sign[cursor[u]++] = -1;
            P.ninc_edge[cursor[v]] = e;
            P.ninc_sign[cursor[v]++] = +1;
        }
    }

    // pattern of K = C_f^T Q C_f in ascending free order (the order of the reference; exposed through the ABI)
    const int nf = P.nf;
    RowSet nat;
    build_rowset(P, P.free_nodes, nat);
    P.nnz = nat.nnz;
    P.rowptr = nat.rowptr;
    P.colidx = nat.colidx;
    P.diagslot = nat.diagslot;
    P.ell_w = nat.ell_w;
    P.ell_col = nat.ell_col;
    P.finc_ptr = nat.finc_ptr;
    P.finc_edge = nat.finc_edge;
    P.finc_code = nat.finc_code;
    auto slot_of = [&](int row, int col) {
        auto b = P.colidx.begin() + P.rowptr[row], e = P.colidx.begin() + P.rowptr[row + 1];
        return (int)(std::lower_bound(b, e, col) - P.colidx.begin());
    };

    build_pcg_order(P);

    // per-edge scatter slots (uu, vv, uv, vu)
    P.edge_slots.assign(4 * (size_t)E, -1);
    for (int e = 0; e < E; ++e) {
        int u = edges[2 * e], v = edges[2 * e + 1];
        if (u == v) continue;
        int fu = P.node_src[u], fv = P.node_src[v];
        if (fu >= 0) P.edge_slots[4 * e + 0] = P.diagslot[fu];
        if (fv >= 0) P.edge_slots[4 * e + 1] = P.diagslot[fv];
        if (fu >= 0 && fv >= 0) {
            P.edge_slots[4 * e + 2] = slot_of(fu, fv);
            P.edge_slots[4 * e + 3] = slot_of(fv, fu);
        }
    }

    // ordering for the banded direct solver
    std::vector<int32_t> ident(nf);
    std::iota(ident.begin(), ident.end(), 0);
    P.w_nat = bandwidth(nf, P.rowptr, P.colidx, ident);
    std::vector<int32_t> rperm, rpos(nf);
    rcm_order(nf, P.rowptr, P.colidx, rperm);
    for (int p = 0; p < nf; ++p) rpos[rperm[p]] = p;
    int w_rcm = bandwidth(nf, P.rowptr, P.colidx, rpos);
    if (w_rcm < P.w_nat) {
        P.use_rcm = true;
        P.perm = rperm;
        P.pos = rpos;
        P.w = w_rcm;
    } else {
        P.perm = ident;
        P.pos = ident;
        P.w = P.w_nat;
    }

    // incidence by solver position with band offsets
    P.sinc_ptr.assign(nf + 1, 0);
    for (int p = 0; p < nf; ++p) {
        int f = P.perm[p];
        P.sinc_ptr[p + 1] = P.sinc_ptr[p] + (P.finc_ptr[f + 1] - P.finc_ptr[f]);
    }
    P.sinc_edge.resize(P.sinc_ptr[nf]);
    P.sinc_code.resize(P.sinc_ptr[nf]);
    for (int p = 0; p < nf; ++p) {
        int f = P.perm[p];
        int i = P.free_nodes[f];
        int o = P.sinc_ptr[p];
        for (int k = P.ninc_ptr[i]; k < P.ninc_ptr[i + 1]; ++k, ++o) {
            int e = P.ninc_edge[k];
            int other = edges[2 * e] == i ? edges[2 * e + 1] : edges[2 * e];
            P.sinc_edge[o] = e;
            if (P.node_src[other] >= 0) {
                int pj = P.pos[P.node_src[other]];
                P.sinc_code[o] = pj < p ? p - pj : 0;
            } else {
                P.sinc_code[o] = P.node_src[other];
            }
        }
    }
    return DFDM_OK;
}

}  // namespace dfdm

namespace dfdm { const char* get_error(); }

extern "C" {

const char* dfdm_last_error(void) { return dfdm::get_error(); }
int32_t dfdm_abi_version(void) { return DFDM_ABI_VERSION; }

int dfdm_plan_get_info(const dfdm_plan* plan, dfdm_plan_info* info) {
    if (!plan || !info) {
        dfdm::set_error("dfdm_plan_get_info: NULL argument");
        return DFDM_EINVAL;
    }
    std::memset(info, 0, sizeof(*info));
    info->n = plan->n;
    info->n_edges = plan->E;
    info->n_free = plan->nf;
    info->n_fixed = plan->nx;
    info->nnz = plan->nnz;
    info->bandwidth_natural = plan->w_nat;
    info->bandwidth_solver = plan->w;
    info->direct_smem_bytes = dfdm::direct_smem_bytes(plan->nf, plan->w);
    info->direct_fits = info->direct_smem_bytes <= dfdm::kDirectSmemLimit ? 1 : 0;
    info->auto_solver = info->direct_fits ? DFDM_SOLVER_DIRECT : DFDM_SOLVER_PCG;
    info->mg_levels = (int32_t)plan->mg.size();
    info->mg_n_coarse = plan->mg.empty() ? 0 : plan->mg.front().n;
    info->mg_n_coarsest = plan->mg.empty() ? 0 : plan->mg.back().n;
    return DFDM_OK;
}

const int32_t* dfdm_plan_array(const dfdm_plan* plan, int32_t which, int64_t* length) {
    if (!plan) return nullptr;
    const std::vector<int32_t>* v = nullptr;
    switch (which) {
        case DFDM_ARR_FREE: v = &plan->free_nodes; break;
        case DFDM_ARR_FIXED: v = &plan->fixed_nodes; break;
        case DFDM_ARR_FREEFIXED: v = &plan->freefixed; break;
        case DFDM_ARR_ROWPTR: v = &plan->rowptr; break;
        case DFDM_ARR_COLIDX: v = &plan->colidx; break;
        case DFDM_ARR_EDGE_SLOTS: v = &plan->edge_slots; break;
        case DFDM_ARR_SOLVER_PERM: v = &plan->perm; break;
        case DFDM_ARR_NODE_INC_PTR: v = &plan->ninc_ptr; break;
        case DFDM_ARR_NODE_INC_EDGE: v = &plan->ninc_edge; break;
        case DFDM_ARR_NODE_INC_SIGN: v = &plan->ninc_sign; break;
        case DFDM_ARR_MG_ROWS: v = &plan->mg_rows; break;
        case DFDM_ARR_MG_NNZ: v = &plan->mg_nnz; break;
        case DFDM_ARR_PCG_PERM: v = &plan->pg_perm; break;
        default:
            if (which >= DFDM_ARR_MG_AGG && which < DFDM_ARR_MG_AGG + (int32_t)plan->mg.size()) {
                v = &plan->mg[which - DFDM_ARR_MG_AGG].agg;
                break;
            }
            dfdm::set_error("dfdm_plan_array: unknown array id");
            return nullptr;
    }
    if (length) *length = (int64_t)v->size();
    return v->data();
}

}  // extern "C"
