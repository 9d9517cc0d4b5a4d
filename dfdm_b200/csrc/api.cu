// C ABI (include/dfdm_b200.h): plan / goal-program lifetime, device mirrors, dispatch.
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include "common.cuh"

namespace dfdm {

std::atomic<long long> g_launches{0};
void last_pcg_iterations(int* fwd, int* adj);
const char* last_direct_kernel();
void profile_enable(int on);
void profile_read(double* ms, long long* count);

// ---- device mirrors ------------------------------------------------------------------------------

static int upload_plan(const Plan& P, int device, PlanDev& out, PlanDev& out_pcg) {
    std::lock_guard<std::mutex> lock(P.mu);
    auto it = P.dev.find(device);
    if (it != P.dev.end()) {
        out = it->second;
        out_pcg = P.dev_pcg[device];
        return DFDM_OK;
    }
    const std::vector<int32_t>* arrays[] = {&P.edges, &P.node_src, &P.free_nodes, &P.rowptr, &P.colidx, &P.diagslot,
                                            &P.finc_ptr, &P.finc_edge, &P.finc_code, &P.ninc_ptr, &P.ninc_edge, &P.ninc_sign,
                                            &P.perm, &P.pos, &P.sinc_ptr, &P.sinc_edge, &P.sinc_code, &P.ell_col,
                                            &P.pg.free_nodes, &P.pg.node_src, &P.pg.rowptr, &P.pg.colidx, &P.pg.diagslot, &P.pg.ell_col,
                                            &P.pg.finc_ptr, &P.pg.finc_edge, &P.pg.finc_code, &P.pg.inc_edge, &P.pg.inc_col};
    constexpr int NA = sizeof(arrays) / sizeof(arrays[0]);
    int64_t offs[NA], total = 0;
    for (int k = 0; k < NA; ++k) {
        offs[k] = total;
        total += ((int64_t)arrays[k]->size() * 4 + 255) & ~(int64_t)255;
    }
    char* block = nullptr;
    DFDM_CUDA_OK(cudaMalloc(&block, (size_t)std::max<int64_t>(total, 256)));
    for (int k = 0; k < NA; ++k)
        if (!arrays[k]->empty())
            DFDM_CUDA_OK(cudaMemcpy(block + offs[k], arrays[k]->data(), arrays[k]->size() * 4, cudaMemcpyHostToDevice));
    auto at = [&](int k) { return reinterpret_cast<const int32_t*>(block + offs[k]); };
    PlanDev d;
    d.n = P.n; d.E = P.E; d.nf = P.nf; d.nx = P.nx; d.nnz = P.nnz; d.w = P.w;
    d.edges = at(0); d.node_src = at(1); d.free_nodes = at(2); d.rowptr = at(3); d.colidx = at(4); d.diagslot = at(5);
    d.finc_ptr = at(6); d.finc_edge = at(7); d.finc_code = at(8); d.ninc_ptr = at(9); d.ninc_edge = at(10); d.ninc_sign = at(11);
    d.perm = at(12); d.pos = at(13); d.sinc_ptr = at(14); d.sinc_edge = at(15); d.sinc_code = at(16);
    d.ell_w = P.ell_w; d.ell_col = at(17);
    P.dev[device] = d;
    PlanDev g = d;                                    // the PCG path's view: same network, rows in its own order
    g.free_nodes = at(18); g.node_src = at(19); g.rowptr = at(20); g.colidx = at(21); g.diagslot = at(22);
    g.ell_w = P.pg.ell_w; g.ell_col = at(23); g.finc_ptr = at(24); g.finc_edge = at(25); g.finc_code = at(26);
    g.inc_w = P.pg.inc_w; g.inc_edge = at(27); g.inc_col = at(28);
    g.perm = nullptr; g.pos = nullptr; g.sinc_ptr = nullptr; g.sinc_edge = nullptr; g.sinc_code = nullptr;
    P.dev_pcg[device] = g;
    P.dev_block[device] = block;
    out = d;
    out_pcg = g;
    return DFDM_OK;
}

static int upload_mg(const Plan& P, int device, MgDev& out) {
    std::lock_guard<std::mutex> lock(P.mu);
    auto it = P.mg_dev.find(device);
    if (it != P.mg_dev.end()) {
        out = it->second;
        return DFDM_OK;
    }
    MgDev M;
    M.n_levels = (int)P.mg.size();
    for (int l = 0; l < M.n_levels; ++l) {
        const MgLevel& L = P.mg[l];
        // sliced ELL tables of this level's operator (k_mg_tail): slices of 32 rows, each as wide as its widest row
        const int slices = (L.n + 31) / 32;
        std::vector<int32_t> tl_off(slices + 1, 0), tl_col, tl_src, tl_agg;
        for (int sl = 0; sl < slices; ++sl) {
            int wmax = 0;
            for (int f = sl * 32; f < std::min(L.n, sl * 32 + 32); ++f) wmax = std::max(wmax, L.rowptr[f + 1] - L.rowptr[f]);
            tl_off[sl + 1] = tl_off[sl] + wmax;
        }
        const int64_t ent = (int64_t)tl_off[slices] * 32;
        tl_col.assign((size_t)ent, 0);
        tl_src.assign((size_t)ent, -1);
        const bool has_next = l + 1 < M.n_levels;
        if (has_next) tl_agg.assign((size_t)ent, 0);
        for (int sl = 0; sl < slices; ++sl)
            for (int u = 0; u < tl_off[sl + 1] - tl_off[sl]; ++u)
                for (int lane = 0; lane < 32; ++lane) {
                    const int f = sl * 32 + lane;
                    const size_t at = ((size_t)tl_off[sl] + u) * 32 + lane;
                    int col = std::min(f, L.n - 1);
                    if (f < L.n && L.rowptr[f] + u < L.rowptr[f + 1]) {
                        col = L.colidx[L.rowptr[f] + u];
                        tl_src[at] = L.rowptr[f] + u;
                    }
                    tl_col[at] = col;
                    if (has_next) tl_agg[at] = P.mg[l + 1].agg[col];
                }
        const std::vector<int32_t>* arrays[] = {&L.agg, &L.mem_ptr, &L.mem_idx, &L.rowptr, &L.ell_col, &L.diagslot, &L.gal_ptr, &L.gal_src,
                                                &tl_off, &tl_col, &tl_src, &tl_agg};
        constexpr int NA = sizeof(arrays) / sizeof(arrays[0]);
        int64_t offs[NA], total = 0;
        for (int k = 0; k < NA; ++k) {
            offs[k] = total;
            total += ((int64_t)arrays[k]->size() * 4 + 255) & ~(int64_t)255;
        }
        char* block = nullptr;
        DFDM_CUDA_OK(cudaMalloc(&block, (size_t)std::max<int64_t>(total, 256)));
        for (int k = 0; k < NA; ++k)
            if (!arrays[k]->empty())
                DFDM_CUDA_OK(cudaMemcpy(block + offs[k], arrays[k]->data(), arrays[k]->size() * 4, cudaMemcpyHostToDevice));
        auto at = [&](int k) { return reinterpret_cast<const int32_t*>(block + offs[k]); };
        MgLevelDev& D = M.lv[l];
        D.n_fine = L.n_fine; D.n = L.n; D.nnz = L.nnz; D.ell_w = L.ell_w;
        D.agg = at(0); D.mem_ptr = at(1); D.mem_idx = at(2); D.rowptr = at(3); D.ell_col = at(4); D.diagslot = at(5);
        D.gal_ptr = at(6); D.gal_src = at(7);
        D.tl_slices = slices; D.tl_ent = ent; D.tl_off = at(8); D.tl_col = at(9); D.tl_src = at(10);
        D.tl_agg = has_next ? at(11) : nullptr;
        P.dev_block[1000 * (l + 1) + device] = block;     // freed with the plan (key unique per level and device)
    }
    P.mg_dev[device] = M;
    out = M;
    return DFDM_OK;
}

static int upload_prog(const GoalProg& G, int device, GoalProgDev& out) {
    std::lock_guard<std::mutex> lock(G.mu);
    auto it = G.dev.find(device);
    if (it != G.dev.end()) {
        out = it->second;
        return DFDM_OK;
    }
    int64_t rec_bytes = ((int64_t)G.recs.size() * sizeof(GoalRec) + 255) & ~(int64_t)255;
    int64_t np_bytes = ((int64_t)G.ngoal_ptr.size() * 4 + 255) & ~(int64_t)255;
    int64_t ep_bytes = ((int64_t)G.egoal_ptr.size() * 4 + 255) & ~(int64_t)255;
    char* block = nullptr;
    DFDM_CUDA_OK(cudaMalloc(&block, (size_t)(rec_bytes + np_bytes + ep_bytes + 256)));
    if (!G.recs.empty()) DFDM_CUDA_OK(cudaMemcpy(block, G.recs.data(), G.recs.size() * sizeof(GoalRec), cudaMemcpyHostToDevice));
    DFDM_CUDA_OK(cudaMemcpy(block + rec_bytes, G.ngoal_ptr.data(), G.ngoal_ptr.size() * 4, cudaMemcpyHostToDevice));
    DFDM_CUDA_OK(cudaMemcpy(block + rec_bytes + np_bytes, G.egoal_ptr.data(), G.egoal_ptr.size() * 4, cudaMemcpyHostToDevice));
    GoalProgDev d;
    d.n_goals = (int)G.recs.size();
    d.n_net = G.n_net;
    d.net_begin = G.net_begin;
    d.recs = reinterpret_cast<const GoalRec*>(block);
    d.ngoal_ptr = reinterpret_cast<const int32_t*>(block + rec_bytes);
    d.egoal_ptr = reinterpret_cast<const int32_t*>(block + rec_bytes + np_bytes);
    d.l2_alpha = G.l2_alpha;
    G.dev[device] = d;
    G.dev_block[device] = block;
    out = d;
    return DFDM_OK;
}

static int resolve_solver(const Plan& P, const dfdm_options* opt, int& solver) {
    // AUTO: the fused shared-memory kernel when the banded factor fits a CTA's shared memory, PCG beyond.  DIRECT asked for
    // by name beyond that limit takes the global-memory variant of the same factorisation (direct.cu: k_direct_global) - an
    // order of magnitude slower per design than PCG on a definite K, but the one path that solves an indefinite K
    // (mixed-sign q) at any size, as the reference's LU does (model.py:55).
    bool fits = direct_smem_bytes(P.nf, P.w) <= kDirectSmemLimit;
    solver = opt ? opt->solver : DFDM_SOLVER_AUTO;
    if (solver == DFDM_SOLVER_AUTO) solver = fits ? DFDM_SOLVER_DIRECT : DFDM_SOLVER_PCG;
    if (solver != DFDM_SOLVER_DIRECT && solver != DFDM_SOLVER_PCG) {
        set_error("unknown solver id");
        return DFDM_EINVAL;
    }
    return DFDM_OK;
}

static int dispatch(const dfdm_plan* plan, const dfdm_goalprog* prog, Eval& ev, const dfdm_options* opt) {
    if (!plan || ev.B <= 0 || !ev.q || !ev.loads || (plan->nx > 0 && !ev.xfix)) {
        set_error("NULL plan / inputs or non-positive batch");
        return DFDM_EINVAL;
    }
    if (prog && (prog->n != plan->n || prog->E != plan->E)) {
        set_error("goal program was compiled for a different plan");
        return DFDM_EINVAL;
    }
    if (ev.adjoint && !ev.dq) {
        set_error("gradient requested without an output buffer");
        return DFDM_EINVAL;
    }
    int device = 0;
    DFDM_CUDA_OK(cudaGetDevice(&device));
    int rc = upload_plan(*plan, device, ev.P, ev.Pg);
    if (rc) return rc;
    rc = upload_mg(*plan, device, ev.M);
    if (rc) return rc;
    ev.has_prog = prog != nullptr;
    if (prog) {
        rc = upload_prog(*prog, device, ev.G);
        if (rc) return rc;
    }
    if (opt) ev.opt = *opt;
    int solver = 0;
    rc = resolve_solver(*plan, opt, solver);
    if (rc) return rc;
    return solver == DFDM_SOLVER_DIRECT ? run_direct(ev) : run_pcg(ev);
}

// cached per-device arenas of the *_host entry points (created on first use; the caller holds `busy` of the slot it uses)
static Plan::ArenaSet& arena_ref(const Plan& P, int device) {
    std::lock_guard<std::mutex> lock(P.mu);
    return P.arena[device];
}

// Host calls from several threads share a device by taking turns with its SMs: the kernels of one call at a time (two
// PCG solves interleaved would each run slower and double the working set), while the copies of the other calls - on
// their own streams, into their own arenas - run underneath.
static std::mutex& compute_turn(int device) {
    static std::mutex mu;
    static std::map<int, std::mutex> turns;
    std::lock_guard<std::mutex> lock(mu);
    return turns[device];
}

static int arena_prepare(Plan::Arena& a, int64_t bytes) {
    if (a.bytes < bytes) {
        if (a.ptr) cudaFree(a.ptr);
        a.ptr = nullptr;
        a.bytes = 0;
        DFDM_CUDA_OK(cudaMalloc(&a.ptr, (size_t)bytes));
        a.bytes = bytes;
    }
    for (int k = 0; k < 3; ++k)
        if (!a.stream[k]) {
            cudaStream_t st = nullptr;
            DFDM_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            a.stream[k] = st;
        }
    for (int k = 0; k < 3; ++k)
        if (!a.ev_end[k]) {
            cudaEvent_t e = nullptr;
            DFDM_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
            a.ev_end[k] = e;
        }
    for (int k = 0; k < Plan::kHostChunksMax; ++k) {
        if (!a.ev_in[k]) {
            cudaEvent_t e = nullptr;
            DFDM_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            a.ev_in[k] = e;
        }
        if (!a.ev_done[k]) {
            cudaEvent_t e = nullptr;
            DFDM_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            a.ev_done[k] = e;
        }
    }
    return DFDM_OK;
}

// How a host call splits its batch: the copies of chunk k+1 (in) and chunk k-1 (out) run under the kernels of chunk k,
// so only the first chunk's way in and the last chunk's way out are exposed.  Chunks are multiples of 32 designs where
// the batch allows, so every chunk keeps whole warps of design lanes.  DFDM_HOST_CHUNKS overrides the count (tuning).
static int host_chunks(int B, bool pcg, int64_t bytes_per_design, int* b0 /* [kHostChunksMax + 1] */) {
    // Measured (profiles/README.md, round 2): on the PCG path a batch cut in two evaluates 26 % slower than whole (its
    // small multigrid levels are latency bound and do not shrink with the batch), far more than the hidden copies return,
    // so PCG calls stay whole.  The direct path is one kernel per chunk with no host polling: cutting pays once a chunk's
    // copies outweigh a launch (>= 32 MB of q per chunk).
    static const int forced = std::getenv("DFDM_HOST_CHUNKS") ? std::atoi(std::getenv("DFDM_HOST_CHUNKS")) : 0;
    int k = forced > 0 ? forced : (pcg ? 1 : (int)std::min<int64_t>(4, ((int64_t)B * bytes_per_design) / (32ll << 20)));
    k = std::max(1, std::min(std::min(k, Plan::kHostChunksMax), B));
    const int unit = B >= 64 * k ? 32 : (B >= 4 * k ? 2 : 1);
    const int units = (B + unit - 1) / unit;
    for (int c = 0; c <= k; ++c) b0[c] = std::min(B, (int)(((int64_t)units * c / k) * unit));
    b0[k] = B;
    return k;
}

}  // namespace dfdm

using namespace dfdm;

extern "C" {

int dfdm_plan_create(int32_t n, int32_t n_edges, const int32_t* edges, const uint8_t* is_fixed, dfdm_plan** plan) {
    if (!plan) {
        set_error("dfdm_plan_create: NULL output pointer");
        return DFDM_EINVAL;
    }
    *plan = nullptr;
    dfdm_plan* p = new dfdm_plan();
    int rc = build_plan(n, n_edges, edges, is_fixed, *p);
    if (rc) {
        delete p;
        return rc;
    }
    *plan = p;
    return DFDM_OK;
}

void dfdm_plan_destroy(dfdm_plan* plan) {
    if (!plan) return;
    for (auto& kv : plan->dev_block) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(kv.first % 1000);
        cudaFree(kv.second);
        cudaSetDevice(prev);
    }
    for (auto& kv : plan->arena) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(kv.first);
        for (Plan::Arena& a : kv.second.slot) {
            if (a.ptr) cudaFree(a.ptr);
            for (void* st : a.stream)
                if (st) cudaStreamDestroy(static_cast<cudaStream_t>(st));
            for (void* e : a.ev_end)
                if (e) cudaEventDestroy(static_cast<cudaEvent_t>(e));
            for (int k = 0; k < Plan::kHostChunksMax; ++k) {
                if (a.ev_in[k]) cudaEventDestroy(static_cast<cudaEvent_t>(a.ev_in[k]));
                if (a.ev_done[k]) cudaEventDestroy(static_cast<cudaEvent_t>(a.ev_done[k]));
            }
        }
        cudaSetDevice(prev);
    }
    delete plan;
}

int dfdm_goalprog_create(const dfdm_plan* plan, int32_t n_goals, const int32_t* kind, const int32_t* element,
                         const int32_t* term, const double* weight, const double* params, int32_t n_terms,
                         const int32_t* term_kind, const double* term_alpha, double l2_alpha, dfdm_goalprog** prog) {
    if (!plan || !prog || n_goals < 0 || n_terms < 0 || (n_goals > 0 && (!kind || !element || !term || !weight || !params)) ||
        (n_terms > 0 && (!term_kind || !term_alpha))) {
        set_error("dfdm_goalprog_create: NULL or negative argument");
        return DFDM_EINVAL;
    }
    *prog = nullptr;
    // entries per term (vector goals count 3) for the mean of MeanSquaredError (losses.py:55)
    std::vector<int64_t> entries(n_terms, 0);
    for (int g = 0; g < n_goals; ++g) {
        if (kind[g] < 0 || kind[g] >= DFDM_N_GOAL_KINDS || term[g] < 0 || term[g] >= n_terms) {
            set_error("dfdm_goalprog_create: goal " + std::to_string(g) + " has an unknown kind or term");
            return DFDM_EINVAL;
        }
        bool node = kind[g] <= DFDM_GOAL_NODE_RESIDUAL_DIRECTION, net = kind[g] == DFDM_GOAL_NETWORK_LOAD_PATH;
        if (!net && (element[g] < 0 || element[g] >= (node ? plan->n : plan->E))) {
            set_error("dfdm_goalprog_create: goal " + std::to_string(g) + " references an element out of range");
            return DFDM_EINVAL;
        }
        bool vec = kind[g] == DFDM_GOAL_NODE_POINT || kind[g] == DFDM_GOAL_NODE_LINE || kind[g] == DFDM_GOAL_NODE_PLANE ||
                   kind[g] == DFDM_GOAL_NODE_RESIDUAL_VECTOR || kind[g] == DFDM_GOAL_NODE_RESIDUAL_DIRECTION ||
                   kind[g] == DFDM_GOAL_EDGE_DIRECTION;
        entries[term[g]] += vec ? 3 : 1;
    }
    for (int t = 0; t < n_terms; ++t)
        if (term_kind[t] < DFDM_TERM_SQUARED_ERROR || term_kind[t] > DFDM_TERM_PREDICTION_ERROR) {
            set_error("dfdm_goalprog_create: unknown term kind");
            return DFDM_EINVAL;
        }
    dfdm_goalprog* G = new dfdm_goalprog();
    G->n = plan->n;
    G->E = plan->E;
    G->l2_alpha = l2_alpha;
    // stable bucket sort: node goals by node, edge goals by edge, then network goals
    std::vector<int32_t> order(n_goals);
    for (int g = 0; g < n_goals; ++g) order[g] = g;
    auto key = [&](int g) -> int64_t {
        if (kind[g] <= DFDM_GOAL_NODE_RESIDUAL_DIRECTION) return element[g];
        if (kind[g] == DFDM_GOAL_NETWORK_LOAD_PATH) return (int64_t)plan->n + plan->E;
        return (int64_t)plan->n + element[g];
    };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
    G->ngoal_ptr.assign(plan->n + 1, 0);
    G->egoal_ptr.assign(plan->E + 1, 0);
    G->recs.resize(n_goals);
    int n_node = 0, n_edge = 0;
    for (int s = 0; s < n_goals; ++s) {
        int g = order[s];
        GoalRec& r = G->recs[s];
        r.kind = kind[g];
        r.elem = element[g];
        int tk = term_kind[term[g]];
        r.mode = tk == DFDM_TERM_PREDICTION_ERROR ? 1 : 0;
        r.pad = 0;
        r.scale = term_alpha[term[g]];
        if (tk == DFDM_TERM_MEAN_SQUARED_ERROR) r.scale /= (double)entries[term[g]];
        r.weight = weight[g];
        for (int k = 0; k < DFDM_GOAL_NPARAMS; ++k) r.p[k] = params[(size_t)g * DFDM_GOAL_NPARAMS + k];
        if (kind[g] <= DFDM_GOAL_NODE_RESIDUAL_DIRECTION) {
            G->ngoal_ptr[element[g] + 1]++;
            n_node++;
        } else if (kind[g] != DFDM_GOAL_NETWORK_LOAD_PATH) {
            G->egoal_ptr[element[g] + 1]++;
            n_edge++;
        }
    }
    for (int i = 0; i < plan->n; ++i) G->ngoal_ptr[i + 1] += G->ngoal_ptr[i];
    G->egoal_ptr[0] = n_node;
    for (int e = 0; e < plan->E; ++e) G->egoal_ptr[e + 1] += G->egoal_ptr[e];
    G->net_begin = n_node + n_edge;
    G->n_net = n_goals - G->net_begin;
    *prog = G;
    return DFDM_OK;
}

void dfdm_goalprog_destroy(dfdm_goalprog* prog) {
    if (!prog) return;
    for (auto& kv : prog->dev_block) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(kv.first);
        cudaFree(kv.second);
        cudaSetDevice(prev);
    }
    delete prog;
}

int64_t dfdm_workspace_bytes(const dfdm_plan* plan, int32_t batch, const dfdm_options* options) {
    if (!plan || batch <= 0) {
        set_error("dfdm_workspace_bytes: NULL plan or non-positive batch");
        return DFDM_EINVAL;
    }
    int solver = 0;
    int rc = resolve_solver(*plan, options, solver);
    if (rc) return rc;
    return solver == DFDM_SOLVER_DIRECT ? direct_workspace(*plan, batch, true, true) : pcg_workspace(*plan, batch);
}

int dfdm_forward(const dfdm_plan* plan, int32_t batch, const double* q, const double* loads, const double* xyz_fixed,
                 double* xyz, double* residuals, double* vectors, double* lengths, double* forces, int32_t* status,
                 void* workspace, int64_t workspace_bytes, const dfdm_options* options, void* stream) {
    Eval ev;
    ev.B = batch; ev.q = q; ev.loads = loads; ev.xfix = xyz_fixed;
    ev.o_xyz = xyz; ev.o_res = residuals; ev.o_vec = vectors; ev.o_len = lengths; ev.o_for = forces;
    ev.status = status; ev.ws = static_cast<char*>(workspace); ev.ws_bytes = workspace_bytes;
    ev.stream = static_cast<cudaStream_t>(stream);
    return dispatch(plan, nullptr, ev, options);
}

int dfdm_loss_and_grad(const dfdm_plan* plan, const dfdm_goalprog* prog, int32_t batch, const double* q, const double* loads,
                       const double* xyz_fixed, double* loss, double* dq, double* xyz, double* residuals, double* vectors,
                       double* lengths, double* forces, int32_t* status, void* workspace, int64_t workspace_bytes,
                       const dfdm_options* options, void* stream) {
    if (!prog || !loss) {
        set_error("dfdm_loss_and_grad: NULL goal program or loss buffer");
        return DFDM_EINVAL;
    }
    Eval ev;
    ev.B = batch; ev.q = q; ev.loads = loads; ev.xfix = xyz_fixed;
    ev.o_xyz = xyz; ev.o_res = residuals; ev.o_vec = vectors; ev.o_len = lengths; ev.o_for = forces;
    ev.loss = loss; ev.dq = dq; ev.adjoint = dq != nullptr;
    ev.status = status; ev.ws = static_cast<char*>(workspace); ev.ws_bytes = workspace_bytes;
    ev.stream = static_cast<cudaStream_t>(stream);
    return dispatch(plan, prog, ev, options);
}

int dfdm_vjp(const dfdm_plan* plan, int32_t batch, const double* q, const double* loads, const double* xyz_fixed,
             const double* xyz_bar, const double* residuals_bar, const double* vectors_bar, const double* lengths_bar,
             const double* forces_bar, double* dq, int32_t* status, void* workspace, int64_t workspace_bytes,
             const dfdm_options* options, void* stream) {
    Eval ev;
    ev.B = batch; ev.q = q; ev.loads = loads; ev.xfix = xyz_fixed;
    ev.c_xyz = xyz_bar; ev.c_res = residuals_bar; ev.c_vec = vectors_bar; ev.c_len = lengths_bar; ev.c_for = forces_bar;
    ev.dq = dq; ev.adjoint = true;
    ev.status = status; ev.ws = static_cast<char*>(workspace); ev.ws_bytes = workspace_bytes;
    ev.stream = static_cast<cudaStream_t>(stream);
    return dispatch(plan, nullptr, ev, options);
}

// ---- host-buffer variants -------------------------------------------------------------------------

static int host_eval(const dfdm_plan* plan, const dfdm_goalprog* prog, int32_t B, const double* q, const double* loads,
                     const double* xyz_fixed, double* loss, double* dq, double* xyz, double* residuals, double* vectors,
                     double* lengths, double* forces, int32_t* status, const dfdm_options* options) {
    if (!plan || B <= 0 || !q || !loads) {
        set_error("host evaluation: NULL plan / inputs or non-positive batch");
        return DFDM_EINVAL;
    }
    int device = 0;
    DFDM_CUDA_OK(cudaGetDevice(&device));
    const int64_t n = plan->n, E = plan->E, nx = plan->nx;
    int solver = 0;
    int rc = resolve_solver(*plan, options, solver);
    if (rc) return rc;
    int cb[Plan::kHostChunksMax + 1];
    const int nchunks = host_chunks(B, solver == DFDM_SOLVER_PCG, 8 * E, cb);
    int widest = 0;
    for (int c = 0; c < nchunks; ++c) widest = std::max(widest, cb[c + 1] - cb[c]);
    int64_t ws_bytes = dfdm_workspace_bytes(plan, widest, options);
    if (ws_bytes < 0) return (int)ws_bytes;
    double *d_q, *d_loads, *d_xfix, *d_loss, *d_dq, *d_xyz, *d_res, *d_vec, *d_len, *d_for;
    int32_t* d_status;
    char* d_ws;
    auto layout = [&](Bump& bump) {
        d_q = bump.take<double>(B * E);
        d_loads = bump.take<double>(n * 3);
        d_xfix = bump.take<double>(std::max<int64_t>(nx * 3, 1));
        d_loss = loss ? bump.take<double>(B) : nullptr;
        d_dq = dq ? bump.take<double>(B * E) : nullptr;
        d_xyz = xyz ? bump.take<double>(B * n * 3) : nullptr;
        d_res = residuals ? bump.take<double>(B * n * 3) : nullptr;
        d_vec = vectors ? bump.take<double>(B * E * 3) : nullptr;
        d_len = lengths ? bump.take<double>(B * E) : nullptr;
        d_for = forces ? bump.take<double>(B * E) : nullptr;
        d_status = bump.take<int32_t>(B);
        d_ws = bump.take<char>(ws_bytes);
    };
    Bump count(nullptr);
    layout(count);
    // an arena and its streams belong to this call until it returns.  The first free slot: a caller on its own always
    // gets slot 0 (one arena's worth of memory); concurrent callers spread over the slots and queue on slot 0 beyond.
    Plan::ArenaSet& slots = arena_ref(*plan, device);
    Plan::Arena* slot = nullptr;
    std::unique_lock<std::mutex> busy;
    for (int k = 0; k < Plan::kArenaSlots && !slot; ++k) {
        std::unique_lock<std::mutex> attempt(slots.slot[k].busy, std::try_to_lock);
        if (attempt.owns_lock()) { busy = std::move(attempt); slot = &slots.slot[k]; }
    }
    if (!slot) { busy = std::unique_lock<std::mutex>(slots.slot[0].busy); slot = &slots.slot[0]; }
    Plan::Arena& A = *slot;
    rc = arena_prepare(A, count.off);
    if (rc) return rc;
    Bump real(static_cast<char*>(A.ptr));
    layout(real);
    cudaStream_t s_in = static_cast<cudaStream_t>(A.stream[0]), s_cmp = static_cast<cudaStream_t>(A.stream[1]),
                 s_out = static_cast<cudaStream_t>(A.stream[2]);
    // Everything that enqueues work sits in `enqueue`, so that an error on the way (reported through its return code) still
    // reaches the drain below: the arena is released on return, and nothing of this call may be in flight by then.
    std::unique_lock<std::mutex> turn(compute_turn(device), std::defer_lock);
    auto enqueue = [&]() -> int {
    // way in: the design-independent arrays, then q chunk by chunk
    DFDM_CUDA_OK(cudaMemcpyAsync(d_loads, loads, sizeof(double) * n * 3, cudaMemcpyHostToDevice, s_in));
    if (nx > 0) DFDM_CUDA_OK(cudaMemcpyAsync(d_xfix, xyz_fixed, sizeof(double) * nx * 3, cudaMemcpyHostToDevice, s_in));
    for (int c = 0; c < nchunks; ++c) {
        const int64_t b0 = cb[c], nb = cb[c + 1] - cb[c];
        DFDM_CUDA_OK(cudaMemcpyAsync(d_q + b0 * E, q + b0 * E, sizeof(double) * nb * E, cudaMemcpyHostToDevice, s_in));
        DFDM_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(A.ev_in[c]), s_in));
    }
    auto at = [](double* p, int64_t off) { return p ? p + off : nullptr; };
    // the kernels of this call, in its turn on the device; the copies above and the wait below stay outside the turn, so
    // another thread's call copies in while this one computes and computes while this one copies out
    turn.lock();
    int rc = DFDM_OK;
    for (int c = 0; c < nchunks && rc == DFDM_OK; ++c) {
        const int64_t b0 = cb[c], nb = cb[c + 1] - cb[c];
        DFDM_CUDA_OK(cudaStreamWaitEvent(s_cmp, static_cast<cudaEvent_t>(A.ev_in[c]), 0));
        if (prog)
            rc = dfdm_loss_and_grad(plan, prog, (int32_t)nb, d_q + b0 * E, d_loads, d_xfix, d_loss + b0, at(d_dq, b0 * E),
                                    at(d_xyz, b0 * n * 3), at(d_res, b0 * n * 3), at(d_vec, b0 * E * 3), at(d_len, b0 * E),
                                    at(d_for, b0 * E), d_status + b0, d_ws, ws_bytes, options, s_cmp);
        else
            rc = dfdm_forward(plan, (int32_t)nb, d_q + b0 * E, d_loads, d_xfix, at(d_xyz, b0 * n * 3), at(d_res, b0 * n * 3),
                              at(d_vec, b0 * E * 3), at(d_len, b0 * E), at(d_for, b0 * E), d_status + b0, d_ws, ws_bytes, options, s_cmp);
        if (rc) break;
        DFDM_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(A.ev_done[c]), s_cmp));
        // way out of this chunk, under the kernels of the next one
        DFDM_CUDA_OK(cudaStreamWaitEvent(s_out, static_cast<cudaEvent_t>(A.ev_done[c]), 0));
        auto back = [&](double* host, const double* dev, int64_t per_design) -> int {
            if (host) DFDM_CUDA_OK(cudaMemcpyAsync(host + b0 * per_design, dev + b0 * per_design, sizeof(double) * nb * per_design,
                                                   cudaMemcpyDeviceToHost, s_out));
            return DFDM_OK;
        };
        if ((rc = back(loss, d_loss, 1)) || (rc = back(dq, d_dq, E)) || (rc = back(xyz, d_xyz, n * 3)) ||
            (rc = back(residuals, d_res, n * 3)) || (rc = back(vectors, d_vec, E * 3)) || (rc = back(lengths, d_len, E)) ||
            (rc = back(forces, d_for, E)))
            break;
        if (status) DFDM_CUDA_OK(cudaMemcpyAsync(status + b0, d_status + b0, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, s_out));
    }
    return rc;
    };
    rc = enqueue();
    if (turn.owns_lock()) turn.unlock();
    // A PCG call lasts tens of milliseconds: it waits on blocking-sync events, so that this thread sleeps through its copies
    // instead of spinning next to the thread whose call has the kernels and is feeding the device.  A direct-path call is
    // a sub-millisecond kernel: waking up would cost as much as the kernel, so it spins (cudaStreamSynchronize).
    const bool sleepy = solver == DFDM_SOLVER_PCG;
    auto drain = [&](cudaStream_t s, void* ev) -> cudaError_t {
        if (!sleepy) return cudaStreamSynchronize(s);
        cudaError_t e = cudaEventRecord(static_cast<cudaEvent_t>(ev), s);
        return e != cudaSuccess ? e : cudaEventSynchronize(static_cast<cudaEvent_t>(ev));
    };
    cudaError_t e1 = drain(s_in, A.ev_end[0]), e2 = drain(s_cmp, A.ev_end[1]), e3 = drain(s_out, A.ev_end[2]);
    if (rc) return rc;
    DFDM_CUDA_OK(e1);
    DFDM_CUDA_OK(e2);
    DFDM_CUDA_OK(e3);
    return DFDM_OK;
}

int dfdm_forward_host(const dfdm_plan* plan, int32_t batch, const double* q, const double* loads, const double* xyz_fixed,
                      double* xyz, double* residuals, double* vectors, double* lengths, double* forces, int32_t* status,
                      const dfdm_options* options) {
    return host_eval(plan, nullptr, batch, q, loads, xyz_fixed, nullptr, nullptr, xyz, residuals, vectors, lengths, forces, status,
                     options);
}

int dfdm_loss_and_grad_host(const dfdm_plan* plan, const dfdm_goalprog* prog, int32_t batch, const double* q, const double* loads,
                            const double* xyz_fixed, double* loss, double* dq, int32_t* status, const dfdm_options* options) {
    if (!prog || !loss) {
        set_error("dfdm_loss_and_grad_host: NULL goal program or loss buffer");
        return DFDM_EINVAL;
    }
    return host_eval(plan, prog, batch, q, loads, xyz_fixed, loss, dq, nullptr, nullptr, nullptr, nullptr, nullptr, status, options);
}

const char* dfdm_last_direct_kernel(void) { return last_direct_kernel(); }

int64_t dfdm_launch_count(void) { return (int64_t)g_launches.load(); }

void dfdm_last_pcg_iterations(int32_t* forward_iters, int32_t* adjoint_iters) {
    int f = 0, a = 0;
    last_pcg_iterations(&f, &a);
    if (forward_iters) *forward_iters = f;
    if (adjoint_iters) *adjoint_iters = a;
}

void dfdm_profile_enable(int32_t on) { profile_enable(on); }

void dfdm_profile_read(double* ms, int64_t* count) {
    double m[8];
    long long c[8];
    profile_read(m, c);
    for (int k = 0; k < 8; ++k) {
        if (ms) ms[k] = m[k];
        if (count) count[k] = (int64_t)c[k];
    }
}

}  // extern "C"
