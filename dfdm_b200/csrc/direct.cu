// Direct path: one CTA per design, the whole evaluation fused in one kernel.
//
//   assemble K (banded, solver ordering) and the load term in shared memory
//   -> banded LDL^T in place -> 3 right-hand sides -> equilibrium state (edge vectors, lengths,
//   forces, nodal residuals) -> goals / loss -> reverse pass -> one more solve with the same factor
//   -> dL/dq
//
// Reference being replaced: src/dfdm/equilibrium/model.py:36-79 (forward), src/dfdm/losses/losses.py
// and src/dfdm/goals/* (loss), autograd's tape replay behind src/dfdm/optimization.py:61 (gradient).
//
// Shared-memory layout (doubles): D[(w+1)][LD]  banded factor, D[k][i] = K(i, i-k); LD = nf | 1
//                                 rhs[3][nf]    right-hand sides / solutions, by solver position
//                                 lcol[w+1], lsc[w+1]  column scratch of the factorisation
// LDL^T needs no sign handling: K is negative definite for compression (q < 0), positive definite
// for tension, and merely needs non-zero pivots when q has mixed sign.
#include "common.cuh"

namespace dfdm {

struct DirectArgs {
    PlanDev P;
    GoalProgDev G;
    int has_prog, adjoint, B;
    int tx_shift;                    // threads along a band row in the trailing update = 1 << tx_shift
    int k_shift;                     // 1 << k_shift >= bandwidth: band offsets per right-hand side in the solves
    const double* q;
    const double* loads;
    const double* xfix;
    double *X, *R, *U, *L, *F;       // [B][..] design-major state (caller's outputs or workspace)
    double *Rb, *Xb, *Ub;            // cotangent scratch
    const double *c_xyz, *c_res, *c_vec, *c_len, *c_for;
    double* loss;
    double* dq;
    int32_t* status;
};

__device__ __forceinline__ void cta_sync(int nt) {
    if (nt <= 32) __syncwarp(); else __syncthreads();
}

__device__ double cta_sum(double v, double* red, int nt) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (nt <= 32) return __shfl_sync(0xffffffffu, v, 0);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < (nt >> 5); ++k) s += red[k];   // fixed order: deterministic
    return s;
}

// Solve (L D L^T) x = rhs for the 3 columns in shared memory.  A thread owns (column c, band offset k)
// with k in the low kshift bits of its index, so the hot loops carry no integer division.
__device__ void band_solve(const double* __restrict__ D, double* __restrict__ rhs, int nf, int w, int LD, int nt, int kshift) {
    const int tid = threadIdx.x;
    if (w == 1) {
        // chains (bandwidth 1): a sweep is one dependent multiply-add per row.  Thread c walks right-hand side c on its
        // own - no barrier per column - with the next row's operands loaded ahead of the dependent chain.
        if (tid < 3) {
            double* r = rhs + (size_t)tid * nf;
            const double* l = D + LD;
            double y = r[0];
            double ln = nf > 1 ? l[1] : 0.0, rn = nf > 1 ? r[1] : 0.0;
            for (int j = 1; j < nf; ++j) {
                const double lj = ln, rj = rn;
                if (j + 1 < nf) { ln = l[j + 1]; rn = r[j + 1]; }
                y = rj - lj * y;
                r[j] = y;
            }
        }
        cta_sync(nt);
        for (int i = tid; i < nf; i += nt) {
            double d = D[i];
            rhs[i] /= d;
            rhs[nf + i] /= d;
            rhs[2 * nf + i] /= d;
        }
        cta_sync(nt);
        if (tid < 3) {
            double* r = rhs + (size_t)tid * nf;
            const double* l = D + LD;
            double x = r[nf - 1];
            double ln = nf > 1 ? l[nf - 1] : 0.0, rn = nf > 1 ? r[nf - 2] : 0.0;
            for (int j = nf - 1; j > 0; --j) {
                const double lj = ln, rj = rn;
                if (j > 1) { ln = l[j - 1]; rn = r[j - 2]; }
                x = rj - lj * x;
                r[j - 1] = x;
            }
        }
        cta_sync(nt);
        return;
    }
    const int span = 3 << kshift, kmask = (1 << kshift) - 1;
    for (int j = 0; j < nf - 1; ++j) {                       // forward: L y = b
        int m = min(w, nf - 1 - j);
        for (int t = tid; t < span; t += nt) {
            int c = t >> kshift, k = (t & kmask) + 1;
            if (k <= m) rhs[c * nf + j + k] -= D[k * LD + j + k] * rhs[c * nf + j];
        }
        cta_sync(nt);
    }
    for (int i = tid; i < nf; i += nt) {                     // D z = y
        double d = D[i];
        rhs[i] /= d;
        rhs[nf + i] /= d;
        rhs[2 * nf + i] /= d;
    }
    cta_sync(nt);
    for (int j = nf - 1; j > 0; --j) {                       // backward: L^T x = z
        int m = min(w, j);
        for (int t = tid; t < span; t += nt) {
            int c = t >> kshift, k = (t & kmask) + 1;
            if (k <= m) rhs[c * nf + j - k] -= D[k * LD + j] * rhs[c * nf + j];
        }
        cta_sync(nt);
    }
}

__global__ void __launch_bounds__(256) k_direct_eval(DirectArgs a) {
    extern __shared__ double smem[];
    const PlanDev& P = a.P;
    const int nf = P.nf, n = P.n, E = P.E, w = P.w;
    const int LD = nf | 1;
    const int nt = blockDim.x, tid = threadIdx.x;
    double* D = smem;
    double* rhs = D + (size_t)(w + 1) * LD;
    double* lcol = rhs + 3 * (size_t)nf;
    double* lsc = lcol + (w + 1);
    double* red = lsc + (w + 1);                 // 16 doubles
    __shared__ int s_bad;

    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
        const double* q = a.q + (size_t)b * E;
        double* X = a.X + (size_t)b * n * 3;
        double* U = a.U + (size_t)b * E * 3;
        double* Lg = a.L + (size_t)b * E;
        double* F = a.F + (size_t)b * E;
        double* R = a.R + (size_t)b * n * 3;
        if (tid == 0) s_bad = 0;

        // ---- assembly: gather form of the scatter-add, one segment (row) per thread ----------------
        for (int idx = tid; idx < (w + 1) * LD; idx += nt) D[idx] = 0.0;
        cta_sync(nt);
        for (int p = tid; p < nf; p += nt) {
            int node = P.free_nodes[P.perm[p]];
            double diag = 0.0;
            double r0 = a.loads[node * 3 + 0], r1 = a.loads[node * 3 + 1], r2 = a.loads[node * 3 + 2];
            for (int k = P.sinc_ptr[p]; k < P.sinc_ptr[p + 1]; ++k) {
                double qe = q[P.sinc_edge[k]];
                int code = P.sinc_code[k];
                diag += qe;
                if (code > 0) {
                    D[code * LD + p] -= qe;                       // K(p, p - code) = -q
                } else if (code < 0) {
                    const double* xk = a.xfix + (size_t)(-1 - code) * 3;
                    r0 += qe * xk[0];                             // b = P_f - C_f^T Q C_x X_x
                    r1 += qe * xk[1];
                    r2 += qe * xk[2];
                }
            }
            D[p] = diag;
            rhs[p] = r0;
            rhs[nf + p] = r1;
            rhs[2 * nf + p] = r2;
        }
        cta_sync(nt);

        // ---- banded LDL^T, right-looking, in place ---------------------------------------------------
        if (w == 1) {
            // chains: d_(j+1) = K(j+1,j+1) - l^2 d_j with l = K(j+1,j) / d_j is one division and one multiply-add per row,
            // strictly sequential; one thread runs it without the two barriers per column of the general path
            if (tid == 0) {
                const double* diag = D;
                double* sub = D + LD;
                double d = diag[0];
                int bad = (d == 0.0 || !isfinite(d)) ? 1 : 0;
                double en = nf > 1 ? sub[1] : 0.0, bn = nf > 1 ? diag[1] : 0.0;
                for (int j = 0; j < nf - 1; ++j) {
                    const double e = en, b2 = bn;
                    if (j + 2 < nf) { en = sub[j + 2]; bn = diag[j + 2]; }
                    const double l = e / d;
                    d = b2 - l * e;
                    sub[j + 1] = l;
                    D[j + 1] = d;
                    if (d == 0.0 || !isfinite(d)) bad = 1;
                }
                if (bad) s_bad = 1;
            }
            cta_sync(nt);
        } else {
            const int TX = 1 << a.tx_shift, tx = tid & (TX - 1), ty = tid >> a.tx_shift, TY = nt >> a.tx_shift;
            for (int j = 0; j < nf - 1; ++j) {
                int m = min(w, nf - 1 - j);
                double dj = D[j];
                if (tid == 0 && (dj == 0.0 || !isfinite(dj))) s_bad = 1;
                for (int k = 1 + tid; k <= m; k += nt) {
                    double v = D[k * LD + j + k];
                    lcol[k] = v;                 // l_k d_j
                    v /= dj;
                    lsc[k] = v;                  // l_k
                    D[k * LD + j + k] = v;
                }
                cta_sync(nt);
                // K(j+a, j+a-kk) -= l_a d_j l_(a-kk), kk = 0..m-1, a = kk+1..m  (rows contiguous in a)
                for (int kk = ty; kk < m; kk += TY)
                    for (int r = kk + 1 + tx; r <= m; r += TX) D[kk * LD + j + r] -= lsc[r] * lcol[r - kk];
                cta_sync(nt);
            }
            if (tid == 0) {
                double dl = D[nf - 1];
                if (dl == 0.0 || !isfinite(dl)) s_bad = 1;
            }
        }

        // ---- forward solve and equilibrium state ------------------------------------------------------
        band_solve(D, rhs, nf, w, LD, nt, a.k_shift);
        for (int i = tid; i < n; i += nt) {
            int src = P.node_src[i];
            if (src >= 0) {
                int p = P.pos[src];
                X[i * 3 + 0] = rhs[p];
                X[i * 3 + 1] = rhs[nf + p];
                X[i * 3 + 2] = rhs[2 * nf + p];
            } else {
                const double* xk = a.xfix + (size_t)(-1 - src) * 3;
                X[i * 3 + 0] = xk[0];
                X[i * 3 + 1] = xk[1];
                X[i * 3 + 2] = xk[2];
            }
        }
        cta_sync(nt);
        double lp_part = 0.0;
        for (int e = tid; e < E; e += nt) {
            int u = P.edges[2 * e], v = P.edges[2 * e + 1];
            double u0 = X[v * 3 + 0] - X[u * 3 + 0], u1 = X[v * 3 + 1] - X[u * 3 + 1], u2 = X[v * 3 + 2] - X[u * 3 + 2];
            double len = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
            double f = q[e] * len;
            U[e * 3 + 0] = u0;
            U[e * 3 + 1] = u1;
            U[e * 3 + 2] = u2;
            Lg[e] = len;
            F[e] = f;
            lp_part += fabs(len * f);
        }
        cta_sync(nt);

        double loss_part = 0.0;
        const bool seeds = a.adjoint;
        double* Rb = seeds ? a.Rb + (size_t)b * n * 3 : nullptr;
        double* Xb = seeds ? a.Xb + (size_t)b * n * 3 : nullptr;
        for (int i = tid; i < n; i += nt) {
            double r[3] = {a.loads[i * 3 + 0], a.loads[i * 3 + 1], a.loads[i * 3 + 2]};
            for (int k = P.ninc_ptr[i]; k < P.ninc_ptr[i + 1]; ++k) {
                int e = P.ninc_edge[k];
                double s = (double)P.ninc_sign[k] * q[e];
                r[0] -= s * U[e * 3 + 0];
                r[1] -= s * U[e * 3 + 1];
                r[2] -= s * U[e * 3 + 2];
            }
            R[i * 3 + 0] = r[0];
            R[i * 3 + 1] = r[1];
            R[i * 3 + 2] = r[2];
            if (a.has_prog || seeds) {
                double xb[3] = {0.0, 0.0, 0.0}, rb[3] = {0.0, 0.0, 0.0};
                if (a.has_prog) {
                    double x[3] = {X[i * 3 + 0], X[i * 3 + 1], X[i * 3 + 2]};
                    loss_part += node_goals(a.G.recs, a.G.ngoal_ptr[i], a.G.ngoal_ptr[i + 1], x, r, xb, rb);
                }
                if (seeds) {
                    if (a.c_xyz) for (int c = 0; c < 3; ++c) xb[c] += a.c_xyz[((size_t)b * n + i) * 3 + c];
                    if (a.c_res) for (int c = 0; c < 3; ++c) rb[c] += a.c_res[((size_t)b * n + i) * 3 + c];
                    for (int c = 0; c < 3; ++c) { Xb[i * 3 + c] = xb[c]; Rb[i * 3 + c] = rb[c]; }
                }
            }
        }

        // network goals need the total load path before the edge cotangents
        double lpbar = 0.0;
        if (a.has_prog && a.G.n_net > 0) {
            double lp = cta_sum(lp_part, red, nt);
            double net_loss = network_goals(a.G.recs, a.G.net_begin, a.G.net_begin + a.G.n_net, lp, lpbar);
            if (tid == 0) loss_part += net_loss;
        }
        cta_sync(nt);

        // ---- reverse pass ------------------------------------------------------------------------------
        double* Ub = seeds ? a.Ub + (size_t)b * E * 3 : nullptr;
        double* dq = seeds ? a.dq + (size_t)b * E : nullptr;
        if (a.has_prog || seeds) {
            for (int e = tid; e < E; e += nt) {
                double u[3] = {U[e * 3 + 0], U[e * 3 + 1], U[e * 3 + 2]};
                double len = Lg[e], f = F[e], qe = q[e];
                double ub[3] = {0.0, 0.0, 0.0}, lb = 0.0, fb = 0.0;
                if (a.has_prog) {
                    loss_part += edge_goals(a.G.recs, a.G.egoal_ptr[e], a.G.egoal_ptr[e + 1], u, len, f, ub, lb, fb);
                    loss_part += a.G.l2_alpha * qe * qe;
                }
                if (seeds) {
                    if (a.c_vec) for (int c = 0; c < 3; ++c) ub[c] += a.c_vec[((size_t)b * E + e) * 3 + c];
                    if (a.c_len) lb += a.c_len[(size_t)b * E + e];
                    if (a.c_for) fb += a.c_for[(size_t)b * E + e];
                    int un = P.edges[2 * e], vn = P.edges[2 * e + 1];
                    double g[3] = {Rb[vn * 3 + 0] - Rb[un * 3 + 0], Rb[vn * 3 + 1] - Rb[un * 3 + 1], Rb[vn * 3 + 2] - Rb[un * 3 + 2]};
                    double d = edge_adjoint(qe, u, len, f, g, lpbar, ub, lb, fb);
                    Ub[e * 3 + 0] = ub[0];
                    Ub[e * 3 + 1] = ub[1];
                    Ub[e * 3 + 2] = ub[2];
                    dq[e] = d + 2.0 * a.G.l2_alpha * qe;
                }
            }
        }
        if (a.loss) {
            double total = cta_sum(loss_part, red, nt);
            if (tid == 0) a.loss[b] = total;
        }
        if (seeds) {
            cta_sync(nt);
            // Xbar = direct + C^T Ubar; right-hand side of the adjoint solve on the free rows
            for (int i = tid; i < n; i += nt) {
                int src = P.node_src[i];
                if (src < 0) continue;
                double xb[3] = {Xb[i * 3 + 0], Xb[i * 3 + 1], Xb[i * 3 + 2]};
                for (int k = P.ninc_ptr[i]; k < P.ninc_ptr[i + 1]; ++k) {
                    int e = P.ninc_edge[k];
                    double s = (double)P.ninc_sign[k];
                    xb[0] += s * Ub[e * 3 + 0];
                    xb[1] += s * Ub[e * 3 + 1];
                    xb[2] += s * Ub[e * 3 + 2];
                }
                int p = P.pos[src];
                rhs[p] = xb[0];
                rhs[nf + p] = xb[1];
                rhs[2 * nf + p] = xb[2];
            }
            cta_sync(nt);
            band_solve(D, rhs, nf, w, LD, nt, a.k_shift);       // K symmetric: same factor
            for (int e = tid; e < E; e += nt) {
                int su = P.node_src[P.edges[2 * e]], sv = P.node_src[P.edges[2 * e + 1]];
                double d = 0.0;
                for (int c = 0; c < 3; ++c) {
                    double lv = sv >= 0 ? rhs[c * nf + P.pos[sv]] : 0.0;
                    double lu = su >= 0 ? rhs[c * nf + P.pos[su]] : 0.0;
                    d += (lv - lu) * U[e * 3 + c];
                }
                dq[e] -= d;                               // dK/dq_e = c_e c_e^T  =>  -(C_f lambda)_e . U_e
            }
        }
        cta_sync(nt);
        if (tid == 0 && a.status) a.status[b] = s_bad ? DFDM_STATUS_SINGULAR : DFDM_STATUS_OK;
        cta_sync(nt);
    }
}

int64_t direct_workspace(const Plan& P, int B, bool adjoint, bool has_prog) {
    (void)has_prog;
    Bump bump(nullptr);
    int64_t n = P.n, E = P.E;
    bump.take<double>(B * n * 3);   // X
    bump.take<double>(B * n * 3);   // R
    bump.take<double>(B * E * 3);   // U
    bump.take<double>(B * E);       // L
    bump.take<double>(B * E);       // F
    if (adjoint) {
        bump.take<double>(B * n * 3);   // Rb
        bump.take<double>(B * n * 3);   // Xb
        bump.take<double>(B * E * 3);   // Ub
    }
    return bump.off;
}

int run_direct(Eval& ev) {
    const PlanDev& P = ev.P;
    const int64_t n = P.n, E = P.E, B = ev.B;
    Bump bump(ev.ws);
    DirectArgs a{};
    a.P = P;
    a.G = ev.G;
    a.has_prog = ev.has_prog ? 1 : 0;
    a.adjoint = ev.adjoint ? 1 : 0;
    a.B = ev.B;
    a.q = ev.q;
    a.loads = ev.loads;
    a.xfix = ev.xfix;
    double* wX = bump.take<double>(B * n * 3);
    double* wR = bump.take<double>(B * n * 3);
    double* wU = bump.take<double>(B * E * 3);
    double* wL = bump.take<double>(B * E);
    double* wF = bump.take<double>(B * E);
    a.X = ev.o_xyz ? ev.o_xyz : wX;
    a.R = ev.o_res ? ev.o_res : wR;
    a.U = ev.o_vec ? ev.o_vec : wU;
    a.L = ev.o_len ? ev.o_len : wL;
    a.F = ev.o_for ? ev.o_for : wF;
    if (ev.adjoint) {
        a.Rb = bump.take<double>(B * n * 3);
        a.Xb = bump.take<double>(B * n * 3);
        a.Ub = bump.take<double>(B * E * 3);
    }
    if (bump.off > ev.ws_bytes) {
        set_error("workspace too small for the direct path: need " + std::to_string(bump.off) + " bytes");
        return DFDM_ENOMEM;
    }
    a.c_xyz = ev.c_xyz;
    a.c_res = ev.c_res;
    a.c_vec = ev.c_vec;
    a.c_len = ev.c_len;
    a.c_for = ev.c_for;
    a.loss = ev.loss;
    a.dq = ev.dq;
    a.status = ev.status;

    const int w = P.w;
    int nt = w <= 6 ? 32 : (w <= 12 ? 64 : (w <= 48 ? 128 : 256));
    if (w == 1 && P.nf > 256) nt = 256;      // chains: the sweeps are single-threaded anyway, the node / edge passes want the threads
    int tx = 4;
    while (tx < 32 && tx < w) tx <<= 1;
    if (tx > nt) tx = nt;
    int shift = 0;
    while ((1 << shift) < tx) ++shift;
    a.tx_shift = shift;
    int kshift = 0;
    while ((1 << kshift) < w) ++kshift;
    a.k_shift = kshift;
    int64_t smem = direct_smem_bytes(P.nf, w);
    int dev = 0, sms = 0, per_sm = 0;
    DFDM_CUDA_OK(cudaGetDevice(&dev));
    if (smem > 48 * 1024) {                               // opt-in is per device and per function: set it wherever we launch
        static std::mutex mu;
        static std::map<int, int64_t> configured;
        std::lock_guard<std::mutex> lock(mu);
        if (configured[dev] < smem) {
            DFDM_CUDA_OK(cudaFuncSetAttribute(k_direct_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev] = smem;
        }
    }
    DFDM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DFDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_direct_eval, nt, (size_t)smem));
    if (per_sm < 1) per_sm = 1;
    int grid = (int)std::min<int64_t>(B, (int64_t)sms * per_sm);
    k_direct_eval<<<grid, nt, (size_t)smem, ev.stream>>>(a);
    DFDM_LAUNCH_OK("k_direct_eval");
    return DFDM_OK;
}

}  // namespace dfdm
