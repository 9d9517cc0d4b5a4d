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
//
// Two launch shapes share the per-design body (eval_design):
//   k_direct_eval  one CTA per design, column-by-column LDL^T with block barriers: wide bands (w > 28), long chains
//                  (w = 1) and designs whose factor takes most of a CTA's shared memory
//   k_direct_warp  one WARP per design, several designs per CTA, for 2 <= w <= 28: nothing but __syncwarp between the
//                  stages; the factorisation is BLOCKED by four columns - the panel of a block is eliminated in
//                  registers (a lane per row, pivots and multipliers by shuffle) and the trailing window takes the
//                  block's rank-4 update as FP64 tensor-core products (mma.sync.m8n8k4.f64, operands read straight from
//                  the band) - and the substitutions keep the three right-hand sides in registers, a 64-row sliding
//                  window over the lanes, so a column costs a shuffle and a multiply-add instead of a shared-memory trip.
// After the forward solve both check the residual of the solve itself (R on the free nodes, which is computed anyway)
// against the terms it is a difference of: an unpivoted factorisation of an indefinite K can lose accuracy through a tiny
// pivot without ever hitting an exact zero, and must not report DFDM_STATUS_OK then.
#include <cstdlib>

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
    double *X, *R, *U, *L, *F;       // [B][..] design-major state in global memory (workspace; unused when the state is in shared memory)
    double *Rb, *Xb, *Ub;            // cotangent scratch, likewise
    double *oX, *oR, *oU, *oL, *oF;  // the caller's outputs (nullable)
    int chain_off;                   // >= 0 (w = 1 only): scratch of the partition method at this offset (doubles); -1: the serial sweeps
    int state_off;                   // >= 0: the per-design state lives in shared memory, at this offset (doubles) of the group's block:
                                     // q[E], X[3n] (later Xbar), U[3E], L[E], F[E], Rbar[3n], Ubar[3E] - every pass then runs on shared
                                     // memory and only q comes in and loss / dq (and requested outputs) go out
    double* band;                    // non-NULL: the band and the right-hand sides of CTA k live in GLOBAL memory at band + k * band_stride
    int64_t band_stride;             // (doubles) - factors too large for shared memory (k_direct_global)
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

constexpr double kDirectResidualTol = 1e-9;     // largest |R| on a free node, relative to the largest term of any nodal balance

__device__ double cta_max(double v, double* red, int nt) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    if (nt <= 32) return __shfl_sync(0xffffffffu, v, 0);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < (nt >> 5); ++k) s = fmax(s, red[k]);
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

// ---- chains (w = 1): partition method ------------------------------------------------------------------------------------
// K is tridiagonal.  The classic recurrences (d_i = a_i - b_i^2 / d_(i-1), one substitution step per row) are n_f
// DEPENDENT steps - 25 000 for the factorisation and the two solves of a 5 000-segment arch.  Instead the chain is cut
// into C <= 256 parts of m rows (m odd: conflict-free strides); the last row of every part but the last is an INTERFACE.
// Every part's interior is an independent tridiagonal block T_c that ONE thread factors and solves on its own (m
// dependent steps, all parts at once); the interfaces see each other through the Schur complement
//     S = A_II - B T^-1 B^T,   tridiagonal of order C - 1,
// whose entries need only (T_c^-1)_first,first = sum_i y_i^2 / d_i, (T_c^-1)_last,last = 1 / d_last and
// (T_c^-1)_first,last = y_last / d_last (y: the unit-lower recurrence started at e_first), and which is solved by parallel
// cyclic reduction, one thread per interface, log2(C) steps.  A solve is: one store-free sweep per part for the two end
// values of T_c^-1 r_c, the reduced system, then the interior solves with the interface values moved to the right-hand
// side.  Storage: the band in place (1 / d on the diagonal of interior rows, multipliers on the sub-diagonal; interface
// rows keep a and b) plus 11 arrays of 256 doubles.
constexpr int kChainParts = 256;
constexpr int kChainScratch = 11 * kChainParts;          // doubles

struct ChainGeom { int m, C; };
__host__ __device__ inline ChainGeom chain_geom(int nf) {
    ChainGeom g;
    g.m = max(3, ((nf + kChainParts - 1) / kChainParts) | 1);
    g.C = (nf + g.m - 1) / g.m;
    return g;
}

// returns non-zero (in every thread of the CTA) on a zero / non-finite pivot
__device__ int chain_factor(double* __restrict__ D, int nf, int LD, double* __restrict__ scr, int* s_flag) {
    const ChainGeom g = chain_geom(nf);
    const int c = threadIdx.x;
    double *alpha = scr, *beta = scr + kChainParts, *gamma = scr + 2 * kChainParts, *sd = scr + 3 * kChainParts, *so = scr + 4 * kChainParts;
    int bad = 0;
    if (c < g.C) {
        const int s = c * g.m, e = min(nf, s + g.m), t = c < g.C - 1 ? e - 1 : e;
        double d = D[s];
        if (d == 0.0 || !isfinite(d)) bad = 1;
        double rd = 1.0 / d, y = 1.0, al = rd;
        D[s] = rd;
        for (int i = s + 1; i < t; ++i) {
            const double b = D[LD + i];
            const double l = b * rd;
            d = D[i] - l * b;
            if (d == 0.0 || !isfinite(d)) bad = 1;
            rd = 1.0 / d;
            D[LD + i] = l;
            D[i] = rd;
            y = -l * y;
            al += y * y * rd;
        }
        alpha[c] = al;
        beta[c] = rd;
        gamma[c] = y * rd;
    }
    if (bad) *s_flag = 1;
    __syncthreads();
    if (c < g.C - 1) {
        const int I = min(nf, (c + 1) * g.m) - 1;
        const double bl = D[LD + I], br = D[LD + I + 1];
        sd[c] = D[I] - bl * bl * beta[c] - br * br * alpha[c + 1];
        double off = 0.0;
        if (c + 1 < g.C - 1) {
            const int I2 = min(nf, (c + 2) * g.m) - 1;
            off = -br * gamma[c + 1] * D[LD + I2];
        }
        so[c] = off;
    }
    __syncthreads();
    return *s_flag;
}

__device__ void chain_solve(const double* __restrict__ D, double* __restrict__ rhs, int nf, int LD, double* __restrict__ scr) {
    const ChainGeom g = chain_geom(nf);
    const int c = threadIdx.x, N = g.C - 1;
    double *yf = scr, *sd = scr + 3 * kChainParts, *so = scr + 4 * kChainParts;          // yf[3][parts] takes the place of alpha / beta / gamma
    double *pa = scr + 5 * kChainParts, *pd = scr + 6 * kChainParts, *pc = scr + 7 * kChainParts, *pr = scr + 8 * kChainParts;
    const int s = c * g.m, e = min(nf, s + g.m), t = c < g.C - 1 ? e - 1 : e;
    double yl[3] = {0.0, 0.0, 0.0};
    if (c < g.C) {
        // end values of T^-1 r for this part, no stores: x_last = z_last, x_first = sum_i y_i z_i
        double v[3], xf[3];
        const double rd0 = D[s];
#pragma unroll
        for (int k = 0; k < 3; ++k) { v[k] = rhs[k * nf + s]; xf[k] = v[k] * rd0; }
        double y = 1.0, rd = rd0;
        for (int i = s + 1; i < t; ++i) {
            const double l = D[LD + i];
            rd = D[i];
            y = -l * y;
            const double yr = y * rd;
#pragma unroll
            for (int k = 0; k < 3; ++k) { v[k] = rhs[k * nf + i] - l * v[k]; xf[k] += yr * v[k]; }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { yl[k] = v[k] * rd; yf[k * kChainParts + c] = xf[k]; }
    }
    __syncthreads();
    // reduced system on the interfaces: parallel cyclic reduction, own coefficients in registers
    double a = 0.0, d = 1.0, cc = 0.0, r[3] = {0.0, 0.0, 0.0};
    const int I = min(nf, (c + 1) * g.m) - 1;
    if (c < N) {
        const double bl = D[LD + I], br = D[LD + I + 1];
        a = c > 0 ? so[c - 1] : 0.0;
        d = sd[c];
        cc = so[c];
#pragma unroll
        for (int k = 0; k < 3; ++k) r[k] = rhs[k * nf + I] - bl * yl[k] - br * yf[k * kChainParts + c + 1];
    }
    for (int h = 1; h < N; h <<= 1) {
        __syncthreads();
        if (c < N) { pa[c] = a; pd[c] = d; pc[c] = cc; pr[c] = r[0]; pr[kChainParts + c] = r[1]; pr[2 * kChainParts + c] = r[2]; }
        __syncthreads();
        if (c < N) {
            const int lo = c - h, hi = c + h;
            double k1 = 0.0, k2 = 0.0, na = 0.0, nc = 0.0;
            if (lo >= 0) {
                k1 = a / pd[lo];
                na = -k1 * pa[lo];
                d -= k1 * pc[lo];
#pragma unroll
                for (int k = 0; k < 3; ++k) r[k] -= k1 * pr[k * kChainParts + lo];
            }
            if (hi < N) {
                k2 = cc / pd[hi];
                nc = -k2 * pc[hi];
                d -= k2 * pa[hi];
#pragma unroll
                for (int k = 0; k < 3; ++k) r[k] -= k2 * pr[k * kChainParts + hi];
            }
            a = na;
            cc = nc;
        }
    }
    if (c < N) {
        const double rd = 1.0 / d;
#pragma unroll
        for (int k = 0; k < 3; ++k) rhs[k * nf + I] = r[k] * rd;
    }
    __syncthreads();
    // interiors, with the interface values on the right-hand side
    if (c < g.C) {
        const double bs = c > 0 ? D[LD + s] : 0.0, bt = c < g.C - 1 ? D[LD + t] : 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double* x = rhs + (size_t)k * nf;
            const double xl = c > 0 ? x[s - 1] : 0.0, xr = c < g.C - 1 ? x[t] : 0.0;
            x[s] -= bs * xl;
            x[t - 1] -= bt * xr;
            double v = x[s];
            x[s] = v * D[s];
            for (int i = s + 1; i < t; ++i) {
                v = x[i] - D[LD + i] * v;
                x[i] = v * D[i];                              // z = y / d
            }
            double xn = x[t - 1];
            for (int i = t - 2; i >= s; --i) {
                xn = x[i] - D[LD + i + 1] * xn;
                x[i] = xn;
            }
        }
    }
    __syncthreads();
}

// ---- warp-per-design building blocks (2 <= w <= kWarpMaxW) ------------------------------------------------------------

constexpr int kWarpMaxW = 28;            // a block's panel has w + 4 rows: one lane each
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Banded LDL^T of one design by one warp, blocked by four columns.  D[k * LD + i] = K(i, i - k) on entry; on exit the
// diagonal holds d and the sub-diagonals the unit-lower multipliers.  Returns non-zero (in every lane) on a zero or
// non-finite pivot.
__device__ int band_ldlt_warp(double* __restrict__ D, int nf, int w, int LD) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;                    // mma fragment coordinates
    int bad = 0;
    for (int j0 = 0; j0 < nf; j0 += 4) {
        const int nb = min(4, nf - j0);
        const int row = j0 + lane;
        // ---- the panel: rows j0 .. j0 + w + 3, columns j0 .. j0 + nb - 1, a lane per row, in registers ----------------
        double p[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool in = c < nb && c <= lane && lane - c <= w && row < nf;
            p[c] = in ? D[(lane - c) * LD + row] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c >= nb) break;
            const double piv = __shfl_sync(kFull, p[c], c);   // lane c holds the diagonal entry of column j0 + c
            if (piv == 0.0 || !isfinite(piv)) bad = 1;
            const double rinv = 1.0 / piv;
            const double v = lane > c ? p[c] : 0.0;           // l d of this row (zero above the diagonal and outside the band)
            const double l = v * rinv;
#pragma unroll
            for (int c2 = c + 1; c2 < 4; ++c2) {
                const double lc2 = __shfl_sync(kFull, l, c2); // multiplier of row j0 + c2 in column j0 + c
                if (c2 < nb && lane >= c2) p[c2] -= v * lc2;
            }
            if (lane > c) p[c] = l;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const bool in = c < nb && c <= lane && lane - c <= w && row < nf;
            if (in) D[(lane - c) * LD + row] = p[c];
        }
        __syncwarp();
        // ---- the trailing window (rows and columns j0 + 4 .. j0 + 3 + w) takes the rank-4 update on the tensor cores ----
        const int W0 = j0 + 4;
        const int count = min(w, nf - W0);
        if (nb == 4 && count > 0) {
            const double dk = D[j0 + t];                      // pivot of this lane's k index
            const int tiles = (count + 7) >> 3;
            for (int ti = 0; ti < tiles; ++ti) {
                const int ra = W0 + 8 * ti + g;               // A fragment: row ra, column j0 + t, scaled by -d
                const int ka = ra - (j0 + t);
                const double af = (ra < nf && ka <= w) ? -D[ka * LD + ra] * dk : 0.0;
                for (int tj = 0; tj <= ti; ++tj) {
                    const int cb = W0 + 8 * tj + g;           // B fragment: k = t, column cb of the window = row cb of the panel
                    const int kb = cb - (j0 + t);
                    const double bf = (cb < nf && kb <= w) ? D[kb * LD + cb] : 0.0;
                    const int rc = W0 + 8 * ti + g, cc = W0 + 8 * tj + 2 * t;
                    const bool in0 = rc < nf && cc <= rc && rc - cc <= w, in1 = rc < nf && cc + 1 <= rc && rc - cc - 1 <= w;
                    double c0 = in0 ? D[(rc - cc) * LD + rc] : 0.0, c1 = in1 ? D[(rc - cc - 1) * LD + rc] : 0.0;
                    dmma_884(c0, c1, af, bf);
                    if (in0) D[(rc - cc) * LD + rc] = c0;
                    if (in1) D[(rc - cc - 1) * LD + rc] = c1;
                }
            }
        }
        __syncwarp();
    }
    return __any_sync(kFull, bad);
}

// (L D L^T) x = rhs for the three columns, one warp: the rows travel through registers in two 32-row slots (the current
// one and its neighbour), a lane per row; a column costs one shuffle per right-hand side and a multiply-add - no branch
// (every lane updates exactly one row per column, in one slot or the other: the multiplier is routed by a select) and
// the multiplier of the next column is in flight while this column's chain (shuffle -> multiply-add) completes.
__device__ void band_solve_warp(const double* __restrict__ D, double* __restrict__ rhs, int nf, int w, int LD) {
    const int lane = threadIdx.x & 31;
    const int nslots = (nf + 31) >> 5;
    double cur[3], nxt[3];
    // ---- forward: L y = b, then z = y / d on the way out of the registers -----------------------------------------
#pragma unroll
    for (int c = 0; c < 3; ++c) cur[c] = lane < nf ? rhs[c * nf + lane] : 0.0;
    for (int s = 0; s < nslots; ++s) {
        const int r0 = 32 * s + lane, r1 = r0 + 32;
#pragma unroll
        for (int c = 0; c < 3; ++c) nxt[c] = r1 < nf ? rhs[c * nf + r1] : 0.0;
        const int jmax = min(32, nf - 32 * s);
        // multiplier of this lane's row for column lj: l(row, 32 s + lj); row in the current slot when lane > lj
        auto mult = [&](int lj) {
            const bool own = lane > lj;
            const int k = own ? lane - lj : lane - lj + 32, row = own ? r0 : r1;
            return (lj < jmax && k <= w && row < nf) ? D[k * LD + row] : 0.0;
        };
        double l = mult(0);
        for (int lj = 0; lj < jmax; ++lj) {
            const double ln = mult(lj + 1);                     // next column's multiplier: independent of the chain
            const double y0 = __shfl_sync(kFull, cur[0], lj), y1 = __shfl_sync(kFull, cur[1], lj), y2 = __shfl_sync(kFull, cur[2], lj);
            const double lo = lane > lj ? l : 0.0, lx = lane > lj ? 0.0 : l;
            cur[0] -= lo * y0; cur[1] -= lo * y1; cur[2] -= lo * y2;
            nxt[0] -= lx * y0; nxt[1] -= lx * y1; nxt[2] -= lx * y2;
            l = ln;
        }
        if (r0 < nf) {
            const double dinv = 1.0 / D[r0];
#pragma unroll
            for (int c = 0; c < 3; ++c) rhs[c * nf + r0] = cur[c] * dinv;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) cur[c] = nxt[c];
    }
    __syncwarp();
    // ---- backward: L^T x = z, slots from the last to the first; `nxt` now holds the slot BELOW the current one ----------
    {
        const int rl = 32 * (nslots - 1) + lane;
#pragma unroll
        for (int c = 0; c < 3; ++c) cur[c] = rl < nf ? rhs[c * nf + rl] : 0.0;
    }
    for (int s = nslots - 1; s >= 0; --s) {
        const int r0 = 32 * s + lane, rp = r0 - 32;
#pragma unroll
        for (int c = 0; c < 3; ++c) nxt[c] = s > 0 ? rhs[c * nf + rp] : 0.0;
        const int jtop = min(32, nf - 32 * s) - 1;
        // multiplier l(j, row) of column j = 32 s + lj for this lane's row: j - k, in the current slot when lane < lj
        auto mult = [&](int lj) {
            const bool own = lane < lj;
            const int k = own ? lj - lane : lj - lane + 32;
            return (lj >= 0 && k <= w && (own || s > 0)) ? D[k * LD + 32 * s + lj] : 0.0;
        };
        double l = mult(jtop);
        for (int lj = jtop; lj >= 0; --lj) {
            const double ln = mult(lj - 1);
            const double x0 = __shfl_sync(kFull, cur[0], lj), x1 = __shfl_sync(kFull, cur[1], lj), x2 = __shfl_sync(kFull, cur[2], lj);
            const double lo = lane < lj ? l : 0.0, lx = lane < lj ? 0.0 : l;
            cur[0] -= lo * x0; cur[1] -= lo * x1; cur[2] -= lo * x2;
            nxt[0] -= lx * x0; nxt[1] -= lx * x1; nxt[2] -= lx * x2;
            l = ln;
        }
        if (r0 < nf) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rhs[c * nf + r0] = cur[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) cur[c] = nxt[c];
    }
    __syncwarp();
}

// leading dimension of the band: odd for the CTA kernel (its trailing update walks band rows); for the warp kernel
// LD = 2 (mod 4), so that a lane-per-row walk (stride LD + 1, odd) is conflict free and a lane-per-offset walk (stride LD)
// meets two-way conflicts at most
__host__ __device__ inline int direct_ld(int nf, bool warp) { return warp ? ((nf + 1) & ~3) + 2 : (nf | 1); }

// One design, start to finish, by the `nt` threads of a group (a whole CTA, or one warp of it when WARP): tid = index in
// the group, smem = the group's shared memory.
// MODE 0: a CTA with the column-by-column factorisation; 1: one warp on its own; 2: a CTA (a "team") whose first warp
// factors and substitutes with the warp routines while all its warps share the node / edge passes
template <int MODE>
__device__ void eval_design(const DirectArgs& a, const int b, double* smem, const int tid, const int nt, int* s_bad) {
    constexpr bool WARP = MODE == 1, TEAM = MODE == 2;
    const PlanDev& P = a.P;
    const int nf = P.nf, n = P.n, E = P.E, w = P.w;
    const int LD = direct_ld(nf, WARP || TEAM);
    double* D = (MODE == 0 && a.band) ? a.band + (size_t)blockIdx.x * a.band_stride : smem;
    double* rhs = D + (size_t)(w + 1) * LD;
    double* lcol = (MODE == 0 && a.band) ? smem : rhs + 3 * (size_t)nf;         // CTA kernel only
    double* lsc = lcol + (w + 1);
    double* red = lsc + (w + 1);                 // 32 doubles: one per warp of the largest CTA
    int bad_w = 0;                               // warp kernel: the verdict lives in registers
    {
        const bool in_smem = a.state_off >= 0;
        double* st = smem + (in_smem ? a.state_off : 0);
        const double* q = a.q + (size_t)b * E;
        double* X = in_smem ? st + E : a.X + (size_t)b * n * 3;
        double* U = in_smem ? X + 3 * (size_t)n : a.U + (size_t)b * E * 3;
        double* Lg = in_smem ? U + 3 * (size_t)E : a.L + (size_t)b * E;
        double* F = in_smem ? Lg + E : a.F + (size_t)b * E;
        double* oX = a.oX ? a.oX + (size_t)b * n * 3 : nullptr;
        double* oR = a.oR ? a.oR + (size_t)b * n * 3 : nullptr;
        double* oU = a.oU ? a.oU + (size_t)b * E * 3 : nullptr;
        double* oL = a.oL ? a.oL + (size_t)b * E : nullptr;
        double* oF = a.oF ? a.oF + (size_t)b * E : nullptr;
        if (!WARP && tid == 0) *s_bad = 0;
        if (in_smem) {                               // q is read by five passes: once from global memory
            for (int e = tid; e < E; e += nt) st[e] = q[e];
            q = st;
            cta_sync(nt);
        }

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
        if (WARP) {
            bad_w = band_ldlt_warp(D, nf, w, LD);
        } else if (TEAM) {
            if (tid < 32 && band_ldlt_warp(D, nf, w, LD) && tid == 0) *s_bad = 1;
            __syncthreads();
        } else if (w == 1 && a.chain_off >= 0) {
            chain_factor(D, nf, LD, smem + a.chain_off, s_bad);
        } else if (w == 1) {
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
                if (bad) *s_bad = 1;
            }
            cta_sync(nt);
        } else {
            const int TX = 1 << a.tx_shift, tx = tid & (TX - 1), ty = tid >> a.tx_shift, TY = nt >> a.tx_shift;
            for (int j = 0; j < nf - 1; ++j) {
                int m = min(w, nf - 1 - j);
                double dj = D[j];
                if (tid == 0 && (dj == 0.0 || !isfinite(dj))) *s_bad = 1;
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
                if (dl == 0.0 || !isfinite(dl)) *s_bad = 1;
            }
        }

        // ---- forward solve and equilibrium state ------------------------------------------------------
        if (WARP) band_solve_warp(D, rhs, nf, w, LD);
        else if (TEAM) { if (tid < 32) band_solve_warp(D, rhs, nf, w, LD); __syncthreads(); }
        else if (w == 1 && a.chain_off >= 0) chain_solve(D, rhs, nf, LD, smem + a.chain_off);
        else band_solve(D, rhs, nf, w, LD, nt, a.k_shift);
        for (int i = tid; i < n; i += nt) {
            int src = P.node_src[i];
            double x0, x1, x2;
            if (src >= 0) {
                int p = P.pos[src];
                x0 = rhs[p]; x1 = rhs[nf + p]; x2 = rhs[2 * nf + p];
            } else {
                const double* xk = a.xfix + (size_t)(-1 - src) * 3;
                x0 = xk[0]; x1 = xk[1]; x2 = xk[2];
            }
            X[i * 3 + 0] = x0; X[i * 3 + 1] = x1; X[i * 3 + 2] = x2;
            if (oX) { oX[i * 3 + 0] = x0; oX[i * 3 + 1] = x1; oX[i * 3 + 2] = x2; }
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
            if (oU) { oU[e * 3 + 0] = u0; oU[e * 3 + 1] = u1; oU[e * 3 + 2] = u2; }
            if (oL) oL[e] = len;
            if (oF) oF[e] = f;
            lp_part += fabs(len * f);
        }
        cta_sync(nt);

        double loss_part = 0.0;
        const bool seeds = a.adjoint;
        // cotangents: in shared memory Xbar takes the place of X (a node's position is read by its own thread only, before it
        // writes Xbar of the same node)
        double* Rb = !seeds ? nullptr : (in_smem ? F + E : a.Rb + (size_t)b * n * 3);
        double* Xb = !seeds ? nullptr : (in_smem ? X : a.Xb + (size_t)b * n * 3);
        double res_max = 0.0, term_max = 0.0;        // residual of the solve on the free nodes against the size of its terms
        for (int i = tid; i < n; i += nt) {
            double r[3] = {a.loads[i * 3 + 0], a.loads[i * 3 + 1], a.loads[i * 3 + 2]};
            double terms = fmax(fabs(r[0]), fmax(fabs(r[1]), fabs(r[2])));
            for (int k = P.ninc_ptr[i]; k < P.ninc_ptr[i + 1]; ++k) {
                int e = P.ninc_edge[k];
                double s = (double)P.ninc_sign[k] * q[e];
                r[0] -= s * U[e * 3 + 0];
                r[1] -= s * U[e * 3 + 1];
                r[2] -= s * U[e * 3 + 2];
                terms = fmax(terms, fabs(s) * fmax(fabs(U[e * 3 + 0]), fmax(fabs(U[e * 3 + 1]), fabs(U[e * 3 + 2]))));
            }
            term_max = fmax(term_max, terms);
            if (P.node_src[i] >= 0) res_max = fmax(res_max, fmax(fabs(r[0]), fmax(fabs(r[1]), fabs(r[2]))));
            if (oR) { oR[i * 3 + 0] = r[0]; oR[i * 3 + 1] = r[1]; oR[i * 3 + 2] = r[2]; }     // R is consumed here: stored only on request
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
        double* Ub = !seeds ? nullptr : (in_smem ? F + E + 3 * (size_t)n : a.Ub + (size_t)b * E * 3);
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
            if (WARP) band_solve_warp(D, rhs, nf, w, LD);                                           // K symmetric: same factor
            else if (TEAM) { if (tid < 32) band_solve_warp(D, rhs, nf, w, LD); __syncthreads(); }
            else if (w == 1 && a.chain_off >= 0) chain_solve(D, rhs, nf, LD, smem + a.chain_off);
            else band_solve(D, rhs, nf, w, LD, nt, a.k_shift);
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
        // verdict: a zero / non-finite pivot, or a solve whose own residual is not small against its terms (tiny pivots of an
        // indefinite K: the answer would be inaccurate with every pivot non-zero).  Both map to LinAlgError on the host.
        const double rmax = cta_max(res_max, red, nt), tmax = cta_max(term_max, red, nt);
        const bool inaccurate = !(rmax <= kDirectResidualTol * tmax);
        const int bad = WARP ? bad_w : *s_bad;
        if (tid == 0 && a.status) a.status[b] = (bad || inaccurate) ? DFDM_STATUS_SINGULAR : DFDM_STATUS_OK;
        cta_sync(nt);
    }
}

__global__ void __launch_bounds__(256) k_direct_eval(DirectArgs a) {
    extern __shared__ double smem[];
    __shared__ int s_bad;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) eval_design<0>(a, b, smem, threadIdx.x, blockDim.x, &s_bad);
}

// one CTA per design with the warp routines for the factor and the substitutions: small batches, where a warp per design
// would leave most of every SM idle during the node / edge passes
__global__ void __launch_bounds__(256) k_direct_team(DirectArgs a) {
    extern __shared__ double smem[];
    __shared__ int s_bad;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) eval_design<2>(a, b, smem, threadIdx.x, blockDim.x, &s_bad);
}

// the same body for long chains: 1 024 threads (64 registers) - the partition solver is light, the node / edge passes over
// thousands of elements want the threads
__global__ void __launch_bounds__(1024) k_direct_chain(DirectArgs a) {
    extern __shared__ double smem[];
    __shared__ int s_bad;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) eval_design<0>(a, b, smem, threadIdx.x, blockDim.x, &s_bad);
}

// factors too large for shared memory: the band and the right-hand sides of a CTA's design live in global memory (its own
// slot of the workspace - L2 resident while it is worked on), only the column scratch is in shared memory.  Same body, same
// column-by-column sweep with two block barriers per column: ~2 us per column instead of ~0.2.  This is the path of last
// resort - mixed-sign q (indefinite K, which CG cannot solve) on networks beyond the shared-memory limit.
__global__ void __launch_bounds__(1024) k_direct_global(DirectArgs a) {
    extern __shared__ double smem[];
    __shared__ int s_bad;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) eval_design<0>(a, b, smem, threadIdx.x, blockDim.x, &s_bad);
}

// one warp per design; `stride` doubles of shared memory per warp
__global__ void __launch_bounds__(256) k_direct_warp(DirectArgs a, int stride) {
    extern __shared__ double smem[];
    const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    double* mine = smem + (size_t)warp * stride;
    for (int b = blockIdx.x * warps + warp; b < a.B; b += gridDim.x * warps) eval_design<1>(a, b, mine, threadIdx.x & 31, 32, nullptr);
}

// global-memory variant: CTAs (= band slots in the workspace) and doubles per slot
constexpr int kGlobalSlots = 296;
static int64_t global_band_stride(int nf, int w) {
    return (((int64_t)(w + 1) * direct_ld(nf, false) + 3 * (int64_t)nf) + 31) & ~(int64_t)31;
}

static thread_local const char* g_last_direct = "";
const char* last_direct_kernel() { return g_last_direct; }

int64_t direct_workspace(const Plan& P, int B, bool adjoint, bool has_prog) {
    (void)has_prog;
    Bump bump(nullptr);
    int64_t n = P.n, E = P.E;
    if (direct_smem_bytes(P.nf, P.w) > kDirectSmemLimit) bump.take<double>(std::min<int64_t>(B, kGlobalSlots) * global_band_stride(P.nf, P.w));
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
    const bool global_band = direct_smem_bytes(P.nf, P.w) > kDirectSmemLimit;
    if (global_band) {
        a.band_stride = global_band_stride(P.nf, P.w);
        a.band = bump.take<double>(std::min<int64_t>(B, kGlobalSlots) * a.band_stride);
    }
    a.X = bump.take<double>(B * n * 3);
    a.R = bump.take<double>(B * n * 3);
    a.U = bump.take<double>(B * E * 3);
    a.L = bump.take<double>(B * E);
    a.F = bump.take<double>(B * E);
    a.oX = ev.o_xyz; a.oR = ev.o_res; a.oU = ev.o_vec; a.oL = ev.o_len; a.oF = ev.o_for;
    a.state_off = -1;
    a.chain_off = -1;
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
    int dev = 0, sms = 0, per_sm = 0;
    DFDM_CUDA_OK(cudaGetDevice(&dev));
    DFDM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (global_band) {
        const int nt = 1024;
        int tx = 4;
        while (tx < 32 && tx < w) tx <<= 1;
        int shift = 0;
        while ((1 << shift) < tx) ++shift;
        a.tx_shift = shift;
        int kshift = 0;
        while ((1 << kshift) < w) ++kshift;
        a.k_shift = kshift;
        int64_t smem = 8 * (2 * (int64_t)(w + 1) + 32);
        if (w == 1 && P.nf >= 6) {
            a.chain_off = (int)(smem / 8);
            smem += 8 * (int64_t)kChainScratch;
        }
        if (smem > 200 * 1024) {
            set_error("direct solver: half-bandwidth " + std::to_string(w) + " is beyond the global-memory variant's column scratch");
            return DFDM_EUNSUPPORTED;
        }
        if (smem > 48 * 1024) {
            static std::mutex mu;
            static std::map<int, int64_t> configured;
            std::lock_guard<std::mutex> lock(mu);
            if (configured[dev] < smem) {
                DFDM_CUDA_OK(cudaFuncSetAttribute(k_direct_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured[dev] = smem;
            }
        }
        DFDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_direct_global, nt, (size_t)smem));
        if (per_sm < 1) per_sm = 1;
        const int grid = (int)std::min<int64_t>(std::min<int64_t>(B, kGlobalSlots), (int64_t)sms * per_sm);
        k_direct_global<<<grid, nt, (size_t)smem, ev.stream>>>(a);
        DFDM_LAUNCH_OK("k_direct_global");
        g_last_direct = "k_direct_global";
        return DFDM_OK;
    }
    // ---- one warp per design: narrow bands whose factor leaves room for several designs per SM ------------------------
    static const bool warp_off = std::getenv("DFDM_DIRECT_WARP") && std::atoi(std::getenv("DFDM_DIRECT_WARP")) == 0;
    const int64_t factor_part = (int64_t)(w + 1) * direct_ld(P.nf, true) + 3 * (int64_t)P.nf + 2 * (int64_t)(w + 1) + 16;   // doubles
    const int64_t state_part = 6 * n + 9 * E;                  // q, X / Xbar, U, L, F, Rbar, Ubar
    const int64_t per_design = factor_part + state_part;
    if (!warp_off && w >= 2 && w <= kWarpMaxW && per_design * 8 <= 100 * 1024) {
        a.state_off = (int)factor_part;
        // A warp per design fills the SMs only when there are many designs per SM.  Below `team_below` designs per SM a
        // CTA per design ("team": warp 0 factors and substitutes, every warp shares the passes over nodes and edges)
        // is the faster shape (dome, 512 designs = 3.5 per SM: 63.6 against 85.8 us; pillow / pringle at 6.9 per SM: a warp per
        // design wins by 1-6 %); DFDM_DIRECT_TEAM_BELOW / DFDM_DIRECT_TEAM_THREADS are there for the sweep that set them
        // (profiles/r02_direct_team_sweep.json).
        static const int team_below = std::getenv("DFDM_DIRECT_TEAM_BELOW") ? std::atoi(std::getenv("DFDM_DIRECT_TEAM_BELOW")) : 5;
        static const int team_threads = std::getenv("DFDM_DIRECT_TEAM_THREADS") ? std::atoi(std::getenv("DFDM_DIRECT_TEAM_THREADS")) : 128;
        if (B < (int64_t)team_below * sms) {
            const size_t smem = (size_t)per_design * 8;
            if (smem > 48 * 1024) {
                static std::mutex mu;
                static std::map<int, size_t> configured;
                std::lock_guard<std::mutex> lock(mu);
                if (configured[dev] < smem) {
                    DFDM_CUDA_OK(cudaFuncSetAttribute(k_direct_team, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    configured[dev] = smem;
                }
            }
            const int nt = std::max(32, std::min(256, team_threads & ~31));
            DFDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_direct_team, nt, smem));
            if (per_sm < 1) per_sm = 1;
            const int grid = (int)std::min<int64_t>(B, (int64_t)sms * per_sm);
            k_direct_team<<<grid, nt, smem, ev.stream>>>(a);
            DFDM_LAUNCH_OK("k_direct_team");
            g_last_direct = "k_direct_team";
            return DFDM_OK;
        }
        // designs per CTA: up to 8 warps, as many as fit ~100 KB so that two CTAs share an SM; fewer when the batch is small
        int warps = (int)std::max<int64_t>(1, std::min<int64_t>(8, (100 * 1024) / (per_design * 8)));
        while (warps > 1 && (B + warps - 1) / warps < sms) warps >>= 1;      // spread a small batch over the SMs first
        const size_t smem = (size_t)warps * per_design * 8;
        if (smem > 48 * 1024) {
            static std::mutex mu;
            static std::map<int, size_t> configured;
            std::lock_guard<std::mutex> lock(mu);
            if (configured[dev] < smem) {
                DFDM_CUDA_OK(cudaFuncSetAttribute(k_direct_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured[dev] = smem;
            }
        }
        DFDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_direct_warp, warps * 32, smem));
        if (per_sm < 1) per_sm = 1;
        const int64_t ctas = (B + warps - 1) / warps;
        const int grid = (int)std::min<int64_t>(ctas, (int64_t)sms * per_sm);
        k_direct_warp<<<grid, warps * 32, smem, ev.stream>>>(a, (int)per_design);
        DFDM_LAUNCH_OK("k_direct_warp");
        g_last_direct = "k_direct_warp";
        return DFDM_OK;
    }
    int nt = w <= 6 ? 32 : (w <= 12 ? 64 : (w <= 48 ? 128 : 256));
    static const bool chain_off = std::getenv("DFDM_DIRECT_CHAIN") && std::atoi(std::getenv("DFDM_DIRECT_CHAIN")) == 0;
    const bool chain = w == 1 && P.nf >= 6 && !chain_off;   // partition method: parts of >= 3 rows, at most 256 of them
    if (w == 1 && P.nf > 256) nt = 256;
    if (chain) nt = P.nf > 1024 ? 1024 : 256;               // one thread per part in the solver; the node / edge passes want the threads
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
    if (chain) {
        a.chain_off = (int)(smem / 8);
        smem += 8 * (int64_t)kChainScratch;
    }
    if (smem + 8 * (6 * n + 9 * E) <= 64 * 1024) {        // small enough to keep the state in shared memory as well
        a.state_off = (int)(smem / 8);
        smem += 8 * (6 * n + 9 * E);
    }
    auto kernel = nt > 256 ? k_direct_chain : k_direct_eval;
    if (smem > 48 * 1024) {                               // opt-in is per device and per function: set it wherever we launch
        static std::mutex mu;
        static std::map<std::pair<int, int>, int64_t> configured;
        std::lock_guard<std::mutex> lock(mu);
        int64_t& have = configured[{dev, nt > 256 ? 1 : 0}];
        if (have < smem) {
            DFDM_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            have = smem;
        }
    }
    DFDM_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, nt, (size_t)smem));
    if (per_sm < 1) per_sm = 1;
    int grid = (int)std::min<int64_t>(B, (int64_t)sms * per_sm);
    kernel<<<grid, nt, (size_t)smem, ev.stream>>>(a);
    DFDM_LAUNCH_OK("k_direct_eval");
    g_last_direct = nt > 256 ? "k_direct_chain" : "k_direct_eval";
    return DFDM_OK;
}

}  // namespace dfdm
