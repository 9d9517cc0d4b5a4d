// Shared device / host helpers: error handling, the evaluation request, goal and adjoint mathematics.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <string>

#include "plan.hpp"

namespace dfdm {

extern std::atomic<long long> g_launches;

#define DFDM_CUDA_OK(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _err = (expr);                                                                  \
        if (_err != cudaSuccess) {                                                                  \
            dfdm::set_error(std::string(#expr) + ": " + cudaGetErrorString(_err));                  \
            return DFDM_ECUDA;                                                                      \
        }                                                                                           \
    } while (0)

#define DFDM_LAUNCH_OK(name)                                                                        \
    do {                                                                                            \
        dfdm::g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        cudaError_t _err = cudaGetLastError();                                                      \
        if (_err != cudaSuccess) {                                                                  \
            dfdm::set_error(std::string("launch of ") + name + ": " + cudaGetErrorString(_err));    \
            return DFDM_ECUDA;                                                                      \
        }                                                                                           \
    } while (0)

// One evaluation request, resolved to device pointers.  All batch arrays are design-major.
struct Eval {
    PlanDev P;                     // rows in ascending free order (direct path, ABI)
    PlanDev Pg;                    // rows in the PCG path's hierarchical order
    MgDev M;
    GoalProgDev G;
    bool has_prog = false;
    bool adjoint = false;          // compute dq
    int B = 0;
    const double* q = nullptr;     // [B][E]
    const double* loads = nullptr; // [n][3]
    const double* xfix = nullptr;  // [nx][3]
    // state outputs requested by the caller (NULL = not requested)
    double* o_xyz = nullptr;
    double* o_res = nullptr;
    double* o_vec = nullptr;
    double* o_len = nullptr;
    double* o_for = nullptr;
    // external cotangents (vjp entry)
    const double* c_xyz = nullptr;
    const double* c_res = nullptr;
    const double* c_vec = nullptr;
    const double* c_len = nullptr;
    const double* c_for = nullptr;
    double* loss = nullptr;        // [B]
    double* dq = nullptr;          // [B][E]
    int32_t* status = nullptr;     // [B]
    char* ws = nullptr;
    int64_t ws_bytes = 0;
    dfdm_options opt{};
    cudaStream_t stream = nullptr;
};

// Bump allocator over the caller's workspace (256-byte aligned blocks).  With base == nullptr it only counts.
struct Bump {
    char* base;
    int64_t off = 0;
    explicit Bump(char* b) : base(b) {}
    template <typename T>
    T* take(int64_t count) {
        int64_t bytes = (count * (int64_t)sizeof(T) + 255) & ~(int64_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

int64_t direct_workspace(const Plan& P, int B, bool adjoint, bool has_prog);
int64_t pcg_workspace(const Plan& P, int B);
int run_direct(Eval& ev);
int run_pcg(Eval& ev);

// ------------------------------------------------------------------------------------------------
// goal mathematics on registers (layout independent); see include/dfdm_b200.h for the kinds
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Contribution of one squared-error / prediction-sum entry: returns the loss term and the cotangent of p.
__device__ __forceinline__ double entry(const GoalRec& g, double p, double t, double& pbar) {
    if (g.mode == 0) {
        double coef = g.scale * g.weight;
        double d = p - t;
        pbar = 2.0 * coef * d;
        return coef * d * d;
    }
    pbar = g.scale;
    return g.scale * p;
}

// All goals attached to one node.  X, R: position and residual of the node.  Accumulates the direct
// cotangents xb (of X) and rb (of R); returns the loss contribution.
__device__ inline double node_goals(const GoalRec* __restrict__ recs, int g0, int g1, const double X[3], const double R[3],
                                    double xb[3], double rb[3]) {
    double loss = 0.0;
    for (int gi = g0; gi < g1; ++gi) {
        const GoalRec g = recs[gi];
        double pb[3];
        switch (g.kind) {
            case DFDM_GOAL_NODE_POINT: {
                for (int c = 0; c < 3; ++c) { loss += entry(g, X[c], g.p[c], pb[c]); xb[c] += pb[c]; }
            } break;
            case DFDM_GOAL_NODE_LINE: {
                // target = a + ((X - a).ab / ab.ab) ab   (compas closest_point_on_line; pointgoal.py:38-44)
                double ab[3] = {g.p[3] - g.p[0], g.p[4] - g.p[1], g.p[5] - g.p[2]};
                double ap[3] = {X[0] - g.p[0], X[1] - g.p[1], X[2] - g.p[2]};
                double s = dot3(ap, ab) / dot3(ab, ab);
                for (int c = 0; c < 3; ++c) { loss += entry(g, X[c], g.p[c] + ab[c] * s, pb[c]); xb[c] += pb[c]; }
            } break;
            case DFDM_GOAL_NODE_PLANE: {
                // target = X - ((X - o).n^) n^   (compas closest_point_on_plane; pointgoal.py:54-61)
                double nn = sqrt(g.p[3] * g.p[3] + g.p[4] * g.p[4] + g.p[5] * g.p[5]);
                double nh[3] = {g.p[3] / nn, g.p[4] / nn, g.p[5] / nn};
                double d = nh[0] * g.p[0] + nh[1] * g.p[1] + nh[2] * g.p[2];
                double k = (dot3(nh, X) - d) / dot3(nh, nh);
                for (int c = 0; c < 3; ++c) { loss += entry(g, X[c], X[c] - k * nh[c], pb[c]); xb[c] += pb[c]; }
            } break;
            case DFDM_GOAL_NODE_RESIDUAL_FORCE: {
                double nr = sqrt(dot3(R, R));
                double sb;
                loss += entry(g, nr, g.p[0], sb);
                for (int c = 0; c < 3; ++c) rb[c] += sb * R[c] / nr;
            } break;
            case DFDM_GOAL_NODE_RESIDUAL_VECTOR: {
                for (int c = 0; c < 3; ++c) { loss += entry(g, R[c], g.p[c], pb[c]); rb[c] += pb[c]; }
            } break;
            case DFDM_GOAL_NODE_RESIDUAL_DIRECTION: {
                double nr = sqrt(dot3(R, R));
                double nt = sqrt(g.p[0] * g.p[0] + g.p[1] * g.p[1] + g.p[2] * g.p[2]);
                double ph[3] = {R[0] / nr, R[1] / nr, R[2] / nr};
                for (int c = 0; c < 3; ++c) loss += entry(g, ph[c], g.p[c] / nt, pb[c]);
                double pp = dot3(pb, ph);
                for (int c = 0; c < 3; ++c) rb[c] += (pb[c] - pp * ph[c]) / nr;
            } break;
            default: break;
        }
    }
    return loss;
}

// All goals attached to one edge.  Accumulates the direct cotangents ub (of U), lb (of L), fb (of F).
__device__ inline double edge_goals(const GoalRec* __restrict__ recs, int g0, int g1, const double U[3], double L, double F,
                                    double ub[3], double& lb, double& fb) {
    double loss = 0.0;
    for (int gi = g0; gi < g1; ++gi) {
        const GoalRec g = recs[gi];
        double sb;
        switch (g.kind) {
            case DFDM_GOAL_EDGE_LENGTH: loss += entry(g, L, g.p[0], sb); lb += sb; break;
            case DFDM_GOAL_EDGE_FORCE: loss += entry(g, F, g.p[0], sb); fb += sb; break;
            case DFDM_GOAL_EDGE_LOAD_PATH: {
                double lf = L * F;
                double sg = lf > 0.0 ? 1.0 : (lf < 0.0 ? -1.0 : 0.0);
                loss += entry(g, fabs(lf), g.p[0], sb);
                lb += sb * sg * F;
                fb += sb * sg * L;
            } break;
            case DFDM_GOAL_EDGE_VECTOR_ANGLE: {
                // degrees(arccos(clamp(U.v / (|U||v|)))) with Python max/min clamp: constant (zero gradient) when clipped
                const double* v = &g.p[1];
                double nu = sqrt(dot3(U, U)), nv = sqrt(dot3(v, v));
                double a = dot3(U, v) / (nu * nv);
                bool clipped = false;
                if (a > 1.0) { a = 1.0; clipped = true; }
                if (a < -1.0) { a = -1.0; clipped = true; }
                const double r2d = 57.295779513082320876798154814105;
                loss += entry(g, acos(a) * r2d, g.p[0], sb);
                if (!clipped) {
                    double dang = -r2d / sqrt(1.0 - a * a);
                    for (int c = 0; c < 3; ++c) ub[c] += sb * dang * (v[c] / (nu * nv) - a * U[c] / (nu * nu));
                }
            } break;
            case DFDM_GOAL_EDGE_DIRECTION: {
                double nu = sqrt(dot3(U, U));
                double nt = sqrt(g.p[0] * g.p[0] + g.p[1] * g.p[1] + g.p[2] * g.p[2]);
                double ph[3] = {U[0] / nu, U[1] / nu, U[2] / nu};
                double pb[3];
                for (int c = 0; c < 3; ++c) loss += entry(g, ph[c], g.p[c] / nt, pb[c]);
                double pp = dot3(pb, ph);
                for (int c = 0; c < 3; ++c) ub[c] += (pb[c] - pp * ph[c]) / nu;
            } break;
            default: break;
        }
    }
    return loss;
}

// Network goals: given the total load path, returns the loss and d(loss)/d(load path).
__device__ inline double network_goals(const GoalRec* __restrict__ recs, int g0, int g1, double loadpath, double& lpbar) {
    double loss = 0.0;
    lpbar = 0.0;
    for (int gi = g0; gi < g1; ++gi) {
        const GoalRec g = recs[gi];
        double sb;
        loss += entry(g, loadpath, g.p[0], sb);
        lpbar += sb;
    }
    return loss;
}

// Edge part of the reverse pass before the solve.
//   in : direct cotangents ub, lb, fb; G = Rbar_v - Rbar_u; q, L, U, F; lpbar = d loss / d(sum |L F|)
//   out: ub = total cotangent of U; returns the partial dq (without the solve term)
__device__ __forceinline__ double edge_adjoint(double q, const double U[3], double L, double F, const double G[3], double lpbar,
                                               double ub[3], double lb, double fb) {
    if (lpbar != 0.0) {
        double lf = L * F;
        double sg = lf > 0.0 ? 1.0 : (lf < 0.0 ? -1.0 : 0.0);
        lb += lpbar * sg * F;
        fb += lpbar * sg * L;
    }
    double dq = fb * L;          // F = q L
    lb += fb * q;
    dq -= dot3(G, U);            // R = P - C^T diag(q) U
    for (int c = 0; c < 3; ++c) ub[c] -= q * G[c];
    if (lb != 0.0) {             // L = |U|
        double s = lb / L;
        for (int c = 0; c < 3; ++c) ub[c] += s * U[c];
    }
    return dq;
}

}  // namespace dfdm
