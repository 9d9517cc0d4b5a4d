/*
 * dfdm_b200 — C ABI of the B200-native force-density hot path.
 *
 * The reference (arpastrana/dfdm) is pure Python and has no FFI of its own; its boundary for this
 * path is the Python API of src/dfdm/equilibrium/model.py, src/dfdm/losses/losses.py and
 * src/dfdm/optimization.py.  Every entry point below names the reference interface it replaces
 * (paths relative to /root/reference).  The binding a maintainer would add on the reference side is
 * the ctypes stub shown in INTEGRATION.md (dfdm_b200/_lib.py is that stub, in use).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / CUDA types in signatures (streams travel as void*,
 *     i.e. a cudaStream_t; NULL = the legacy default stream)
 *   - device entry points take DEVICE pointers, run asynchronously on `stream`, never allocate after
 *     the workspace query, and leave per-design numerical verdicts in `status[B]`
 *   - *_host entry points take HOST pointers and include the host<->device copies
 *   - all floating point data is IEEE double; batches are design-major: q[B][E], xyz[B][n][3], ...
 *   - return value: 0 on success, a negative DFDM_E* code otherwise; dfdm_last_error() explains it
 *   - thread safety: a plan / goal program is immutable after creation and may be shared by any
 *     number of threads, streams and devices; concurrent DEVICE calls must use distinct workspaces.
 *     The *_host calls keep their device memory inside the plan, one arena per device: calls on the
 *     same (plan, device) are serialised by a lock held for the whole call; they run on private
 *     non-blocking streams, never on the legacy default stream
 */
#ifndef DFDM_B200_H
#define DFDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFDM_ABI_VERSION 3

#if defined(__GNUC__)
#define DFDM_API __attribute__((visibility("default")))
#else
#define DFDM_API
#endif

/* error codes */
#define DFDM_OK 0
#define DFDM_EINVAL (-1)     /* bad argument (NULL, negative size, index out of range) */
#define DFDM_ECUDA (-2)      /* CUDA runtime error; see dfdm_last_error() */
#define DFDM_ENOMEM (-3)     /* workspace too small / allocation failed */
#define DFDM_EUNSUPPORTED (-4)

/* per-design status written by the solvers (numpy.linalg.LinAlgError territory in model.py:55) */
#define DFDM_STATUS_OK 0
#define DFDM_STATUS_SINGULAR 1        /* zero / non-finite pivot in the direct factorisation */
#define DFDM_STATUS_NOT_CONVERGED 2   /* PCG hit maxiter (indefinite or ill-conditioned K) */
#define DFDM_STATUS_NONFINITE 3       /* NaN / Inf in the result */

/* solver selection */
#define DFDM_SOLVER_AUTO 0
#define DFDM_SOLVER_DIRECT 1   /* banded LDL^T, whole evaluation fused in one kernel: in shared memory (a warp or a CTA per design) when the
                                  factor fits; asked for by name beyond that, the same factorisation with the band in global memory
                                  (slow; the path that solves the indefinite K of a mixed-sign q at any size) */
#define DFDM_SOLVER_PCG 2      /* CG on batch-interleaved CSR, preconditioned by aggregation multigrid (or Jacobi: DFDM_PRECOND_*) */
/* Precision of the PCG path: the solution x, the residual r, K, K p and every dot product are double precision, and x and r
   move by the same stored vector (x += a p, r -= a K p), so r stays the true residual of x to double-precision rounding and
   the stop tests below bound the true error.  Single precision is used where it cannot reach the answer: the multigrid
   preconditioner (its operators, vectors and its copy of r) and the STORED search direction p (rounding p tilts the
   direction by 6e-8, like an inexact preconditioner: same iteration counts, same errors against SuperLU). */

typedef struct dfdm_plan dfdm_plan;
typedef struct dfdm_goalprog dfdm_goalprog;

typedef struct dfdm_options {
    int32_t solver;          /* DFDM_SOLVER_*; AUTO picks DIRECT when the band fits shared memory */
    int32_t pcg_maxiter;     /* <= 0: default 20 * sqrt(n_free) + 2000 */
    double pcg_rtol;         /* stop test of the forward solve, ||r|| <= rtol ||b|| per design and coordinate; <= 0: the calibrated
                                default DFDM_PCG_RTOL_DEFAULT (below) */
    int32_t pcg_check_every; /* > 0: the host blocks on the device "all converged" counter every k iterations;
                                <= 0 (default): the host stays two iterations ahead of the device and reads the counter of
                                every iteration as it completes, so the device never waits for the host and at most two
                                (early-exiting, empty) iterations are enqueued past convergence */
    int32_t pcg_precond;     /* DFDM_PRECOND_*; AUTO = aggregation multigrid V-cycle when the plan has coarse levels */
    double mg_omega;         /* damping of the Jacobi smoother; <= 0: default 0.8 */
    double mg_over;          /* scaling of the coarse-grid correction (plain aggregation under-corrects) on levels whose coarse
                                problem is solved once or exactly; the scalings of such nested levels multiply.
                                <= 0: 1.3 under a W-cycle, 1.7 for a plain V-cycle */
    int32_t mg_w_levels;     /* coarse levels 1..k are visited twice per cycle (W-cycle); 0: default 3; < 0: plain V-cycle */
    int32_t flags;           /* DFDM_FLAG_* bits; 0: defaults */
    double mg_over_w;        /* the same scaling on levels whose coarse problem is solved TWICE (the finest level and the W
                                levels but the last): two nested corrections of scaling s combine to 1 - (1 - s)^2 <= 1, so any
                                value below 2 keeps the preconditioner definite.  <= 0: 1.7 */
    double pcg_rtol_adjoint; /* the same stop test for the adjoint solve K lambda = xbar (gradients within 1e-7 need fewer digits
                                than coordinates within 1e-9); <= 0: DFDM_PCG_RTOL_ADJOINT_DEFAULT (_JACOBI under that preconditioner) */
} dfdm_options;

/* Calibrated against SuperLU / the dense oracle (tools/rtol_sweep.py, profiles/r02_rtol_sweep.jsonl and the GPU parity suite
   run with every tolerance tightened tenfold, DFDM_TEST_TOL_SCALE=0.1):
     forward 1e-12  coordinates within 5e-13 on the 258 x 258 grid, and edge lengths (differences of neighbouring
                    coordinates, 250 x smaller than the coordinates themselves) still within 1e-10; 1e-11 would save one
                    iteration of 17 and leave the lengths at 1.5e-9, so it stays
     adjoint 5e-9   under the multigrid preconditioner (the gradient contract is 1e-7): gradient error 2e-11 on the 258 x 258
                    grid, 12 instead of 17 iterations there
             5e-10  under the Jacobi preconditioner (networks without coarse levels, or DFDM_PRECOND_JACOBI): with a weak
                    preconditioner the error can exceed the residual by up to the condition number - measured 1.2 ... 3.6 x
                    the stop test on the small networks - and its iterations are cheap */
#define DFDM_PCG_RTOL_DEFAULT 1e-12
#define DFDM_PCG_RTOL_ADJOINT_DEFAULT 5e-9
#define DFDM_PCG_RTOL_ADJOINT_JACOBI 5e-10

/* option flags */
#define DFDM_FLAG_SPMV_MATRIX_FREE 1   /* PCG: K p straight from q over the node-edge incidence (8 E bytes per design) instead of
                                          the stored CSR values (8 nnz): 7 % faster SpMV on the quad grid, 0.8 % per evaluation */

#define DFDM_FLAG_MG_NO_TAIL 2         /* PCG: keep the coarse multigrid levels on the multi-launch path (one kernel per level and
                                          pass, cluster kernel at the bottom) instead of the one-CTA-per-design-pair tail kernel;
                                          same arithmetic, for A/B measurements and tests */

/* preconditioner of the PCG path */
#define DFDM_PRECOND_AUTO 0
#define DFDM_PRECOND_JACOBI 1      /* z = r / diag(K) */
#define DFDM_PRECOND_MULTIGRID 2   /* one V(1,1) cycle of plain-aggregation multigrid, damped-Jacobi smoothing */

typedef struct dfdm_plan_info {
    int32_t n, n_edges, n_free, n_fixed;
    int32_t nnz;                 /* stored entries of K = C_f^T Q C_f (full symmetric CSR, sorted columns) */
    int32_t bandwidth_natural;   /* half-bandwidth of K in ascending free-node order */
    int32_t bandwidth_solver;    /* half-bandwidth in the ordering the direct solver uses (RCM or natural) */
    int32_t direct_fits;         /* 1 if the banded factor + 3 right-hand sides fit one CTA's shared memory */
    int64_t direct_smem_bytes;
    int32_t auto_solver;         /* what DFDM_SOLVER_AUTO resolves to */
    int32_t mg_levels;           /* coarse levels of the aggregation multigrid hierarchy (0: Jacobi only) */
    int32_t mg_n_coarse;         /* rows of the first coarse level */
    int32_t mg_n_coarsest;       /* rows of the last coarse level */
} dfdm_plan_info;

/* index arrays exposed for bit-exact parity tests (host pointers owned by the plan) */
#define DFDM_ARR_FREE 0          /* int32[n_free]   structure.py:74-81  free_nodes  */
#define DFDM_ARR_FIXED 1         /* int32[n_fixed]  structure.py:83-90  fixed_nodes */
#define DFDM_ARR_FREEFIXED 2     /* int32[n]        structure.py:92-106 freefixed_nodes */
#define DFDM_ARR_ROWPTR 3        /* int32[n_free+1] CSR of K in free order */
#define DFDM_ARR_COLIDX 4        /* int32[nnz] */
#define DFDM_ARR_EDGE_SLOTS 5    /* int32[E][4]: CSR slots (uu, vv, uv, vu) an edge scatters into, -1 if the endpoint is fixed */
#define DFDM_ARR_SOLVER_PERM 6   /* int32[n_free]: free index at each position of the direct solver's ordering */
#define DFDM_ARR_NODE_INC_PTR 7  /* int32[n+1]  node -> incident edges */
#define DFDM_ARR_NODE_INC_EDGE 8 /* int32[2E] */
#define DFDM_ARR_NODE_INC_SIGN 9 /* int32[2E]: C[e, node] = -1 (tail u) or +1 (head v); model.py:23,34 */
#define DFDM_ARR_MG_ROWS 10      /* int32[mg_levels+1]: rows of every multigrid level, level 0 = n_free (empty without a hierarchy) */
#define DFDM_ARR_MG_NNZ 11       /* int32[mg_levels+1]: stored entries of every level's operator */
#define DFDM_ARR_PCG_PERM 12     /* int32[n_free]: free index of every row in the PCG path's hierarchical row order */
#define DFDM_ARR_MG_AGG 100      /* + l (0 <= l < mg_levels): int32[rows of level l], row of level l -> its aggregate on level l+1,
                                    rows in the PCG path's order; lets a host reference rebuild the hierarchy (tests) */

DFDM_API const char* dfdm_last_error(void);
DFDM_API int32_t dfdm_abi_version(void);

/*
 * Topology compiler.  Replaces EquilibriumStructure (src/dfdm/equilibrium/structure.py:11-106):
 * node_index / edge_index are the caller's row numbering; connectivity (structure.py:63-72) is kept
 * as the edge list itself (C[e,u] = -1, C[e,v] = +1); free_nodes / fixed_nodes / freefixed_nodes are
 * DFDM_ARR_FREE / FIXED / FREEFIXED.  Also builds, once, what model.py:36-55 rebuilds on every call:
 * the sparsity pattern of K and the scatter maps into it.
 */
DFDM_API int dfdm_plan_create(int32_t n, int32_t n_edges, const int32_t* edges /* [E][2] (u, v) */,
                     const uint8_t* is_fixed /* [n] */, dfdm_plan** plan);
DFDM_API void dfdm_plan_destroy(dfdm_plan* plan);
DFDM_API int dfdm_plan_get_info(const dfdm_plan* plan, dfdm_plan_info* info);
DFDM_API const int32_t* dfdm_plan_array(const dfdm_plan* plan, int32_t which, int64_t* length);

/*
 * Goal / loss program.  Replaces the object graph Loss -> LossTerm -> [Goal]
 * (src/dfdm/losses/losses.py:10-87, src/dfdm/goals/goal.py:23-33, goals/helpers.py:19-37) by a flat
 * table evaluated on the device.  Goal kinds: one per reference goal class.
 */
#define DFDM_GOAL_NODE_POINT 0               /* goals/nodegoal/pointgoal.py:11-28     params: target xyz */
#define DFDM_GOAL_NODE_LINE 1                /* pointgoal.py:31-44                    params: a xyz, b xyz */
#define DFDM_GOAL_NODE_PLANE 2               /* pointgoal.py:47-61                    params: origin xyz, normal xyz */
#define DFDM_GOAL_NODE_RESIDUAL_FORCE 3      /* goals/nodegoal/residualgoal.py:12-25  params: target */
#define DFDM_GOAL_NODE_RESIDUAL_VECTOR 4     /* residualgoal.py:28-39                 params: target xyz */
#define DFDM_GOAL_NODE_RESIDUAL_DIRECTION 5  /* residualgoal.py:42-70                 params: target xyz (normalised on device) */
#define DFDM_GOAL_EDGE_LENGTH 6              /* goals/edgegoal/lengthgoal.py:7-19     params: target */
#define DFDM_GOAL_EDGE_FORCE 7               /* goals/edgegoal/forcegoal.py:7-19      params: target */
#define DFDM_GOAL_EDGE_LOAD_PATH 8           /* goals/edgegoal/loadpathgoal.py:7-22   params: target */
#define DFDM_GOAL_EDGE_VECTOR_ANGLE 9        /* goals/edgegoal/anglegoal.py:7-32      params: target (degrees), vector xyz */
#define DFDM_GOAL_EDGE_DIRECTION 10          /* goals/edgegoal/directiongoal.py:7-25  params: target xyz */
#define DFDM_GOAL_NETWORK_LOAD_PATH 11       /* goals/networkgoal/loadpathgoal.py:7-23 params: target */
#define DFDM_N_GOAL_KINDS 12
#define DFDM_GOAL_NPARAMS 6

#define DFDM_TERM_SQUARED_ERROR 0        /* losses.py:38-45  alpha * sum(w (p - t)^2)  */
#define DFDM_TERM_MEAN_SQUARED_ERROR 1   /* losses.py:48-55  alpha * mean(w (p - t)^2) over all entries of the term */
#define DFDM_TERM_PREDICTION_ERROR 2     /* losses.py:58-64  alpha * sum(p) */

DFDM_API int dfdm_goalprog_create(const dfdm_plan* plan, int32_t n_goals,
                         const int32_t* kind /* [G] DFDM_GOAL_* */,
                         const int32_t* element /* [G] node or edge row index; ignored for network goals */,
                         const int32_t* term /* [G] index into the term arrays */,
                         const double* weight /* [G] */,
                         const double* params /* [G][DFDM_GOAL_NPARAMS] */,
                         int32_t n_terms, const int32_t* term_kind /* [T] DFDM_TERM_* */,
                         const double* term_alpha /* [T] */,
                         double l2_alpha /* sum of L2Regularizer alphas, regularizers.py:19-20 */,
                         dfdm_goalprog** prog);
DFDM_API void dfdm_goalprog_destroy(dfdm_goalprog* prog);

/* bytes of device scratch an evaluation of B designs needs (state outputs the caller does not ask for live here) */
DFDM_API int64_t dfdm_workspace_bytes(const dfdm_plan* plan, int32_t batch, const dfdm_options* options);

/*
 * EquilibriumModel.__call__ (src/dfdm/equilibrium/model.py:63-79) for B designs.
 * Outputs (any may be NULL): the EquilibriumState fields of state.py:11-18.
 */
DFDM_API int dfdm_forward(const dfdm_plan* plan, int32_t batch,
                 const double* q /* [B][E] */, const double* loads /* [n][3] model.py:18 */,
                 const double* xyz_fixed /* [n_fixed][3] model.py:19 */,
                 double* xyz /* [B][n][3] */, double* residuals /* [B][n][3] */, double* vectors /* [B][E][3] */,
                 double* lengths /* [B][E] */, double* forces /* [B][E] */, int32_t* status /* [B] */,
                 void* workspace, int64_t workspace_bytes, const dfdm_options* options, void* stream);

/*
 * Loss.__call__ (losses.py:76-81) and grad(loss) (src/dfdm/optimization.py:61) in one pass:
 * forward, goal / loss evaluation, closed-form adjoint with one extra solve on the same operator.
 * dq may be NULL (value only).  State outputs as in dfdm_forward (any may be NULL).
 */
DFDM_API int dfdm_loss_and_grad(const dfdm_plan* plan, const dfdm_goalprog* prog, int32_t batch,
                       const double* q, const double* loads, const double* xyz_fixed,
                       double* loss /* [B] */, double* dq /* [B][E] */,
                       double* xyz, double* residuals, double* vectors, double* lengths, double* forces,
                       int32_t* status, void* workspace, int64_t workspace_bytes,
                       const dfdm_options* options, void* stream);

/*
 * Generic reverse mode (what autograd's tape computes for an arbitrary scalar of the state):
 * q_bar for given cotangents of (xyz, residuals, vectors, lengths, forces); any cotangent may be NULL.
 * Used for constraint Jacobians (optimization.py:115-130) and torch.autograd interop.
 */
DFDM_API int dfdm_vjp(const dfdm_plan* plan, int32_t batch,
             const double* q, const double* loads, const double* xyz_fixed,
             const double* xyz_bar /* [B][n][3] */, const double* residuals_bar /* [B][n][3] */,
             const double* vectors_bar /* [B][E][3] */, const double* lengths_bar /* [B][E] */,
             const double* forces_bar /* [B][E] */,
             double* dq /* [B][E] */, int32_t* status, void* workspace, int64_t workspace_bytes,
             const dfdm_options* options, void* stream);

/*
 * Host-buffer variants: copy inputs to the current device, evaluate, copy results back, synchronise.
 * Device memory is cached inside the library per (plan, device) and grown on demand.
 * The copies run on streams of their own, chained to the compute stream by events, and the batch can be cut into
 * chunks of designs whose copies (chunk k+1 in, chunk k-1 out) run under the kernels of chunk k, one cudaMemcpyAsync
 * per chunk and array: the direct path does so from 32 MB of q per chunk on (at most 4 chunks); the PCG path evaluates
 * the batch whole (measured: its latency-bound multigrid levels make two half batches 26 % slower than one whole,
 * more than the hidden copies return; DFDM_HOST_CHUNKS=k forces k chunks for experiments).  Pinned host buffers are
 * needed for any overlap (pageable ones still work, serialised by the driver).
 * Threads: the calls may be made from several host threads on one plan and device.  Each call in flight owns one of two
 * cached arenas (memory, streams) per (plan, device); the calls take turns with the kernels and copy outside their turn,
 * so with two threads, each evaluating batches of its own, the copies of one call run under the kernels of the other -
 * the way to keep a GPU busy with a stream of batches whose buffers live on the host (bench.py's `e2e`).  A third
 * concurrent call waits for the first arena.
 */
DFDM_API int dfdm_forward_host(const dfdm_plan* plan, int32_t batch, const double* q, const double* loads, const double* xyz_fixed,
                      double* xyz, double* residuals, double* vectors, double* lengths, double* forces,
                      int32_t* status, const dfdm_options* options);
DFDM_API int dfdm_loss_and_grad_host(const dfdm_plan* plan, const dfdm_goalprog* prog, int32_t batch,
                            const double* q, const double* loads, const double* xyz_fixed,
                            double* loss, double* dq, int32_t* status, const dfdm_options* options);

/* kernels launched by this library since load (all threads); bench.py reports the delta as gpu_launches */
DFDM_API int64_t dfdm_launch_count(void);
/* PCG iterations used by the most recent evaluation on this thread (forward solve, adjoint solve) */
DFDM_API void dfdm_last_pcg_iterations(int32_t* forward_iters, int32_t* adjoint_iters);
/* name of the kernel the most recent direct-path evaluation on this thread launched ("k_direct_warp", "k_direct_team",
   "k_direct_chain", "k_direct_eval", "k_direct_global"; "" before the first): which launch shape a plan and batch got */
DFDM_API const char* dfdm_last_direct_kernel(void);

/*
 * Per-kernel timing of the PCG loop for roofline accounting: when enabled, a CUDA event is recorded on
 * the launch stream after every PCG kernel and resolved at the convergence polls.
 * ms / count are indexed by DFDM_PROF_* (8 entries each).  Not thread safe; meant for bench.py.
 */
#define DFDM_PROF_SPMV 0       /* k_pcg_spmv:      Ap = K p, p.Ap            algorithmic bytes 8 nnz + 36 n_free per design (p is stored in single precision) */
#define DFDM_PROF_UPDATE 1     /* k_pcg_update:    r -= alpha Ap, r.r (r.z)  84 n_free (multigrid: with the single-precision copy of r the cycle reads) / 80 n_free (Jacobi) */
#define DFDM_PROF_DIRECTION 2  /* k_pcg_direction: p = z + beta p, with its share of k_pcg_xupdate (x += the alpha p of 8 iterations in one
                                  pass): 54 n_free (multigrid) / 74 n_free (Jacobi) */
#define DFDM_PROF_MG_DOWN 3    /* k_mg_down, finest level: t = K D^-1 r, restricted residual   4 nnz + 24 n_free + 12 n_coarse (single-precision cycle) */
#define DFDM_PROF_MG_UP 4      /* k_mg_up, finest level: z = M r, r.z                          4 nnz + 40 n_free + 12 n_coarse */
#define DFDM_PROF_MG_COARSE 5  /* all kernels below the finest level, per V-cycle */
#define DFDM_PROF_ASSEMBLE 6   /* k_assemble: K values, 1/diag and b from q          8 E + 8 nnz + 32 n_free per design */
#define DFDM_PROF_STATE 7      /* k_positions .. k_adjoint_rhs as one interval (loss_and_grad only): state, goals / loss,
                                  reverse edge passes                                ~ 184 E + 240 n per design */
DFDM_API void dfdm_profile_enable(int32_t on);
DFDM_API void dfdm_profile_read(double* ms /* [8] */, int64_t* count /* [8] */);

/*
 * Peer-memory all-gather for the multi-GPU path (one process per GPU on one node).  The reference evaluates one design
 * per call in one process (src/dfdm/optimization.py:89-97) and has nothing to bind here; these entry points serve
 * dfdm_b200/distributed.py::PeerGather, the one exchange of a sharded batched evaluation (losses and gradients).
 *
 *   dfdm_peer_alloc   device buffer that other processes can map; `handle` receives 64 bytes to send to the peers
 *   dfdm_peer_open    map a peer's buffer from its handle (CUDA IPC, peer access enabled on demand)
 *   dfdm_peer_push    ONE kernel copies `bytes` from src into every dst_bases[d] + dst_offset_bytes (st.global to peer
 *                     addresses over NVLink / NVSwitch; the own buffer may be one of the destinations).  Asynchronous on
 *                     `stream`; the caller orders the readers (e.g. a barrier / tiny all-reduce after it).
 *                     bytes % 8 == 0, src and dst_offset_bytes 16-byte aligned, at most 16 destinations.
 */
DFDM_API int32_t dfdm_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle /* [64] */);
DFDM_API int32_t dfdm_peer_open(const uint8_t* handle /* [64] */, void** ptr);
DFDM_API int32_t dfdm_peer_close(void* ptr);
DFDM_API int32_t dfdm_peer_free(void* ptr);
DFDM_API int32_t dfdm_peer_push(const void* src, int64_t bytes, int64_t dst_offset_bytes, void* const* dst_bases, int32_t n_dst,
                                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFDM_B200_H */
