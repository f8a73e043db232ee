/* stereo_b200.h -- C ABI of libstereo_b200.so
 *
 * B200-native (sm_100a) replacement for the data-parallel hot path of
 * johannesu/stereo: the TRW-S and QPBO fusion solvers behind trws.m / rd.m and
 * the cost-volume / unary / pairwise array builders of dispmap_*.m.
 *
 * Conventions
 *  - plain C, host pointers unless a name says `_dev`; MATLAB (column-major)
 *    layouts exactly as the reference mex gateways receive them;
 *  - every entry point returns 0 on success and a negative SB_E* code on
 *    failure; sb_last_error() returns the message (thread-local);
 *  - there is NO CPU fallback: without a CUDA device every compute entry
 *    point fails with SB_ENODEV;
 *  - node index u = r + H*c (0-based, column-major; dispmap_super.m:281-282);
 *    pairwise terms in the order of dispmap_super.construct_neighborhood
 *    (dispmap_super.m:279-302): vertical down, vertical up, horizontal right,
 *    horizontal left, each column-major over the start node.
 *
 * Each declaration cites the reference interface it replaces.
 */
#ifndef STEREO_B200_H
#define STEREO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK        0
#define SB_EINVAL   -1   /* bad argument (the reference ASSERTs, cppmatrix.h:20-24)      */
#define SB_ENODEV   -2   /* no CUDA device / driver                                      */
#define SB_ECUDA    -3   /* CUDA runtime error                                           */
#define SB_ENOTGRID -4   /* connectivity is not the 4-connected grid of dispmap_super    */
#define SB_ENOMEM   -5   /* device memory exhausted                                      */
#define SB_EUNSUP   -6   /* unsupported size (e.g. more than SB_MAX_LABELS labels)       */

#define SB_MAX_LABELS 256

#if defined(__GNUC__)
#define SB_API __attribute__((visibility("default")))
#else
#define SB_API
#endif

/* ---------------------------------------------------------------- library */

/* "stereo_b200 x.y (sm_100a)" */
SB_API const char *sb_version(void);
/* message of the last failing call on this thread ("" if none) */
SB_API const char *sb_last_error(void);
/* number of visible CUDA devices (0 if none; never fails) */
SB_API int sb_device_count(void);
/* select the device used by subsequent calls of this thread's process */
SB_API int sb_set_device(int device);
/* number of kernels launched by the library since load (bench.py gpu_launches) */
SB_API int64_t sb_kernel_launches(void);

/* ---------------------------------------------------------------- options */

/* Arithmetic type of the TRW-S message sweep.  SB_F32 is the product path
 * (north star: energies within 1e-4 relative of the reference's doubles);
 * SB_F64 runs the same kernels in double for tight parity checks. */
#define SB_F32 0
#define SB_F64 1

typedef struct sb_trws_options {
    double maxiter;       /* trws_mex.cpp:39  default 1000 */
    double max_relgap;    /* trws_mex.cpp:40  default 0    */
    int    precision;     /* SB_F32 (default) | SB_F64     */
    int    fuse_rounding; /* 1 (default): primal rounding of iteration t rides in the
                             forward sweep of t+1; 0: separate sweep   */
    int    col_blocks;    /* grid-native entry, world > 1: column blocks per rank (the blocks are dealt round robin
                             to the ranks; 1 = contiguous bands); 0 = default (1: more blocks shorten the pipeline fill but
                             add NVLink hand-overs to the critical path -- measured slower on 8 GPUs) */
    int    latency_mode;  /* grid-native entry: 0 = automatic, 1 = always, -1 = never run the LATENCY build of the sweep
                             kernel (at most two strip walkers per SM, no register spills, operands taken ahead of the
                             dependent chain: shorter node steps, ~9 % less throughput).  Automatic: when a rank's nodes per SM
                             are fewer than 3 (H + W) (6 (H + W) in a banded run), i.e. when the pass ends with the DAG's
                             critical path rather than with the walkers' work, DESIGN.md 4 / 6 */
    int    reserved[4];
} sb_trws_options;

SB_API void sb_trws_default_options(sb_trws_options *opt);

typedef struct sb_trws_timing {
    double setup_ms;       /* upload, conversion, rank tables, schedule               */
    double solve_ms;       /* all sweeps (CUDA events on the solver stream)           */
    double sweep_ms_avg;   /* solve_ms / iterations                                   */
    double download_ms;
    int64_t kernel_launches;
    double sweep_kernel_ms;       /* sum of the sweep kernels' own durations (CUDA events
                                     around each launch on the solver stream)            */
    int64_t sweep_kernel_launches;
    int64_t reserved[1];
} sb_trws_timing;

/* ---------------------------------------------------------------- TRW-S */

/* Replaces mexFunction of cpp/trws_mex.cpp:149-163 (called from trws.m:33).
 *   kernel   1 = truncated linear (TypeStereoLinear), 2 = truncated quadratic
 *            (TypeStereoQuadratic); anything else -> SB_EINVAL ("Unsupported kernel")
 *   unary    L x N  (label fastest)                       trws_mex.cpp:31
 *   conn     2 x E  uint32, 0-based: conn[2p]=tail, conn[2p+1]=head
 *   q        L x E  positions of the head's labels        trws_mex.cpp:33,101-105
 *   qprim    L x E  positions of the tail's labels        trws_mex.cpp:34,107-111
 *   alphas   E      per-term weight                       trws_mex.cpp:35
 *   tol      lambda (truncation)                          trws_mex.cpp:36-37
 * Pairwise term p: V(k_tail,k_head) = alphas[p]*min(|q[k_head,p]-qprim[k_tail,p]|^kernel, tol)
 * Outputs (trws_mex.cpp:134-143): labels N doubles, 1-BASED; energy; lower bound;
 * iterations.  `timing` may be NULL.
 * The connectivity must be the dispmap_super grid (both directions of every
 * neighbour pair, reference order); otherwise SB_ENOTGRID. */
SB_API int sb_trws_solve(int kernel, int L, int64_t N, int64_t E,
                  const double *unary, const uint32_t *conn,
                  const double *q, const double *qprim,
                  const double *alphas, double tol,
                  const sb_trws_options *opt,
                  double *labels, double *energy, double *lower_bound,
                  double *iterations, sb_trws_timing *timing);

/* Resident-solver form of the same path, for callers that keep the problem in HBM
 * across calls (bench.py; dispmap_super.simultaneous_fusion re-solves):
 *   sb_trws_create     = the graph build of solve_mrf, cpp/trws_mex.cpp:58-121
 *                        (AddNode / AddEdge / SetAutomaticOrdering) -- uploads and
 *                        converts the inputs, builds rank tables and the schedule;
 *   sb_trws_reset      = MRFEnergy::ZeroMessages, cpp/trw-s/MRFEnergy.cpp:115-131;
 *   sb_trws_minimize   = MRFEnergy::Minimize_TRW_S, cpp/trw-s/minimize.cpp:7-116
 *                        (continues from the current messages, like the reference);
 *   sb_trws_get_labels = the GetSolution loop, cpp/trws_mex.cpp:134-139 (1-based);
 *   sb_trws_destroy    = delete mrf, cpp/trws_mex.cpp:146. */
typedef struct sb_trws_solver sb_trws_solver;
SB_API int sb_trws_create(int kernel, int L, int64_t N, int64_t E,
                   const double *unary, const uint32_t *conn,
                   const double *q, const double *qprim,
                   const double *alphas, double tol,
                   const sb_trws_options *opt, sb_trws_solver **out);
SB_API int sb_trws_reset(sb_trws_solver *s);
SB_API int sb_trws_minimize(sb_trws_solver *s, double maxiter, double max_relgap,
                     double *energy, double *lower_bound, double *iterations,
                     sb_trws_timing *timing);
SB_API int sb_trws_get_labels(sb_trws_solver *s, double *labels);
SB_API void sb_trws_destroy(sb_trws_solver *s);

/* Several GPUs of one box: the image rows are split into `world` contiguous
 * bands, rank r sweeps the strips of band r in the SAME order / orientation DAG as the single
 * GPU sweep, and the messages (and rounded-label positions) that cross a band boundary are
 * pushed straight into the neighbouring GPU's memory over NVLink as self-validating 64-bit
 * words (value | launch epoch), so the receiving sweep simply polls its own memory -- no host
 * round trip, no flag, no fence inside a pass.  Every rank is one process and holds the whole
 * problem (static data replicated; dynamic data homed per band); fp32 only.
 *   sb_trws_create_banded  like sb_trws_create, for rank `rank` of `world`
 *   sb_trws_ipc_export     3 x 64 bytes: CUDA IPC handles of this rank's message, mailbox and
 *                          selected-position arrays, to be handed to ranks rank-1 / rank+1
 *   sb_trws_ipc_attach     the handles exported by rank-1 (`up`) and rank+1 (`down`); NULL at the ends
 *   sb_trws_pass           one sweep on this rank: pass 0 forward / 1 backward, mode bit 0 = send
 *                          messages, bit 1 = primal rounding (forward only); acc[0] = this rank's
 *                          part of the energy, acc[1] = of the lower bound.  All ranks must call it
 *                          together and all-reduce acc (the stop rule of minimize.cpp:97-112 needs
 *                          the sums); consecutive passes must be separated by that collective. */
SB_API int sb_trws_create_banded(int kernel, int L, int64_t N, int64_t E,
                          const double *unary, const uint32_t *conn,
                          const double *q, const double *qprim,
                          const double *alphas, double tol,
                          const sb_trws_options *opt, int rank, int world, sb_trws_solver **out);
SB_API int sb_trws_ipc_export(sb_trws_solver *s, unsigned char *handles /* 192 bytes */);
SB_API int sb_trws_ipc_attach(sb_trws_solver *s, const unsigned char *up, const unsigned char *down);
SB_API int sb_trws_pass(sb_trws_solver *s, int pass, int mode, double *acc /* 2 */);

/* ---------------------------------------------------------------- TRW-S, grid-native entry
 *
 * The same solver as sb_trws_solve, but the problem enters the way
 * dispmap_super.simultaneous_fusion HOLDS it (dispmap_super.m:158-188) instead of as the L x E arrays
 * it derives from that: L plane proposals per pixel, one unary slab per proposal, one weight per term.
 * q(:,p) = disparity of the head's plane at the head's point and qprim(:,p) = disparity of the
 * tail's plane at the head's point (dispmap_super.m:180-183) are recomputed inside the sweep kernel
 * from three numbers per label and node (own disparity, disparity step per column / row), which
 * replaces the per-edge q / q' / order storage of typeStereoLinear.h:274-311.  State is 45 bytes per
 * label and node (fp32), so BASELINE config 5 (1980 x 2880 x 192) fits one B200 and config 4
 * (4096 x 4096 x 256) fits eight, each rank holding only its own column band (+ one halo column per
 * inner side).  The bands are COLUMN bands because the sweep's strips run along the image rows
 * (ordering.cpp:7-157 orders the interior row by row, right to left): a rank sweeps its piece of row r
 * while its right-hand neighbour already sweeps row r + 1 -- the ranks are the stages of a pipeline --
 * where row bands would make them take turns (row r + 1 waits for row r).
 *
 *   sb_trws_grid_create      rank `rank` of `world` (world = 1: the whole grid): the columns are cut into blocks that
 *                            are dealt round robin to the ranks (opt.col_blocks per rank); H, W >= 4, W >= 4 x the
 *                            number of blocks; kernel / tol / options as sb_trws_solve
 *   sb_trws_grid_set_labels  proposals l0 .. l0+nl-1: planes = nl consecutive 4 x N arrays ([a; b; c; d0]
 *                            per pixel, MATLAB node order), unary = nl x N (proposal-major);
 *                            d_min / d_step = the disparity normalisation of
 *                            dispmap_globalstereo.m:336-345 (0, 1 for dispmap_ncc / dispmap_super).
 *                            c == 0 -> SB_EINVAL "Infinite disparity" (dispmap_super.m:321-323).
 *                            planes / unary (and alphas below) may be host OR device pointers (unified addressing)
 *   sb_trws_grid_set_weights alphas, E doubles in the reference's term order (trws_mex.cpp:35)
 *   sb_trws_grid_synth       fills every proposal, unary and weight with the seeded synthetic problem of
 *                            SURVEY 8(d) ON the device (for sizes whose inputs do not fit a host)
 *   sb_trws_grid_finalize    builds the sort-rank / merge-count tables (the argsort loop of
 *                            trws_mex.cpp:84-119) on the device; call after the labels are set
 *   sb_trws_grid_get_label   one stored proposal back out: 4 x N doubles [unary | own disparity | step per
 *                            column | step per row] (N each, MATLAB node order; nodes this rank does not
 *                            store are 0) -- inspection / parity tests
 *   sb_trws_grid_get_weights the stored weights, E doubles in the reference's term order
 *   sb_trws_grid_minimize    Minimize_TRW_S (minimize.cpp:7-116), world = 1 only; continues from the
 *                            current messages.  With fuse_rounding and max_relgap > 0 the stop test of
 *                            iteration t runs inside the forward sweep of t + 1, so after an early stop
 *                            the resident messages are half an iteration ahead of the reference's
 *                            (the returned labels / energy / bound are those of iteration t)
 *   sb_trws_grid_get_labels  N doubles, 1-based, MATLAB node order; nodes this rank does not sweep are 0
 *   sb_trws_grid_pass / _ipc_export (2 x 64 bytes: message and selected-position arrays) / _ipc_attach
 *                            (`up` = rank (rank - 1) mod world, which owns the blocks to the left of mine; `down` =
 *                            (rank + 1) mod world; with two ranks both are the same process):
 *                            the multi-GPU protocol of sb_trws_pass, on sharded state: the messages of the
 *                            horizontal terms that cross a band boundary are stored by both ranks and
 *                            written by the sender into both copies (its own, and the neighbour's over
 *                            NVLink); a message word carries the parity of the pass counter in its sign
 *                            bit (min-normalised messages are >= 0), so the receiver polls its own copy --
 *                            no mailbox array, no flag, no fence
 *   sb_trws_grid_launch_pass / _wait: the same pass, asynchronous: launch only enqueues, wait blocks
 *                            until every launched pass has finished and returns (energy, bound
 *                            contribution) per pass in launch order.  Because every message word
 *                            validates itself the ranks need no barrier between passes: with a fixed
 *                            iteration count (max_relgap = 0) a rank launches all its passes at once and
 *                            the per-pass sums are all-reduced once at the end
 *   sb_trws_grid_attach_local neighbours that are solvers of the SAME process on the same device (several
 *                            ranks sharing one GPU: how a one-GPU box runs the banded sweep); `share` =
 *                            solvers that sweep concurrently (each takes 1/share of the resident CTAs)
 *   sb_trws_grid_info        info[0] = bytes of HBM state on this rank, [1] = nodes stored, [2], [3] =
 *                            columns swept [lo, hi), [4], [5] = persistent CTAs (forward, backward), [6] =
 *                            dynamic shared memory per CTA, [7] = padded label count */
typedef struct sb_trws_grid sb_trws_grid;
SB_API int sb_trws_grid_create(int kernel, int H, int W, int L, double tol, const sb_trws_options *opt,
                        int rank, int world, sb_trws_grid **out);
SB_API int sb_trws_grid_set_labels(sb_trws_grid *g, int l0, int nl, const double *planes, const double *unary,
                            double d_min, double d_step);
SB_API int sb_trws_grid_set_weights(sb_trws_grid *g, const double *alphas);
SB_API int sb_trws_grid_synth(sb_trws_grid *g, uint64_t seed);
/* the same generator, the solver's grid being the window at (r_off, c_off) of a scene_H x scene_W scene
 * (bench.py: the crop of the full-size problem that the CPU reference is timed on) */
SB_API int sb_trws_grid_synth_window(sb_trws_grid *g, uint64_t seed, int scene_H, int scene_W, int r_off, int c_off);
SB_API int sb_trws_grid_finalize(sb_trws_grid *g);
SB_API int sb_trws_grid_get_label(sb_trws_grid *g, int l, double *out /* 4 x N */);
SB_API int sb_trws_grid_get_weights(sb_trws_grid *g, double *alphas /* E */);
SB_API int sb_trws_grid_reset(sb_trws_grid *g);
SB_API int sb_trws_grid_minimize(sb_trws_grid *g, double maxiter, double max_relgap,
                          double *energy, double *lower_bound, double *iterations, sb_trws_timing *timing);
SB_API int sb_trws_grid_get_labels(sb_trws_grid *g, double *labels);
SB_API int sb_trws_grid_ipc_export(sb_trws_grid *g, unsigned char *handles /* 128 bytes */);
SB_API int sb_trws_grid_ipc_attach(sb_trws_grid *g, const unsigned char *up, const unsigned char *down);
SB_API int sb_trws_grid_pass(sb_trws_grid *g, int pass, int mode, double *acc /* 2 */);
SB_API int sb_trws_grid_launch_pass(sb_trws_grid *g, int pass, int mode);
SB_API int sb_trws_grid_wait(sb_trws_grid *g, double *acc /* 2 x max_passes, may be null */, int max_passes, int *n_passes);
SB_API int sb_trws_grid_attach_local(sb_trws_grid *g, sb_trws_grid *up, sb_trws_grid *down, int share);
SB_API int sb_trws_grid_info(sb_trws_grid *g, int64_t *info /* 8 */);
/* *on = 1 when this solver sweeps with the latency build (sb_trws_options.latency_mode) */
SB_API int sb_trws_grid_latency_mode(sb_trws_grid *g, int *on);
/* cumulative since creation: out[0] = ms spent in sweep kernels (CUDA events around every launch on the solver
 * stream), out[1] = sweep launches, out[2] = set-up ms (uploads, table build) */
SB_API int sb_trws_grid_counters(sb_trws_grid *g, double *out /* 3 */);
SB_API void sb_trws_grid_destroy(sb_trws_grid *g);
/* Host-only: per pass (forward, backward) six values: strips, segments, node steps, nodes, nodes that
 * need two steps, (messages pushed to rank - 1) * 1e6 + (messages pushed to rank + 1).  12 values. */
SB_API int sb_trws_grid_plan_stats(int H, int W, int rank, int world, int64_t *stats);

/* One Edge::UpdateMessage (typeStereoLinear.h:329-487 / typeStereoQuadratic.h:329-501) run by the sweep
 * kernels' own device routine on one warp -- the unit-level known-answer entry:
 *   msg_out[j] = min(vTrunc, min_i gamma*Di[i] - msg[i] + alpha*|dst_pos[j] - src_pos[i]|^kernel) - vMin.
 * Known deviation: on EXACT ties h_j - h_k == alpha*(q_k - q_j) the reference's cone envelope drops cone k
 * (its `s <= qj -> break` path, typeStereoLinear.h:443-446) and so returns values ABOVE the exact min-plus
 * message; the device computes the exact minimum.  Continuous data never ties; integer-valued unaries with
 * integer positions do (tests/test_update_message_gpu.py quantifies it). */
SB_API int sb_trws_update_message(int kernel, int L, const double *Di, const double *msg,
                           const double *src_pos, const double *dst_pos, double alpha, double lambda,
                           double gamma, int precision, double *msg_out, double *vmin_out);

/* Node ordering of MRFEnergy::SetAutomaticOrdering (cpp/trw-s/ordering.cpp:7-157)
 * on the H x W grid: ordering[r + H*c] in [0, H*W).  Closed form for H,W >= 4
 * (SURVEY Appendix A.1), literal greedy scan otherwise.  Host-only. */
SB_API int sb_trws_grid_ordering(int H, int W, int32_t *ordering);

/* Host-only introspection of the sweep schedule (strips -> segments, trws_order.cpp) for an
 * H x W grid as rank `rank` of `world` row bands sees it (world = 1: the whole grid):
 * stats[0] = strips of the whole schedule; then per pass (forward, backward) six values:
 * segments, nodes covered, rows fetched by the helper warps, nodes that send on more than four
 * terms, messages pushed to rank - 1, messages pushed to rank + 1.  13 values. */
SB_API int sb_trws_plan_stats(int H, int W, int rank, int world, int64_t *stats);

/* Infer (H, W) from a connectivity list and verify it is the reference grid.
 * Host-only.  Returns SB_ENOTGRID when it is not. */
SB_API int sb_grid_from_connectivity(int64_t N, int64_t E, const uint32_t *conn, int *H, int *W);

/* ---------------------------------------------------------------- QPBO / roof duality */

/* Replaces mexFunction of cpp/rd_mex.cpp:14-101 (called from rd.m:21).
 *   U0, U1            N      unary cost of keeping / switching           rd_mex.cpp:24-25
 *   E00,E01,E10,E11   E      pairwise tables, E_ab(p) = cost of (x_tail = a, x_head = b)
 *   conn              2 x E  uint32, 0-based (rd.m:21 subtracts 1)        rd_mex.cpp:31
 *   improve           QPBO-I on the nodes left unlabelled                 rd_mex.cpp:34,91-92
 * Outputs (rd_mex.cpp:72-100): labels N doubles in {0, 1, negative = unlabelled}; energy of
 * the labelling with unlabelled -> 0; roof-dual lower bound; number of nodes unlabelled
 * after Solve + ComputeWeakPersistencies (before Improve).
 * The connectivity must be the dispmap_super grid, else SB_ENOTGRID.  Improve draws its node
 * permutation from libc rand() exactly like QPBO_extra.cpp:13-27.
 * Known deviation: with improve != 0 and unlabelled nodes, `lower_bound` is the roof-dual bound of the
 * problem (what Solve yields); the reference evaluates ComputeTwiceLowerBound (QPBO.cpp:897-917) on the
 * residual graph its forced-label max-flows leave behind (QPBO_extra.cpp:1151-1232), a number that
 * depends on the particular augmenting paths BK took.  Labels, energy and num_unlabelled agree. */
SB_API int sb_rd_solve(int64_t N, int64_t E, const double *U0, const double *U1,
                const double *E00, const double *E01, const double *E10, const double *E11,
                const uint32_t *conn, int improve,
                double *labels, double *energy, double *lower_bound, double *num_unlabelled);

/* dispmap_super.binary_fusion (dispmap_super.m:61-84) as one grid-native call: the four
 * tables of all_pairwise_costs (dispmap_super.m:236-262) are built ON the device from the two plane fields
 * and feed the QPBO build directly, so neither the 4 x E table doubles nor the 2 x E connectivity cross the
 * host boundary (14 N doubles in instead of 23 N + the connectivity).
 *   assignment, proposal  4 x N plane fields ([a; b; c; d0] per pixel, MATLAB node order)
 *   U0, U1                N   unary cost of keeping / of taking the proposal (dispmap_super.m:66-67)
 *   weights               E   smooth weights, the reference's term order
 *   kernel, tol, d_min, d_step   as sb_pairwise_tables
 *   on_device             != 0: every array pointer (inputs and `labels`) is a DEVICE pointer -- a fusion loop
 *                         that keeps its fields resident pays no copies at all
 *   labels / energy / lower_bound / num_unlabelled   as sb_rd_solve
 *   stats (may be null)   [0] push / relabel rounds, [1] exact relabellings, [2] BFS sweeps, [3] ms from the
 *                         graph build to the labels */
SB_API int sb_binary_fusion_grid(int H, int W, int kernel, const double *assignment, const double *proposal,
                          const double *U0, const double *U1, const double *weights, double tol, double d_min,
                          double d_step, int improve, int on_device, double *labels, double *energy,
                          double *lower_bound, double *num_unlabelled, double *stats /* 4 */);

/* dispmap_super.binary_fuse_until_convergence (dispmap_super.m:85-152) as ONE call over device-resident fields
 * (SURVEY 8(f) rank 3): the reference's loop makes one MATLAB round trip per fusion (unary_cost x 2, all_pairwise_costs,
 * rd); here the proposals and their unary costs go to the device once, every move is sb_binary_fusion_grid in
 * device-pointer mode, a kernel adopts the accepted planes, and only the energy crosses the host per fusion.
 *   proposals      n_proposals x (4 x N) plane fields        (proposal_cell, :91)
 *   unaries        n_proposals x N   unary_cost of each proposal field (:66-67; per-pixel functions of the plane at
 *                  that pixel in dispmap_ncc / dispmap_globalstereo, so the fused field's cost is a selection)
 *   assignment     4 x N   in: the current assignment, out: the fused one
 *   unary          N       in: unary_cost(assignment), out: that of the result
 *   maxiter        self.maxiter (:107)
 *   ids, n_ids     the visiting order the CALLER builds at :96-101 (1:n, then 5 maxiter randi draws, repeats removed):
 *                  1-based proposal numbers.  The random stream stays MATLAB's own.
 *   on_device      != 0: proposals, unaries, assignment, unary and weights are device pointers
 *   energies       maxiter + 1 doubles: E of :104 / :128 (energies[0] = energy before the first move)
 *   n_energies     number_of_iterations = length(E) (:151)
 *   stats          (may be null) [0] fusion moves made, [1] push / relabel rounds, [2] solver ms, [3] pixels adopted
 * The loop keeps the reference's bookkeeping: `iter = iter + 1` (:116) starts it at ids(2); a proposal is marked visited
 * when a move leaves E unchanged, all marks are cleared when E changes, and it stops when every proposal is marked. */
SB_API int sb_binary_fuse_until_convergence_grid(int H, int W, int kernel, int n_proposals, const double *proposals,
                          const double *unaries, double *assignment, double *unary, const double *weights, double tol,
                          double d_min, double d_step, int improve, int maxiter, const int32_t *ids, int64_t n_ids,
                          int on_device, double *energies, int *n_energies, double *stats /* 4 */);

/* ------------------------------------------- cost volume / unary / pairwise builders
 *
 * The dense arrays dispmap_super.binary_fusion / simultaneous_fusion hand to rd() / trws()
 * are produced by the methods below in the reference; each entry point replaces one of
 * them.  Images are H x W x C doubles (MATLAB layout, values 0..255), planes are 4 x M
 * ([a; b; c; d0] per column), points 2 x M ([x; y] = [column; row], 1-based).
 * d_min / d_step are the disparity normalisation of dispmap_globalstereo.m:336-345
 * (pass 0 and 1 for dispmap_ncc / dispmap_super). */

/* dispmap_ncc.compute_ncc (dispmap_ncc.m:116-198): NCC volume H x W x D over the joint
 * (2*patchsize+1)^2 x 3 window; the reference hard-codes patchsize = 2 (dispmap_ncc.m:24). */
SB_API int sb_ncc_volume(int H, int W, int C, const double *im0, const double *im1,
                  int D, const double *disparities, int patchsize, double *ncc_out);
/* The same volume kept ON the device (fp32, H x W x D) behind a handle, for the methods that only sample it
 * (dispmap_ncc.m:107-115, 208-276): nothing of size H x W x D crosses the host boundary unless sb_ncc_vol_get asks.
 * For 8-bit images and integer disparities (every case the reference's examples run, example_ncc.m:13-16) the
 * volume comes from ONE pass over the two images for all levels (ncc_volume.cu: exact integer running window
 * sums, image columns staged by TMA tensor-map copies); otherwise from the general per-level kernel.
 *   sb_ncc_vol_best_disp / _sample  = sb_ncc_best_disp / sb_ncc_sample on the resident volume (bit-identical)
 *   sb_ncc_vol_info   info[0] = ms the volume kernels took, [1] = 1 if the one-pass kernel ran, [2] = bytes held */
typedef struct sb_ncc_vol sb_ncc_vol;
SB_API int sb_ncc_vol_create(int H, int W, int C, const double *im0, const double *im1,
                      int D, const double *disparities, int patchsize, sb_ncc_vol **out);
SB_API int sb_ncc_vol_get(sb_ncc_vol *v, double *ncc_out /* H x W x D */);
SB_API int sb_ncc_vol_best_disp(sb_ncc_vol *v, double *best_disp);
SB_API int sb_ncc_vol_sample(sb_ncc_vol *v, const double *disps, double unary_weight, int as_unary, double *out);
SB_API int sb_ncc_vol_info(sb_ncc_vol *v, double *info /* 3 */);
SB_API void sb_ncc_vol_destroy(sb_ncc_vol *v);
/* dispmap_ncc.best_disp_from_ncc (dispmap_ncc.m:208-221): WTA level + parabola refinement. */
SB_API int sb_ncc_best_disp(int H, int W, int D, const double *ncc, const double *disparities,
                     double *best_disp);
/* dispmap_ncc.sample_ncc_from_disp (dispmap_ncc.m:222-245); with as_unary != 0 the output is
 * unary_weight * (1 - nccs), i.e. dispmap_ncc.unary_cost (dispmap_ncc.m:107-115). */
SB_API int sb_ncc_sample(int H, int W, int D, const double *ncc, const double *disparities,
                  const double *disps, double unary_weight, int as_unary, double *out);
/* dispmap_super.disparitymap_from_assignment (dispmap_super.m:318-328) and its override
 * (dispmap_globalstereo.m:336-345); c == 0 -> SB_EINVAL "Infinite disparity". */
SB_API int sb_plane_disparity(int64_t M, const double *planes, const double *points,
                       double d_min, double d_step, double *out);
/* vgg_interp2(A, X, Y, 'linear', oobv) (imrender/vgg/vgg_interp2.cxx:246-322); B is n x col. */
SB_API int sb_interp2_linear(const double *A, int h, int w, int col, const double *X, const double *Y,
                      int64_t n, double oobv, double *B);
/* The window-matching volume of dispmap_globalstereo.segpln (dispmap_globalstereo.m:83-117): for
 * every disparity level the photo cost ephoto(colour of images{a} at the point projected with that disparity - colour
 * of the reference) summed over the images, box mean over the (2 window + 1)^2 window (conv2 'valid'), normalised by
 * X(1) = ephoto(-1000 - R(1, 1, :)) * n_images, first maximum over the levels, matches scoring below min_corr (0.07 in
 * the reference) set to disparity 0, symmetric padding back to H x W.
 *   images   n_images x (H x W x C) doubles, images[0] the reference;  P  n_images x (3 x 4) camera matrices (column-major)
 *   disps    D disparity levels (self.disps);  window = options.window;  col_thresh = options.col_thresh
 *   corr     H x W: the winning disparity per pixel;  score (may be null)  (H - 2 window) x (W - 2 window): info.corr */
SB_API int sb_segpln_wta(int H, int W, int C, int n_images, const double *images, const double *P, int D,
                  const double *disps, int window, double col_thresh, double min_corr, double *corr, double *score);
/* dispmap_ncc.generate_new_plane_RANSAC + fit_plane_to_points (dispmap_ncc.m:48-91; SURVEY 8(f) rank 1, the
 * dispmap_ncc producer of proposal plane fields): the plane through the points [col; row; disp(row, col)] within radius r
 * of (x, y) -- normal = right singular vector of the smallest singular value of the centred point matrix, re-weighted
 * 20 times by sqrt(|residual|) for kernel 1 (IRLS, :76-83), once for kernel 2 --, p(4) = -(p(1:3)' * mean), p / p(3).
 * Computed as the eigenvector of the smallest eigenvalue of the 3 x 3 scatter matrix (fp64 block reductions, Jacobi):
 * agrees with the SVD form to rounding (the sign of V(:, end) cancels in p / p(3)).
 *   disp       H x W doubles (best_disp_from_ncc);  on_device != 0: `disp` and `proposal` are device pointers
 *   plane      4 doubles [a; b; 1; d0];  proposal (may be null) 4 x N: repmat(p, [1 N]) (:65), so that a fusion loop
 *              gets its proposal field without it ever being a host array;  n_points (may be null) points used
 * Fewer than 3 points within the radius: SB_EINVAL. */
SB_API int sb_plane_from_disparity(int H, int W, const double *disp, double x, double y, double r, int kernel,
                            int on_device, double *plane, double *proposal, double *n_points);
/* dispmap_globalstereo.preprocess, the part after the segmentation (dispmap_globalstereo.m:396-401):
 * weights[p] = scale * (lambda_h if the two pixels of term p lie in the same segment else lambda_l), with
 * scale = num_in / ((connect == 8) + 1), terms in dispmap_super.construct_neighborhood order.  segment: H x W uint32
 * labels (vgg_segment_ms output, column-major).  The mean-shift segmentation itself stays the caller's. */
SB_API int sb_smooth_weights(int H, int W, const uint32_t *segment, double lambda_h, double lambda_l, double scale,
                      double *weights);
/* dispmap_globalstereo.unary_cost + ephoto (dispmap_globalstereo.m:355-375,405).
 * P2 = self.P(:,:,2): the 4 x 3 transpose of the second camera matrix (:42). */
SB_API int sb_photo_unary(int H, int W, int C, const double *im0, const double *im1, const double *P2,
                   const double *planes, double d_min, double d_step, double col_thresh, double *U);
/* dispmap_super.all_pairwise_costs (dispmap_super.m:226-262).  proposal == NULL computes E00
 * only (the nargout == 1 form used by update_energy). */
SB_API int sb_pairwise_tables(int H, int W, int kernel, const double *assignment, const double *proposal,
                       const double *weights, double tol, double d_min, double d_step,
                       double *E00, double *E01, double *E10, double *E11);
/* q / qprim (L x E) of dispmap_super.simultaneous_fusion (dispmap_super.m:170-183);
 * proposals = L consecutive 4 x N plane arrays. */
SB_API int sb_fusion_positions(int H, int W, int L, const double *proposals, double d_min, double d_step,
                        double *q, double *qprim);
/* dispmap_super.update_energy (dispmap_super.m:263-274): sum(unary) + sum(E00). */
SB_API int sb_energy(int H, int W, int kernel, const double *unary, const double *assignment,
              const double *weights, double tol, double d_min, double d_step, double *energy);

#ifdef __cplusplus
}
#endif
#endif /* STEREO_B200_H */
