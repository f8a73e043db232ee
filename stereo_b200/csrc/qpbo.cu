// qpbo.cu -- roof duality (QPBO) binary fusion on the GPU behind sb_rd_solve
// (include/stereo_b200.h).  Replaces, for the 4-connected dispmap_super grid,
//   mexFunction of cpp/rd_mex.cpp:14-101
//   QPBO::AddPairwiseTerm / ComputeWeights / AddUnaryTerm   QPBO.cpp:408-507, QPBO.h:615-624,760-807
//   QPBO::MergeParallelEdges                                QPBO.cpp:786-816, QPBO_extra.cpp:137-239
//   QPBO::Solve / maxflow                                   QPBO.cpp:818-845, QPBO_maxflow.cpp:477-617
//   QPBO::ComputeWeakPersistencies                          QPBO_postprocessing.cpp:10-120
//   QPBO::Improve                                           QPBO_extra.cpp:1151-1232
//   ComputeTwiceEnergy / ComputeTwiceLowerBound             QPBO.cpp:847-917
//
// Design.  The reference runs Boykov-Kolmogorov augmenting paths on a pointer graph; the
// quantities it returns, however, are properties of the energy, not of the algorithm:
//   * strong labels (Solve): label(i) = what_segment(i) unless it equals what_segment(i'),
//     where what_segment(v) = 1 iff v is in the sink tree, i.e. iff v can still reach the
//     sink in the residual graph of a maximum flow -- the same set for every maximum flow;
//   * the roof-dual bound = constant + value of the maximum flow.
// So the device runs a different max-flow: synchronous (Jacobi) push-relabel on the implicit
// 2N-node grid network, fp64 capacities, one thread per node, periodic exact global relabelling
// by parallel BFS from the sink.  Phase 1 (maximum preflow) already fixes the sink-reachable
// set; the excess stranded on the source side is only returned (phase 2) when weak
// persistencies have to be computed on a true flow.  The two directed terms of every
// neighbour pair are summed into one table before the normal form (the reference merges them
// afterwards, QPBO.cpp:786-816; the represented energy is the same).
//
// Layout: node v = u + side * N (side 1 = the mate i'), u = r + H*c.  Pair P: vertical pairs
// (r,c)-(r+1,c) first, P = c*(H-1)+r, then horizontal pairs (r,c)-(r,c+1), P = nV + c*H + r;
// i = first node of the pair.  Four residual capacities per pair:
//   [0] i -> J0   [1] J0 -> i   [2] J1 -> i'   [3] i' -> J1
// with (J0, J1) = (j, j') for a submodular pair and (j', j) otherwise (QPBO.cpp:458-500).
#include "sb_common.h"
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <chrono>

namespace sb {
namespace qpbo {

constexpr int HINF = 0x3fffffff;

struct Graph {
    int H, W;
    long long N, nV, nH, nP;
    double *tr0;      // [2N] initial signed terminal capacity (> 0: from the source, < 0: to the sink)
    double *excess;   // [2N]
    double *tsink;    // [2N] residual capacity v -> sink
    double *tsrc;     // [2N] residual capacity v -> source (flow received from the source)
    double *r;        // [nP][4]
    double *pushed;   // [nP][4] flow sent over each arc in the current round
    unsigned char *sub; // [nP]
    int *h, *h2;      // [2N] labels (double buffered)
    int *flag;        // device flags: [0] active, [1] changed
};

// ComputeWeights, QPBO.h:760-807
__device__ __forceinline__ void compute_weights(double A, double B, double C, double D, double &ci, double &cj,
                                                double &cij, double &cji)
{
    ci = D - A;
    B -= A;
    C -= D;
    if (B < 0) {
        ci += -B;
        cj = B;
        cji = B + C;
        cij = 0;
    } else if (C < 0) {
        ci += C;
        cj = -C;
        cij = B + C;
        cji = 0;
    } else {
        cj = 0;
        cij = B;
        cji = C;
    }
}

// One thread per neighbour pair: sum the two directed terms, normal form, arc capacities.
__global__ void build_pairs_kernel(Graph g, const double *__restrict__ E00, const double *__restrict__ E01,
                                   const double *__restrict__ E10, const double *__restrict__ E11,
                                   double *__restrict__ pair_ci, double *__restrict__ pair_cj,
                                   double *__restrict__ pair_t00)
{
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= g.nP) return;
    long long p1, p2;
    if (P < g.nV) { p1 = P; p2 = g.nV + P; }                 // VD, VU (dispmap_super.m:284-288)
    else { p1 = 2 * g.nV + (P - g.nV); p2 = p1 + g.nH; }     // HR, HL (:290-294)
    const double t00 = E00[p1] + E00[p2], t01 = E01[p1] + E10[p2], t10 = E10[p1] + E01[p2], t11 = E11[p1] + E11[p2];
    double ci, cj, cij, cji;
    const bool sub = (t01 + t10 >= t00 + t11);               // QPBO.cpp:433
    if (sub) compute_weights(t00, t01, t10, t11, ci, cj, cij, cji);
    else { compute_weights(t01, t00, t11, t10, ci, cj, cij, cji); cj = -cj; }
    g.sub[P] = sub ? 1 : 0;
    double *r = g.r + 4 * P;
    r[0] = cij; r[1] = cji; r[2] = cij; r[3] = cji;
    pair_ci[P] = ci;
    pair_cj[P] = cj;
    pair_t00[P] = t00;
}

// One thread per pixel: terminal capacity = unary difference + the pairs' contributions, in
// a fixed order (up, down, left, right).
__global__ void build_nodes_kernel(Graph g, const double *__restrict__ U0, const double *__restrict__ U1,
                                   const double *__restrict__ pair_ci, const double *__restrict__ pair_cj)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= g.N) return;
    const int H = g.H, W = g.W;
    const int r = (int)(u % H), c = (int)(u / H);
    double t = U1[u] - U0[u];
    if (r > 0) t += pair_cj[(long long)c * (H - 1) + (r - 1)];
    if (r < H - 1) t += pair_ci[(long long)c * (H - 1) + r];
    if (c > 0) t += pair_cj[g.nV + (long long)(c - 1) * H + r];
    if (c < W - 1) t += pair_ci[g.nV + (long long)c * H + r];
    g.tr0[u] = t;
    g.tr0[u + g.N] = -t;
}

__global__ void init_preflow_kernel(Graph g)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * g.N) return;
    const double t = g.tr0[v];
    g.excess[v] = t > 0 ? t : 0.0;
    g.tsink[v] = t < 0 ? -t : 0.0;
    g.tsrc[v] = t > 0 ? t : 0.0;
    g.h[v] = 0;
    g.h2[v] = 0;
}

// Arc of node v in grid direction d (0 up, 1 down, 2 left, 3 right): pair, this node's
// outgoing slot (the incoming one is slot ^ 1) and the node at the other end.
struct Arc { long long P; int slot; long long w; bool valid; };
__device__ __forceinline__ Arc arc_of(const Graph &g, long long v, int d)
{
    Arc a;
    const long long N = g.N;
    const int side = v >= N ? 1 : 0;
    const long long u = v - (side ? N : 0);
    const int H = g.H, W = g.W;
    const int r = (int)(u % H), c = (int)(u / H);
    bool first;
    long long other;
    if (d == 0) { a.valid = r > 0; a.P = (long long)c * (H - 1) + (r - 1); first = false; other = u - 1; }
    else if (d == 1) { a.valid = r < H - 1; a.P = (long long)c * (H - 1) + r; first = true; other = u + 1; }
    else if (d == 2) { a.valid = c > 0; a.P = g.nV + (long long)(c - 1) * H + r; first = false; other = u - H; }
    else { a.valid = c < W - 1; a.P = g.nV + (long long)c * H + r; first = true; other = u + H; }
    if (!a.valid) { a.slot = 0; a.w = 0; return a; }
    const bool sub = g.sub[a.P] != 0;
    int oside;
    if (first) {           // i / i'
        a.slot = side ? 3 : 0;
        oside = side ? (sub ? 1 : 0) : (sub ? 0 : 1);
    } else {               // j / j'
        const bool isJ0 = (side == 0) == sub;   // J0 = j if sub else j'
        a.slot = isJ0 ? 1 : 2;
        oside = isJ0 ? 0 : 1;
    }
    a.w = other + (oside ? N : 0);
    return a;
}

// Push round: every active node sends along admissible arcs (label exactly one lower).
// TO_SOURCE selects phase 2 (the stranded excess flows back to the source).
template <bool TO_SOURCE> __global__ void push_kernel(Graph g)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * g.N) return;
    double e = g.excess[v];
    const int hv = g.h[v];
    if (!(e > 0) || hv >= HINF) return;
    double *term = TO_SOURCE ? g.tsrc : g.tsink;
    const double tc = term[v];
    if (tc > 0 && hv == 1) {
        const double dl = fmin(e, tc);
        term[v] = tc - dl;
        e -= dl;
    }
#pragma unroll
    for (int d = 0; d < 4; d++) {
        if (!(e > 0)) break;
        const Arc a = arc_of(g, v, d);
        if (!a.valid) continue;
        const double rr = g.r[4 * a.P + a.slot];
        if (rr > 0 && g.h[a.w] == hv - 1) {
            const double dl = fmin(e, rr);
            g.r[4 * a.P + a.slot] = rr - dl;
            e -= dl;
            g.pushed[4 * a.P + a.slot] = dl;
        }
    }
    g.excess[v] = e;
}

// Collect what the neighbours pushed, then relabel (Jacobi: reads h, writes h2).
template <bool TO_SOURCE> __global__ void collect_relabel_kernel(Graph g)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * g.N) return;
    double e = g.excess[v];
    int minh = HINF;
    const double tc = (TO_SOURCE ? g.tsrc : g.tsink)[v];
    if (tc > 0) minh = 0;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const Arc a = arc_of(g, v, d);
        if (!a.valid) continue;
        const long long in = 4 * a.P + (a.slot ^ 1);
        const double dl = g.pushed[in];
        double rr = g.r[4 * a.P + a.slot];
        if (dl != 0.0) {
            e += dl;
            rr += dl;
            g.r[4 * a.P + a.slot] = rr;
            g.pushed[in] = 0.0;
        }
        if (rr > 0) minh = min(minh, g.h[a.w]);
    }
    g.excess[v] = e;
    int hv = g.h[v];
    if (e > 0 && hv < HINF) {
        const int hn = minh >= HINF ? HINF : minh + 1;
        if (hn > hv) hv = hn;
        if (hv < HINF) g.flag[0] = 1;
    }
    g.h2[v] = hv;
}

// Exact labels: BFS distance to the terminal in the residual graph (relaxation sweeps).
template <bool TO_SOURCE> __global__ void bfs_init_kernel(Graph g)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * g.N) return;
    g.h[v] = ((TO_SOURCE ? g.tsrc : g.tsink)[v] > 0) ? 1 : HINF;
}
__global__ void bfs_sweep_kernel(Graph g)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * g.N) return;
    int hv = g.h[v];
    int best = hv;
#pragma unroll
    for (int d = 0; d < 4; d++) {
        const Arc a = arc_of(g, v, d);
        if (!a.valid) continue;
        if (g.r[4 * a.P + a.slot] > 0) {
            const int hw = g.h[a.w];
            if (hw < HINF && hw + 1 < best) best = hw + 1;
        }
    }
    if (best < hv) {
        g.h[v] = best;
        g.flag[1] = 1;
    }
}
// The same relaxation, run to a LOCAL fixed point on a 32 x 8 tile (both node copies) in shared memory before anything
// is written back: one launch carries a distance across a whole tile instead of one node, so the number of global
// sweeps drops from the residual graph's diameter to roughly diameter / tile size (long thin paths on 1980 x 2880
// grids needed thousands of single-step sweeps).  Halo labels are read once and stay fixed during the launch; labels
// only ever decrease towards the BFS distances, so any schedule reaches the same fixed point.
constexpr int BT_R = 32, BT_C = 8;
__global__ void __launch_bounds__(BT_R * BT_C) bfs_tile_kernel(Graph g)
{
    __shared__ int hs[2][BT_C + 2][BT_R + 2];
    const int H = g.H, W = g.W;
    const long long N = g.N;
    const int r0 = (int)blockIdx.x * BT_R, c0 = (int)blockIdx.y * BT_C;
    for (int i = threadIdx.x; i < 2 * (BT_C + 2) * (BT_R + 2); i += blockDim.x) {
        const int lr = i % (BT_R + 2), lc = (i / (BT_R + 2)) % (BT_C + 2), sd = i / ((BT_R + 2) * (BT_C + 2));
        const int r = r0 + lr - 1, c = c0 + lc - 1;
        hs[sd][lc][lr] = (r >= 0 && r < H && c >= 0 && c < W) ? g.h[(long long)r + (long long)H * c + (sd ? N : 0)] : HINF;
    }
    const int lr = threadIdx.x % BT_R, lc = threadIdx.x / BT_R;
    const int r = r0 + lr, c = c0 + lc;
    const bool inside = r < H && c < W;
    const long long u = (long long)r + (long long)H * c;
    // the residual arcs out of my two nodes: the shared-memory cell of the head, or null
    int *nb[2][4];
    int hv[2] = {HINF, HINF};
#pragma unroll
    for (int sd = 0; sd < 2; sd++)
#pragma unroll
        for (int d = 0; d < 4; d++) {
            nb[sd][d] = nullptr;
            if (!inside) continue;
            const Arc a = arc_of(g, u + (sd ? N : 0), d);
            if (!a.valid || !(g.r[4 * a.P + a.slot] > 0)) continue;
            const int os = a.w >= N ? 1 : 0;
            const int dr = d == 0 ? -1 : d == 1 ? 1 : 0, dc = d == 2 ? -1 : d == 3 ? 1 : 0;
            nb[sd][d] = &hs[os][lc + 1 + dc][lr + 1 + dr];
        }
    __syncthreads();
    if (inside) { hv[0] = hs[0][lc + 1][lr + 1]; hv[1] = hs[1][lc + 1][lr + 1]; }
    bool any = false;
    for (int it = 0; it < 2 * (BT_R + BT_C); it++) {
        int best[2] = {hv[0], hv[1]};
#pragma unroll
        for (int sd = 0; sd < 2; sd++)
#pragma unroll
            for (int d = 0; d < 4; d++)
                if (nb[sd][d]) {
                    const int hw = *(volatile int *)nb[sd][d];
                    if (hw < HINF && hw + 1 < best[sd]) best[sd] = hw + 1;
                }
        __syncthreads();
        bool ch = false;
#pragma unroll
        for (int sd = 0; sd < 2; sd++)
            if (best[sd] < hv[sd]) { hv[sd] = best[sd]; hs[sd][lc + 1][lr + 1] = best[sd]; ch = true; }
        any = any || ch;
        if (!__syncthreads_or(ch ? 1 : 0)) break;
    }
    if (any) {
        g.h[u] = hv[0];
        g.h[u + N] = hv[1];
        g.flag[1] = 1;
    }
}
__global__ void any_active_kernel(Graph g)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= 2 * g.N) return;
    if (g.excess[v] > 0 && g.h[v] < HINF) g.flag[0] = 1;
}

// label(i) = what_segment(i), -1 when it equals what_segment(i') (QPBO.cpp:839-843);
// what_segment = 1 iff the node can reach the sink.
__global__ void labels_kernel(Graph g, int *__restrict__ label)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= g.N) return;
    const int a = g.h[u] < HINF ? 1 : 0, b = g.h[u + g.N] < HINF ? 1 : 0;
    label[u] = (a == b) ? -1 : a;
}

// Weak persistencies of the trivial kind: an unlabelled node whose arcs (of i and of its mate i')
// all have zero residual in both directions is a singleton component in both depth-first passes
// of ComputeWeakPersistencies; the reference visits all i before all i' (QPBO_postprocessing.cpp:
// 39-42), so i' finishes later, gets the smaller region number and i is labelled 0 (:113).
// Exact input ties (proposal == current plane) produce exactly these nodes.  Everything else
// that is still open is counted for the general (host) path.
__global__ void isolated_open_kernel(Graph g, int *__restrict__ label, int *__restrict__ complex_open)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= g.N || label[u] >= 0) return;
    bool isolated = true;
    for (int s = 0; s < 2 && isolated; s++)
        for (int d = 0; d < 4; d++) {
            const Arc a = arc_of(g, u + (s ? g.N : 0), d);
            if (!a.valid) continue;
            if (g.r[4 * a.P + a.slot] != 0.0 || g.r[4 * a.P + (a.slot ^ 1)] != 0.0) { isolated = false; break; }
        }
    if (isolated) label[u] = 0;
    else atomicAdd(complex_open, 1);
}

// energy of a labelling (unlabelled -> 0), ComputeTwiceEnergy / 2 (QPBO.cpp:847-875)
__global__ void energy_terms_kernel(long long N, long long E, const unsigned *__restrict__ conn,
                                    const int *__restrict__ label, const double *__restrict__ U0,
                                    const double *__restrict__ U1, const double *__restrict__ E00,
                                    const double *__restrict__ E01, const double *__restrict__ E10,
                                    const double *__restrict__ E11, double *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N) {
        out[t] = label[t] == 1 ? U1[t] : U0[t];
    } else if (t < N + E) {
        const long long p = t - N;
        const int xi = label[conn[2 * p]] == 1, xj = label[conn[2 * p + 1]] == 1;
        out[t] = xi ? (xj ? E11[p] : E10[p]) : (xj ? E01[p] : E00[p]);
    }
}

__global__ void bound_terms_kernel(Graph g, const double *__restrict__ U0, const double *__restrict__ pair_t00,
                                   double *__restrict__ out)
{
    // twice the initial bound: 2 Z + sum_i min(0, 2 t_i) - sum_nonsub 2 cij  (QPBO.cpp:897-917 on the
    // initial reparameterisation), laid out as [N node terms][nP pair terms]
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < g.N) {
        out[t] = 2.0 * U0[t] + fmin(0.0, 2.0 * g.tr0[t]);
    } else if (t < g.N + g.nP) {
        const long long P = t - g.N;
        out[t] = 2.0 * pair_t00[P];
    }
}
__global__ void flow_terms_kernel(Graph g, const double *__restrict__ r0, double *__restrict__ out)
{
    // flow into the sink per node; the non-submodular E00 caps are subtracted from the pair terms
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 2 * g.N) {
        const double t0 = g.tr0[t];
        out[t] = (t0 < 0 ? -t0 : 0.0) - g.tsink[t];
    } else if (t < 2 * g.N + g.nP) {
        const long long P = t - 2 * g.N;
        out[t] = g.sub[P] ? 0.0 : -2.0 * r0[4 * P];
    }
}

__global__ void reduce_kernel(const double *__restrict__ a, long long n, double *__restrict__ partial)
{
    __shared__ double sh[256];
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += a[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// add a unary term (0, delta) to node u on the device state (AddUnaryTerm in stage 1,
// QPBO.h:615-624): tr(u) += delta, tr(u') -= delta, applied to the preflow state
__global__ void add_unary_kernel(Graph g, long long u, double delta)
{
    if (blockIdx.x || threadIdx.x) return;
    for (int s = 0; s < 2; s++) {
        const long long v = u + (s ? g.N : 0);
        double dl = s ? -delta : delta;
        if (dl > 0) { // more source capacity: cancels sink residual first, the rest is new excess
            const double c = fmin(dl, g.tsink[v]);
            g.tsink[v] -= c;
            g.excess[v] += dl - c;
            g.tsrc[v] += dl - c;
        } else if (dl < 0) { // more sink capacity: absorbs the node's excess first
            dl = -dl;
            const double c = fmin(dl, g.excess[v]);
            g.excess[v] -= c;
            g.tsink[v] += dl - c;
        }
    }
}
// DetermineSaturation(i) + 1 (QPBO_extra.cpp:242-254, 1178-1181) on the current residuals
__global__ void saturation_kernel(Graph g, long long u, double *out)
{
    if (blockIdx.x || threadIdx.x) return;
    // any value that exceeds everything the node's arcs and terminals can carry works; the
    // reference takes max(c1, c2) + 1 over its own residuals
    double c = g.excess[u] + g.tsink[u] + 1.0;
    for (int d = 0; d < 4; d++) {
        const Arc a = arc_of(g, u, d);
        if (!a.valid) continue;
        c += g.r[4 * a.P + a.slot] + g.r[4 * a.P + (a.slot ^ 1)];
    }
    *out = c;
}

inline unsigned nblk(long long n) { return (unsigned)std::max<long long>(1, (n + 255) / 256); }

struct Solver {
    Graph g;
    DevBuf<double> tr0, excess, tsink, tsrc, r, pushed, r0;
    DevBuf<unsigned char> sub;
    DevBuf<int> h, h2, flag, label;
    int *hflag = nullptr;
    int64_t rounds = 0, relabels = 0, bfs_sweeps = 0;

    ~Solver() { if (hflag) cudaFreeHost(hflag); }

    void alloc(int H, int W)
    {
        g.H = H; g.W = W; g.N = (long long)H * W;
        g.nV = (long long)(H - 1) * W; g.nH = (long long)H * (W - 1); g.nP = g.nV + g.nH;
        const size_t n2 = (size_t)2 * g.N, np4 = (size_t)std::max<long long>(g.nP, 1) * 4;
        tr0.alloc(n2); excess.alloc(n2); tsink.alloc(n2); tsrc.alloc(n2); r.alloc(np4); pushed.alloc(np4); r0.alloc(np4);
        sub.alloc((size_t)std::max<long long>(g.nP, 1)); h.alloc(n2); h2.alloc(n2); flag.alloc(2); label.alloc((size_t)g.N);
        g.tr0 = tr0.p; g.excess = excess.p; g.tsink = tsink.p; g.tsrc = tsrc.p; g.r = r.p; g.pushed = pushed.p;
        g.sub = sub.p; g.h = h.p; g.h2 = h2.p; g.flag = flag.p;
        SB_CUDA(cudaMallocHost((void **)&hflag, 2 * sizeof(int)));
        SB_CUDA(cudaMemset(pushed.p, 0, pushed.bytes()));
    }

    // exact labels by BFS from the terminal; returns true if some node with excess can reach it
    template <bool TO_SOURCE> bool global_relabel()
    {
        const unsigned nb = nblk(2 * g.N);
        bfs_init_kernel<TO_SOURCE><<<nb, 256>>>(g);
        count_launch();
        for (;;) {
            SB_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int)));
            if (getenv("SB_QPBO_PLAIN_BFS")) {
                for (int k = 0; k < 16; k++) bfs_sweep_kernel<<<nb, 256>>>(g);
                count_launch(16);
                bfs_sweeps += 16;
            } else {
                const dim3 tg((g.H + BT_R - 1) / BT_R, (g.W + BT_C - 1) / BT_C);
                for (int k = 0; k < 4; k++) bfs_tile_kernel<<<tg, BT_R * BT_C>>>(g);
                count_launch(4);
                bfs_sweeps += 4;
            }
            SB_CUDA(cudaMemcpy(hflag, flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
            if (!hflag[1]) break;
        }
        relabels++;
        SB_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int)));
        any_active_kernel<<<nb, 256>>>(g);
        count_launch();
        SB_CUDA(cudaMemcpy(hflag, flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
        return hflag[0] != 0;
    }

    // maximum preflow (phase 1) or return of the stranded excess (phase 2)
    template <bool TO_SOURCE> void maxflow()
    {
        const unsigned nb = nblk(2 * g.N);
        while (global_relabel<TO_SOURCE>()) {
            // a burst of push / relabel rounds between two exact relabellings; the activity flag is
            // read back every 16 rounds so an exhausted preflow ends the burst early.  Bursts stay
            // short: nodes cut off from the terminal only leave the active set at an exact relabelling.
            for (int chunk = 0; chunk < 4; chunk++) {
                SB_CUDA(cudaMemsetAsync(flag.p, 0, 2 * sizeof(int)));
                for (int k = 0; k < 16; k++) {
                    push_kernel<TO_SOURCE><<<nb, 256>>>(g);
                    collect_relabel_kernel<TO_SOURCE><<<nb, 256>>>(g);
                    std::swap(g.h, g.h2);
                }
                count_launch(32);
                rounds += 16;
                SB_CUDA(cudaMemcpy(hflag, flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost));
                if (!hflag[0]) break;
            }
        }
    }
};

double device_sum(const double *d, long long n)
{
    const int nbk = 148 * 4;
    DevBuf<double> part(nbk);
    reduce_kernel<<<nbk, 256>>>(d, n, part.p);
    SB_CUDA(cudaGetLastError());
    count_launch();
    std::vector<double> hh(nbk);
    SB_CUDA(cudaMemcpy(hh.data(), part.p, nbk * sizeof(double), cudaMemcpyDeviceToHost));
    double s = 0;
    for (double v : hh) s += v;
    return s;
}

// ComputeWeakPersistencies (QPBO_postprocessing.cpp:10-120) on the host for the nodes Solve
// left unlabelled: Kosaraju's two depth-first passes over the residual graph restricted to
// them, start nodes and region numbering in the reference's order (all i, then all i').
void weak_persistencies(const Graph &g, const std::vector<double> &r, const std::vector<unsigned char> &sub,
                        std::vector<int> &label)
{
    const long long N = g.N;
    const int H = g.H, W = g.W;
    auto arc_of_host = [&](long long v, int d, long long &P, int &slot, long long &w) -> bool {
        const int side = v >= N ? 1 : 0;
        const long long u = v - (side ? N : 0);
        const int rr = (int)(u % H), c = (int)(u / H);
        bool first;
        long long other;
        if (d == 0) { if (rr <= 0) return false; P = (long long)c * (H - 1) + (rr - 1); first = false; other = u - 1; }
        else if (d == 1) { if (rr >= H - 1) return false; P = (long long)c * (H - 1) + rr; first = true; other = u + 1; }
        else if (d == 2) { if (c <= 0) return false; P = g.nV + (long long)(c - 1) * H + rr; first = false; other = u - H; }
        else { if (c >= W - 1) return false; P = g.nV + (long long)c * H + rr; first = true; other = u + H; }
        const bool sb_ = sub[P] != 0;
        int oside;
        if (first) { slot = side ? 3 : 0; oside = side ? (sb_ ? 1 : 0) : (sb_ ? 0 : 1); }
        else { const bool isJ0 = (side == 0) == sb_; slot = isJ0 ? 1 : 2; oside = isJ0 ? 0 : 1; }
        w = other + (oside ? N : 0);
        return true;
    };
    std::vector<int> region((size_t)2 * N, 0);
    std::vector<char> visited((size_t)2 * N, 1);
    bool any = false;
    for (long long u = 0; u < N; u++)
        if (label[u] < 0) { visited[u] = visited[u + N] = 0; region[u] = region[u + N] = -1; any = true; }
    if (!any) return;
    // first DFS: finishing order
    std::vector<long long> order;
    std::vector<std::pair<long long, int>> st;
    for (long long s0 = 0; s0 < 2 * N; s0++) {
        if (visited[s0]) continue;
        visited[s0] = 1;
        st.push_back({s0, 0});
        while (!st.empty()) {
            auto &top = st.back();
            if (top.second >= 4) { order.push_back(top.first); st.pop_back(); continue; }
            const int d = top.second++;
            long long P, w; int slot;
            if (!arc_of_host(top.first, d, P, slot, w)) continue;
            if (!(r[4 * P + slot] > 0) || visited[w]) continue;
            visited[w] = 1;
            st.push_back({w, 0});
        }
    }
    // second DFS over reversed arcs, most recently finished first
    int component = 0;
    for (long long k = (long long)order.size() - 1; k >= 0; k--) {
        const long long s0 = order[k];
        if (region[s0] > 0) continue;
        region[s0] = ++component;
        st.push_back({s0, 0});
        while (!st.empty()) {
            auto &top = st.back();
            if (top.second >= 4) { st.pop_back(); continue; }
            const int d = top.second++;
            long long P, w; int slot;
            if (!arc_of_host(top.first, d, P, slot, w)) continue;
            if (!(r[4 * P + (slot ^ 1)] > 0) || region[w] >= 0) continue;
            region[w] = component;
            st.push_back({w, 0});
        }
    }
    for (long long u = 0; u < N; u++)
        if (label[u] < 0) {
            if (region[u] > region[u + N]) label[u] = 0;
            else if (region[u] < region[u + N]) label[u] = 1;
        }
}


// The four pairwise tables and the unaries of one fusion, resident on the device
struct DeviceProblem {
    int H, W;
    int64_t N, E;
    const double *U0, *U1;
    const double *Et[4];     // E00, E01, E10, E11 in the reference's term order
    const unsigned *conn;    // 2 x E, 0-based
};

// connectivity of dispmap_super.construct_neighborhood (dispmap_super.m:284-294), 0-based
__global__ void grid_conn_kernel(int H, int W, long long E, unsigned *__restrict__ conn)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    const long long nV = (long long)(H - 1) * W, nH = (long long)H * (W - 1);
    long long a, b;
    if (p < 2 * nV) {
        const long long e = p < nV ? p : p - nV;
        const long long c = e / (H - 1), r = e % (H - 1);
        const long long up = r + (long long)H * c;
        if (p < nV) { a = up; b = up + 1; } else { a = up + 1; b = up; }
    } else {
        const long long q = p - 2 * nV;
        const long long e = q < nH ? q : q - nH;
        if (q < nH) { a = e; b = e + H; } else { a = e + H; b = e; }
    }
    conn[2 * p] = (unsigned)a;
    conn[2 * p + 1] = (unsigned)b;
}

__global__ void labels_to_double_kernel(const int *__restrict__ lab, long long N, double *__restrict__ out)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < N) out[u] = (double)lab[u];
}

// rd_mex.cpp:55-100 from device-resident inputs.  lab: final labels on the host (always filled);
// dlabels: optional DEVICE array of N doubles that receives them as well; stats (optional, 4 doubles):
// push/relabel rounds, exact relabellings, BFS sweeps, milliseconds from the graph build to the labels.
void solve_device(const DeviceProblem &dp, int improve, std::vector<int> &lab, double *dlabels, double *energy,
                  double *lower_bound, double *num_unlabelled, double *stats)
{
    const int H = dp.H, W = dp.W;
    const int64_t N = dp.N, E = dp.E;
    const double *dU0 = dp.U0, *dU1 = dp.U1;
    Solver S;
    S.alloc(H, W);
    Graph &g = S.g;
    DevBuf<double> pci((size_t)std::max<long long>(g.nP, 1)), pcj((size_t)std::max<long long>(g.nP, 1)),
        pt00((size_t)std::max<long long>(g.nP, 1));
    const bool prof = getenv("SB_QPBO_PROFILE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_a = now();
    const double t_begin = t_a;
    if (g.nP) build_pairs_kernel<<<nblk(g.nP), 256>>>(g, dp.Et[0], dp.Et[1], dp.Et[2], dp.Et[3], pci.p, pcj.p, pt00.p);
    build_nodes_kernel<<<nblk(N), 256>>>(g, dU0, dU1, pci.p, pcj.p);
    init_preflow_kernel<<<nblk(2 * N), 256>>>(g);
    SB_CUDA(cudaGetLastError());
    count_launch(3);
    SB_CUDA(cudaMemcpy(S.r0.p, S.r.p, S.r.bytes(), cudaMemcpyDeviceToDevice));

    if (prof) { cudaDeviceSynchronize(); fprintf(stderr, "[sb qpbo] build %.1f ms\n", now() - t_a); t_a = now(); }
    // ---- Solve(): maximum preflow, sink-reachable set, strong labels
    S.maxflow<false>();
    if (prof) { cudaDeviceSynchronize(); fprintf(stderr, "[sb qpbo] maxflow %.1f ms\n", now() - t_a); t_a = now(); }
    labels_kernel<<<nblk(N), 256>>>(g, S.label.p);
    count_launch();
    lab.resize((size_t)N);
    SB_CUDA(cudaMemcpy(lab.data(), S.label.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    // open nodes that are trivially weakly persistent are settled on the device
    int complex_open = 0;
    {
        SB_CUDA(cudaMemsetAsync(S.flag.p, 0, 2 * sizeof(int)));
        isolated_open_kernel<<<nblk(N), 256>>>(g, S.label.p, S.flag.p);
        count_launch();
        SB_CUDA(cudaMemcpy(&complex_open, S.flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    }

    // ---- lower bound: (initial bound + flow value) / 2 (ComputeTwiceLowerBound, QPBO.cpp:897-917)
    {
        DevBuf<double> terms((size_t)(2 * N + g.nP));
        bound_terms_kernel<<<nblk(N + g.nP), 256>>>(g, dU0, pt00.p, terms.p);
        count_launch();
        double twice = device_sum(terms.p, N + g.nP);
        flow_terms_kernel<<<nblk(2 * N + g.nP), 256>>>(g, S.r0.p, terms.p);
        count_launch();
        twice += device_sum(terms.p, 2 * N + g.nP);
        *lower_bound = twice / 2;
    }

    // ---- ComputeWeakPersistencies() on the nodes Solve left open
    if (complex_open == 0) {
        SB_CUDA(cudaMemcpy(lab.data(), S.label.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    } else {
        S.maxflow<true>();   // a true flow: stranded excess back to the source
        std::vector<double> hr((size_t)std::max<long long>(g.nP, 1) * 4);
        std::vector<unsigned char> hsub((size_t)std::max<long long>(g.nP, 1));
        SB_CUDA(cudaMemcpy(hr.data(), S.r.p, hr.size() * 8, cudaMemcpyDeviceToHost));
        SB_CUDA(cudaMemcpy(hsub.data(), S.sub.p, hsub.size(), cudaMemcpyDeviceToHost));
        weak_persistencies(g, hr, hsub, lab);
    }
    double nun = 0;
    for (int64_t u = 0; u < N; u++) nun += lab[u] < 0;   // rd_mex.cpp:83-88
    *num_unlabelled = nun;

    // ---- Improve() (QPBO_extra.cpp:1151-1232): fix still-ambiguous nodes one by one to
    // label 0 in a libc rand() permutation, re-solving after each
    if (improve && nun > 0) {
        std::vector<int> perm((size_t)N);
        for (int64_t i = 0; i < N; i++) perm[i] = (int)i;
        for (int64_t i = 0; i + 1 < N; i++) {   // ComputeRandomPermutation, QPBO_extra.cpp:13-27
            int64_t j = i + (int64_t)((rand() / (1.0 + (double)RAND_MAX)) * (double)(N - i));
            if (j > N - 1) j = N - 1;
            std::swap(perm[j], perm[i]);
        }
        S.maxflow<false>();
        labels_kernel<<<nblk(N), 256>>>(g, S.label.p);
        std::vector<int> cur((size_t)N);
        SB_CUDA(cudaMemcpy(cur.data(), S.label.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
        DevBuf<double> dsat(1);
        for (int64_t pi = 0; pi < N; pi++) {
            const int u = perm[pi];
            if (cur[u] >= 0) continue;     // what_segment(i) != what_segment(i'): decided
            double sat = 0;
            saturation_kernel<<<1, 1>>>(g, u, dsat.p);
            SB_CUDA(cudaMemcpy(&sat, dsat.p, 8, cudaMemcpyDeviceToHost));
            add_unary_kernel<<<1, 1>>>(g, u, sat);   // user_label == 0: forbid label 1
            count_launch(2);
            S.maxflow<false>();
            labels_kernel<<<nblk(N), 256>>>(g, S.label.p);
            count_launch();
            SB_CUDA(cudaMemcpy(cur.data(), S.label.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
        }
        for (int64_t u = 0; u < N; u++) lab[u] = cur[u] < 0 ? 0 : cur[u];   // ambiguous -> user_label (0)
    }

    if (prof) fprintf(stderr, "[sb qpbo] labels+bound+weak %.1f ms\n", now() - t_a);
    if (prof)
        fprintf(stderr, "[sb qpbo] %dx%d: %lld push/relabel rounds, %lld global relabels, %lld bfs sweeps\n", H, W,
                (long long)S.rounds, (long long)S.relabels, (long long)S.bfs_sweeps);
    if (stats) {
        stats[0] = (double)S.rounds; stats[1] = (double)S.relabels; stats[2] = (double)S.bfs_sweeps;
        stats[3] = now() - t_begin;
    }
    // ---- energy of the labelling (unlabelled -> 0), ComputeTwiceEnergy / 2
    {
        SB_CUDA(cudaMemcpy(S.label.p, lab.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
        DevBuf<double> terms((size_t)(N + E));
        energy_terms_kernel<<<nblk(N + E), 256>>>(N, E, dp.conn, S.label.p, dU0, dU1, dp.Et[0], dp.Et[1], dp.Et[2], dp.Et[3], terms.p);
        count_launch();
        *energy = device_sum(terms.p, N + E);
        if (dlabels) {
            labels_to_double_kernel<<<nblk(N), 256>>>(S.label.p, N, dlabels);
            count_launch();
            SB_CUDA(cudaDeviceSynchronize());
        }
    }
}

} // namespace qpbo

bool grid_from_connectivity(int64_t N, int64_t E, const uint32_t *conn, int &H, int &W);
void launch_pairwise_tables(int H, int W, int kernel, const double *cur, const double *prop, const double *weights, double tol,
                            double d_min, double d_step, long long E, double *E00, double *E01, double *E10, double *E11);

} // namespace sb

using namespace sb;
using namespace sb::qpbo;

extern "C" {

int sb_rd_solve(int64_t N, int64_t E, const double *U0, const double *U1, const double *E00, const double *E01,
                const double *E10, const double *E11, const uint32_t *conn, int improve, double *labels,
                double *energy, double *lower_bound, double *num_unlabelled)
{
    return guarded([&] {
        SB_REQUIRE(N >= 1 && E >= 0, SB_EINVAL, "sb_rd_solve: bad sizes N=%lld E=%lld", (long long)N, (long long)E);
        SB_REQUIRE(U0 && U1 && labels && energy && lower_bound && num_unlabelled, SB_EINVAL, "sb_rd_solve: null pointer");
        SB_REQUIRE(E == 0 || (E00 && E01 && E10 && E11 && conn), SB_EINVAL, "sb_rd_solve: null pointer");
        SB_REQUIRE(N < (1LL << 30), SB_EUNSUP, "sb_rd_solve: too many nodes");
        int H = 0, W = 0;
        SB_REQUIRE(grid_from_connectivity(N, E, conn, H, W), SB_ENOTGRID,
                   "sb_rd_solve: connectivity (N=%lld, E=%lld) is not the 4-connected dispmap_super grid; "
                   "general graphs are not supported on the GPU path", (long long)N, (long long)E);
        require_device();
        DevBuf<double> dU0, dU1, dE[4];
        DevBuf<unsigned> dconn;
        auto up = [&](DevBuf<double> &b, const double *hsrc, size_t n) {
            b.alloc(std::max<size_t>(n, 1));
            if (n) SB_CUDA(cudaMemcpy(b.p, hsrc, n * 8, cudaMemcpyHostToDevice));
        };
        up(dU0, U0, (size_t)N); up(dU1, U1, (size_t)N);
        up(dE[0], E00, (size_t)E); up(dE[1], E01, (size_t)E); up(dE[2], E10, (size_t)E); up(dE[3], E11, (size_t)E);
        dconn.alloc((size_t)std::max<int64_t>(2 * E, 1));
        if (E) SB_CUDA(cudaMemcpy(dconn.p, conn, (size_t)E * 8, cudaMemcpyHostToDevice));
        DeviceProblem dp{H, W, N, E, dU0.p, dU1.p, {dE[0].p, dE[1].p, dE[2].p, dE[3].p}, dconn.p};
        std::vector<int> lab;
        solve_device(dp, improve, lab, nullptr, energy, lower_bound, num_unlabelled, nullptr);
        for (int64_t u = 0; u < N; u++) labels[u] = (double)lab[u];
    });
}

// dispmap_super.binary_fusion (dispmap_super.m:61-84) as ONE call on the grid: the four pairwise tables of
// all_pairwise_costs (dispmap_super.m:236-262) are built on the device from the two plane fields and go straight
// into the QPBO build -- the 4 x E doubles of tables (and the 2 x E connectivity) never exist on the host.
int sb_binary_fusion_grid(int H, int W, int kernel, const double *assignment, const double *proposal, const double *U0,
                          const double *U1, const double *weights, double tol, double d_min, double d_step, int improve,
                          int on_device, double *labels, double *energy, double *lower_bound, double *num_unlabelled,
                          double *stats)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1, SB_EINVAL, "sb_binary_fusion_grid: bad sizes H=%d W=%d", H, W);
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unkown kernel type");   // dispmap_super.m:232-233
        SB_REQUIRE(assignment && proposal && U0 && U1 && weights && labels && energy && lower_bound && num_unlabelled, SB_EINVAL,
                   "sb_binary_fusion_grid: null pointer");
        SB_REQUIRE(d_step != 0.0, SB_EINVAL, "sb_binary_fusion_grid: d_step == 0");
        const int64_t N = (int64_t)H * W, E = 2 * ((int64_t)(H - 1) * W + (int64_t)H * (W - 1));
        SB_REQUIRE(N < (1LL << 30), SB_EUNSUP, "sb_binary_fusion_grid: too many nodes");
        require_device();
        DevBuf<double> bcur, bprop, bU0, bU1, bw, tables((size_t)std::max<int64_t>(4 * E, 1)), dlab;
        DevBuf<unsigned> dconn((size_t)std::max<int64_t>(2 * E, 1));
        const double *cur = assignment, *prop = proposal, *u0 = U0, *u1 = U1, *wt = weights;
        if (!on_device) {
            auto up = [&](DevBuf<double> &b, const double *hsrc, size_t n) -> const double * {
                b.alloc(std::max<size_t>(n, 1));
                if (n) SB_CUDA(cudaMemcpyAsync(b.p, hsrc, n * 8, cudaMemcpyHostToDevice, 0));
                return b.p;
            };
            cur = up(bcur, assignment, (size_t)N * 4); prop = up(bprop, proposal, (size_t)N * 4);
            u0 = up(bU0, U0, (size_t)N); u1 = up(bU1, U1, (size_t)N); wt = up(bw, weights, (size_t)E);
        }
        if (E) {
            launch_pairwise_tables(H, W, kernel, cur, prop, wt, tol, d_min, d_step, E, tables.p, tables.p + E, tables.p + 2 * E,
                                   tables.p + 3 * E);
            grid_conn_kernel<<<nblk(E), 256>>>(H, W, E, dconn.p);
            SB_CUDA(cudaGetLastError());
            count_launch();
        }
        DeviceProblem dp{H, W, N, E, u0, u1, {tables.p, tables.p + E, tables.p + 2 * E, tables.p + 3 * E}, dconn.p};
        std::vector<int> lab;
        double *dlabels = nullptr;
        if (on_device) dlabels = labels;
        solve_device(dp, improve, lab, dlabels, energy, lower_bound, num_unlabelled, stats);
        if (!on_device)
            for (int64_t u = 0; u < N; u++) labels[u] = (double)lab[u];
    });
}

} // extern "C"
