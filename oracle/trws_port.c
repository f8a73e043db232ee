/* trws_port.c -- plain-C restatement of the reference's TRW-S path.
 *
 * TEST INFRASTRUCTURE ONLY: this file is the CPU oracle ("port") of the parity
 * tests and of bench.py's cpu_baseline leg.  Nothing under stereo_b200/ links,
 * imports or executes it.  Parity pinning: tests/test_oracle_port.py checks every
 * function below against the UNMODIFIED reference compiled into
 * oracle/_ref/libref_trws.so (and against fixtures generated from it under
 * tests/golden/ where the reference is absent).
 *
 * Restated here (reference file:line):
 *   solve_mrf                      cpp/trws_mex.cpp:27-147
 *   MRFEnergy::AddNode / AddEdge   cpp/trw-s/MRFEnergy.cpp:37-111
 *   SetAutomaticOrdering           cpp/trw-s/ordering.cpp:7-157
 *   CompleteGraphConstruction      cpp/trw-s/MRFEnergy.cpp:137-229
 *   SetMonotonicTrees              cpp/trw-s/treeProbabilities.cpp:12-47
 *   Minimize_TRW_S                 cpp/trw-s/minimize.cpp:7-116
 *   ComputeSolutionAndEnergy       cpp/trw-s/minimize.cpp:223-264
 *   TypeStereoLinear::Edge::UpdateMessage / AddColumn / Smooth
 *                                  cpp/trw-s/typeStereoLinear.h:324-518
 *   TypeStereoQuadratic::Edge::UpdateMessage / Smooth
 *                                  cpp/trw-s/typeStereoQuadratic.h:324-501
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int tail, head;      /* current orientation (after CompleteGraphConstruction) */
    int dir;             /* Edge::m_dir: 1 if Swap() was called               */
    int next_fwd, next_bwd;
    double alpha;
    const double *st0;   /* first stored vector  = q(:,p)     (trws_mex.cpp:101-105) */
    const double *st1;   /* second stored vector = qprim(:,p) (trws_mex.cpp:107-111) */
    int *ord0, *ord1;    /* argsorts of st0 / st1 */
    double *msg;
} edge_t;

typedef struct {
    int ordering;
    int first_fwd, first_bwd;
    int prev, next;
    int solution;
    double gamma;
} node_t;

typedef struct {
    int L, N, E, kernel;
    double lambda;
    node_t *nodes;
    edge_t *edges;
    const double *unary; /* L x N */
    int node_first, node_last;
    double *bufH, *bufZ, *bufDi, *bufDb;
    int *bufV;
} mrf_t;

/* ---- argsort (trws_mex.cpp:84-99; ties: lower label first) */
typedef struct { double v; int i; } pair_t;
static int cmp_pair(const void *a, const void *b)
{
    const pair_t *x = (const pair_t *)a, *y = (const pair_t *)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return (x->i > y->i) - (x->i < y->i);
}
static void argsort(const double *v, int L, int *out, pair_t *tmp)
{
    for (int i = 0; i < L; i++) { tmp[i].v = v[i]; tmp[i].i = i; }
    qsort(tmp, (size_t)L, sizeof(pair_t), cmp_pair);
    for (int i = 0; i < L; i++) out[i] = tmp[i].i;
}

/* ---- ordering.cpp:24-152 (list / listBoundary / ordered, first-minimum tie rule) */
static int set_automatic_ordering(mrf_t *m)
{
    const int N = m->N;
    node_t *nd = m->nodes;
    edge_t *ed = m->edges;
    int *deg = (int *)calloc((size_t)N, sizeof(int));
    char *state = (char *)calloc((size_t)N, 1); /* 0 list, 1 boundary, 2 ordered */
    for (int i = 0; i < N; i++) {
        for (int e = nd[i].first_fwd; e >= 0; e = ed[e].next_fwd) deg[i]++;
        for (int e = nd[i].first_bwd; e >= 0; e = ed[e].next_bwd) deg[i]++;
    }
    int list = m->node_first, boundary = -1, last = -1, count = 0, ok = 1;
    m->node_first = m->node_last = -1;
    while (list >= 0 && ok) {
        int dMin = N, iMin = -1;
        for (int i = list; i >= 0; i = nd[i].next)
            if (dMin > deg[i]) { dMin = deg[i]; iMin = i; }
        if (iMin < 0) { ok = 0; break; } /* reference dereferences garbage here */
        int i = iMin;
        if (nd[i].prev >= 0) nd[nd[i].prev].next = nd[i].next; else list = nd[i].next;
        if (nd[i].next >= 0) nd[nd[i].next].prev = nd[i].prev;
        boundary = i; nd[i].prev = nd[i].next = -1; state[i] = 1;
        while (boundary >= 0) {
            dMin = N; iMin = -1;
            for (i = boundary; i >= 0; i = nd[i].next)
                if (dMin > deg[i]) { dMin = deg[i]; iMin = i; }
            if (iMin < 0) { ok = 0; break; }
            i = iMin;
            if (nd[i].prev >= 0) nd[nd[i].prev].next = nd[i].next; else boundary = nd[i].next;
            if (nd[i].next >= 0) nd[nd[i].next].prev = nd[i].prev;
            if (last >= 0) nd[last].next = i; else m->node_first = i;
            nd[i].ordering = count++;
            nd[i].prev = last; nd[i].next = -1; last = i; state[i] = 2;
            for (int pass = 0; pass < 2; pass++) {
                for (int e = pass ? nd[last].first_bwd : nd[last].first_fwd; e >= 0;
                     e = pass ? ed[e].next_bwd : ed[e].next_fwd) {
                    int j = pass ? ed[e].tail : ed[e].head;
                    if (state[j] == 2) continue;
                    deg[j]--;
                    if (state[j] == 0) {
                        if (nd[j].prev >= 0) nd[nd[j].prev].next = nd[j].next; else list = nd[j].next;
                        if (nd[j].next >= 0) nd[nd[j].next].prev = nd[j].prev;
                        if (boundary >= 0) nd[boundary].prev = j;
                        nd[j].prev = -1; nd[j].next = boundary; boundary = j; state[j] = 1;
                    }
                }
            }
        }
    }
    m->node_last = last;
    free(deg); free(state);
    return ok && count == N;
}

/* ---- MRFEnergy.cpp:178-224 */
static void complete_graph_construction(mrf_t *m)
{
    node_t *nd = m->nodes;
    edge_t *ed = m->edges;
    for (int i = m->node_first; i >= 0; i = nd[i].next) nd[i].first_bwd = -1;
    for (int i = m->node_first; i >= 0; i = nd[i].next) {
        int ePrev = -1;
        for (int e = nd[i].first_fwd; e >= 0;) {
            int j = ed[e].head;
            if (nd[i].ordering < nd[j].ordering) {
                ed[e].next_bwd = nd[j].first_bwd; nd[j].first_bwd = e;
                ePrev = e; e = ed[e].next_fwd;
            } else {
                ed[e].dir = 1 - ed[e].dir; /* Swap() */
                ed[e].tail = j; ed[e].head = i;
                int eNext = ed[e].next_fwd;
                if (ePrev >= 0) ed[ePrev].next_fwd = ed[e].next_fwd; else nd[i].first_fwd = ed[e].next_fwd;
                ed[e].next_fwd = nd[j].first_fwd; nd[j].first_fwd = e;
                ed[e].next_bwd = nd[i].first_bwd; nd[i].first_bwd = e;
                e = eNext;
            }
        }
    }
}

/* ---- treeProbabilities.cpp:21-46 */
static void set_monotonic_trees(mrf_t *m)
{
    node_t *nd = m->nodes;
    edge_t *ed = m->edges;
    for (int i = m->node_first; i >= 0; i = nd[i].next) {
        int nF = 0, nB = 0;
        for (int e = nd[i].first_fwd; e >= 0; e = ed[e].next_fwd) nF++;
        for (int e = nd[i].first_bwd; e >= 0; e = ed[e].next_bwd) nB++;
        int ni = nF > nB ? nF : nB;
        nd[i].gamma = ni ? 1.0 / ni : 1.0;
    }
}

static double smooth(int kernel, double alpha, double lambda, double val)
{
    double c = kernel == 1 ? fabs(val) : val * val;
    return alpha * (c < lambda ? c : lambda);
}

/* ---- UpdateMessage, both kernels.  source = Di of the sending node. */
static double update_message(int kernel, int L, double alpha, double lambda, double *msg,
                             const double *st0, const double *st1, const int *ord0, const int *ord1,
                             int edge_dir, const double *source, double gamma, int dir,
                             double *H, double *z, int *v)
{
    const double *q, *qprim;
    const int *q_order, *qprim_order;
    if (dir == edge_dir) { q = st1; qprim = st0; q_order = ord1; qprim_order = ord0; }
    else                 { q = st0; qprim = st1; q_order = ord0; qprim_order = ord1; }
    double vTrunc = INFINITY, vMin = INFINITY;
    int j = 0, k, l;
    for (k = 0; k < L; k++) {
        H[k] = gamma * source[k] - msg[k];
        if (H[k] < vTrunc) vTrunc = H[k];
    }
    if (alpha == 0) {
        for (k = 0; k < L; k++) msg[k] = vTrunc;
        vMin = vTrunc;
    } else if (kernel == 1) {
        /* typeStereoLinear.h:375-480: lower envelope of cones */
        v[0] = q_order[0]; z[0] = -INFINITY; z[1] = INFINITY;
        vTrunc += alpha * lambda;
        for (k = 1; k < L; k++) {
            double hk = H[q_order[k]], qk = q[q_order[k]];
            for (l = k; l >= 0; l--) {
                double hj = H[v[j]], qj = q[v[j]];
                double dist = alpha * fabs(qk - qj);
                if ((dist + hk) < hj) {
                    if (j == 0) { v[0] = q_order[k]; z[0] = -INFINITY; z[1] = INFINITY; }
                    else j--;
                } else if ((dist + hj) <= hk) {
                    break;
                } else {
                    double s = ((hk - hj) + alpha * (qk + qj)) / (2 * alpha);
                    if (s >= qk) break;
                    if (s <= qj) break;
                    j++; v[j] = q_order[k]; z[j] = s; z[j + 1] = INFINITY;
                    break;
                }
            }
        }
        j = 0;
        for (k = 0; k < L; k++) {
            double qprimk = qprim[qprim_order[k]];
            while (z[j + 1] < qprimk) j++;
            double val = alpha * fabs(qprimk - q[v[j]]) + H[v[j]];
            if (vTrunc < val) val = vTrunc;
            msg[qprim_order[k]] = val;
            if (val < vMin) vMin = val;
        }
    } else {
        /* typeStereoQuadratic.h:405-496: lower envelope of parabolas */
        vTrunc += alpha * lambda;
        v[0] = q_order[0]; z[0] = -INFINITY; z[1] = INFINITY;
        for (k = 1; k < L; k++) {
            double hk = H[q_order[k]], qk = q[q_order[k]];
            for (l = k; l >= 0; l--) {
                double hj = H[v[j]], qj = q[v[j]];
                if ((qk - qj) < 1e-8) {
                    if (hj > hk) {
                        if (j == 0) { v[0] = q_order[k]; z[0] = -INFINITY; z[1] = INFINITY; break; }
                        else j--;
                    } else break;
                } else {
                    double s = ((hk + alpha * qk * qk) - (hj + alpha * qj * qj)) / (2 * alpha * (qk - qj));
                    if (s <= z[j]) j--;
                    else { j++; v[j] = q_order[k]; z[j] = s; z[j + 1] = INFINITY; break; }
                }
            }
        }
        j = 0;
        for (k = 0; k < L; k++) {
            double qprimk = qprim[qprim_order[k]];
            while (z[j + 1] < qprimk) j++;
            double val = qprimk - q[v[j]];
            val = alpha * val * val + H[v[j]];
            if (vTrunc < val) val = vTrunc;
            msg[qprim_order[k]] = val;
        }
        for (k = 0; k < L; k++) if (msg[k] < vMin) vMin = msg[k];
    }
    for (k = 0; k < L; k++) msg[k] -= vMin;
    return vMin;
}

/* ---- AddColumn, typeStereoLinear.h:491-518 */
static void add_column(const mrf_t *m, const edge_t *e, int ksource, double *dest, int dir)
{
    const double *q = e->st0, *qprim = e->st1;
    if (dir == e->dir) for (int k = 0; k < m->L; k++) dest[k] += smooth(m->kernel, e->alpha, m->lambda, qprim[ksource] - q[k]);
    else               for (int k = 0; k < m->L; k++) dest[k] += smooth(m->kernel, e->alpha, m->lambda, qprim[k] - q[ksource]);
}

/* ---- minimize.cpp:223-264 */
static double compute_solution_and_energy(mrf_t *m)
{
    node_t *nd = m->nodes;
    edge_t *ed = m->edges;
    const int L = m->L;
    double E = 0, *DiB = m->bufDb, *Di = m->bufDi;
    for (int i = m->node_first; i >= 0; i = nd[i].next) {
        memcpy(DiB, m->unary + (size_t)i * L, sizeof(double) * L);
        for (int e = nd[i].first_bwd; e >= 0; e = ed[e].next_bwd)
            add_column(m, &ed[e], nd[ed[e].tail].solution, DiB, 0);
        memcpy(Di, DiB, sizeof(double) * L);
        for (int e = nd[i].first_fwd; e >= 0; e = ed[e].next_fwd)
            for (int k = 0; k < L; k++) Di[k] += ed[e].msg[k];
        double vMin = Di[0]; int kMin = 0;
        for (int k = 1; k < L; k++) if (vMin > Di[k]) { vMin = Di[k]; kMin = k; }
        nd[i].solution = kMin;
        E += DiB[kMin];
    }
    return E;
}

/* ---- minimize.cpp:7-116 */
static int minimize_trws(mrf_t *m, int iterMax, double relgapMax, double *lowerBound, double *energy)
{
    node_t *nd = m->nodes;
    edge_t *ed = m->edges;
    const int L = m->L;
    double *Di = m->bufDi;
    int iter;
    set_monotonic_trees(m);
    for (iter = 1;; iter++) {
        for (int i = m->node_first; i >= 0; i = nd[i].next) {
            memcpy(Di, m->unary + (size_t)i * L, sizeof(double) * L);
            for (int e = nd[i].first_fwd; e >= 0; e = ed[e].next_fwd) for (int k = 0; k < L; k++) Di[k] += ed[e].msg[k];
            for (int e = nd[i].first_bwd; e >= 0; e = ed[e].next_bwd) for (int k = 0; k < L; k++) Di[k] += ed[e].msg[k];
            for (int e = nd[i].first_fwd; e >= 0; e = ed[e].next_fwd)
                update_message(m->kernel, L, ed[e].alpha, m->lambda, ed[e].msg, ed[e].st0, ed[e].st1, ed[e].ord0,
                               ed[e].ord1, ed[e].dir, Di, nd[i].gamma, 0, m->bufH, m->bufZ, m->bufV);
        }
        *lowerBound = 0;
        for (int i = m->node_last; i >= 0; i = nd[i].prev) {
            memcpy(Di, m->unary + (size_t)i * L, sizeof(double) * L);
            for (int e = nd[i].first_bwd; e >= 0; e = ed[e].next_bwd) for (int k = 0; k < L; k++) Di[k] += ed[e].msg[k];
            for (int e = nd[i].first_fwd; e >= 0; e = ed[e].next_fwd) for (int k = 0; k < L; k++) Di[k] += ed[e].msg[k];
            double vMin = Di[0];
            for (int k = 1; k < L; k++) if (vMin > Di[k]) vMin = Di[k];
            for (int k = 0; k < L; k++) Di[k] -= vMin;
            *lowerBound += vMin;
            for (int e = nd[i].first_bwd; e >= 0; e = ed[e].next_bwd)
                *lowerBound += update_message(m->kernel, L, ed[e].alpha, m->lambda, ed[e].msg, ed[e].st0, ed[e].st1,
                                              ed[e].ord0, ed[e].ord1, ed[e].dir, Di, nd[i].gamma, 1, m->bufH,
                                              m->bufZ, m->bufV);
        }
        int finish = iter >= iterMax;
        *energy = compute_solution_and_energy(m);
        if ((*energy - *lowerBound) / *energy < relgapMax) finish = 1;
        if (finish) break;
    }
    return iter;
}

static void mrf_free(mrf_t *m)
{
    if (m->edges) {
        for (int p = 0; p < m->E; p++) { free(m->edges[p].ord0); free(m->edges[p].msg); }
    }
    free(m->edges); free(m->nodes); free(m->bufH); free(m->bufZ); free(m->bufV); free(m->bufDi); free(m->bufDb);
}

/* build nodes + edges exactly as trws_mex.cpp:60-119 / MRFEnergy.cpp:37-111 do */
static int mrf_build(mrf_t *m, int kernel, int L, int N, int E, const double *unary, const uint32_t *conn,
                     const double *q, const double *qprim, const double *alphas, double tol, int with_data)
{
    memset(m, 0, sizeof(*m));
    m->L = L; m->N = N; m->E = E; m->kernel = kernel; m->lambda = tol; m->unary = unary;
    m->nodes = (node_t *)calloc((size_t)N, sizeof(node_t));
    m->edges = (edge_t *)calloc((size_t)(E ? E : 1), sizeof(edge_t));
    m->bufH = (double *)malloc(sizeof(double) * (L + 2));
    m->bufZ = (double *)malloc(sizeof(double) * (L + 2));
    m->bufV = (int *)malloc(sizeof(int) * (L + 2));
    m->bufDi = (double *)malloc(sizeof(double) * L);
    m->bufDb = (double *)malloc(sizeof(double) * L);
    for (int i = 0; i < N; i++) {
        m->nodes[i].ordering = i; m->nodes[i].first_fwd = m->nodes[i].first_bwd = -1;
        m->nodes[i].prev = i - 1; m->nodes[i].next = i + 1 < N ? i + 1 : -1;
    }
    m->node_first = N ? 0 : -1; m->node_last = N - 1;
    pair_t *tmp = (pair_t *)malloc(sizeof(pair_t) * L);
    for (int p = 0; p < E; p++) {
        edge_t *e = &m->edges[p];
        int i = (int)conn[2 * p], j = (int)conn[2 * p + 1];
        if (i < 0 || i >= N || j < 0 || j >= N) { free(tmp); return 0; }
        e->tail = i; e->head = j; e->dir = 0;
        e->next_fwd = m->nodes[i].first_fwd; m->nodes[i].first_fwd = p;
        e->next_bwd = m->nodes[j].first_bwd; m->nodes[j].first_bwd = p;
        if (with_data) {
            e->alpha = alphas[p];
            e->st0 = q + (size_t)p * L; e->st1 = qprim + (size_t)p * L;
            e->ord0 = (int *)malloc(sizeof(int) * 2 * L); e->ord1 = e->ord0 + L;
            argsort(e->st0, L, e->ord0, tmp);
            argsort(e->st1, L, e->ord1, tmp);
            e->msg = (double *)calloc((size_t)L, sizeof(double));
        }
    }
    free(tmp);
    return 1;
}

int port_trws_solve(int kernel, int L, int64_t N, int64_t E, const double *unary, const uint32_t *conn,
                    const double *q, const double *qprim, const double *alphas, double tol, double maxiter,
                    double max_relgap, double *labels, double *energy, double *lower_bound, double *iterations)
{
    if (kernel != 1 && kernel != 2) return -1;
    mrf_t m;
    if (!mrf_build(&m, kernel, L, (int)N, (int)E, unary, conn, q, qprim, alphas, tol, 1)) { mrf_free(&m); return -1; }
    if (!set_automatic_ordering(&m)) { mrf_free(&m); return -2; }
    complete_graph_construction(&m);
    double lb = 0, en = 0;
    int it = minimize_trws(&m, (int)maxiter, max_relgap, &lb, &en);
    for (int u = 0; u < (int)N; u++) labels[u] = m.nodes[u].solution + 1;
    *energy = en; *lower_bound = lb; *iterations = it;
    mrf_free(&m);
    return 0;
}

/* grid terms in dispmap_super.construct_neighborhood order (dispmap_super.m:279-302), 0-based */
static uint32_t *grid_conn(int H, int W, int *E_out)
{
    int nV = (H - 1) * W, nH = H * (W - 1), E = 2 * (nV + nH), p = 0;
    uint32_t *c = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (size_t)(E ? E : 1));
    for (int x = 0; x < W; x++) for (int r = 0; r < H - 1; r++, p++) { c[2 * p] = r + H * x; c[2 * p + 1] = r + 1 + H * x; }
    for (int x = 0; x < W; x++) for (int r = 0; r < H - 1; r++, p++) { c[2 * p] = r + 1 + H * x; c[2 * p + 1] = r + H * x; }
    for (int x = 0; x < W - 1; x++) for (int r = 0; r < H; r++, p++) { c[2 * p] = r + H * x; c[2 * p + 1] = r + H * (x + 1); }
    for (int x = 0; x < W - 1; x++) for (int r = 0; r < H; r++, p++) { c[2 * p] = r + H * (x + 1); c[2 * p + 1] = r + H * x; }
    *E_out = E;
    return c;
}

int port_trws_ordering(int H, int W, int32_t *ordering_out)
{
    int E = 0;
    uint32_t *conn = grid_conn(H, W, &E);
    mrf_t m;
    int ok = mrf_build(&m, 1, 1, H * W, E, NULL, conn, NULL, NULL, NULL, 0, 0);
    if (ok) ok = set_automatic_ordering(&m);
    if (ok) for (int u = 0; u < H * W; u++) ordering_out[u] = m.nodes[u].ordering;
    mrf_free(&m);
    free(conn);
    return ok ? 0 : -1;
}

int port_trws_update_message(int kernel, int L, const double *Di, double *msg, const double *stored0,
                             const double *stored1, const int *order0, const int *order1, double alpha,
                             double lambda, double gamma, int dir, int swapped, double *vmin_out)
{
    double *H = (double *)malloc(sizeof(double) * (L + 2)), *z = (double *)malloc(sizeof(double) * (L + 2));
    int *v = (int *)malloc(sizeof(int) * (L + 2));
    *vmin_out = update_message(kernel, L, alpha, lambda, msg, stored0, stored1, order0, order1, swapped, Di, gamma,
                               dir, H, z, v);
    free(H); free(z); free(v);
    return 0;
}
