// trws_order.cpp -- host-side graph logic of the TRW-S path (no CUDA):
//   * recognise the dispmap_super grid in a connectivity list
//     (dispmap_super.m:279-302, as shifted to 0-based by trws.m:33 / rd.m:21);
//   * reproduce the node ordering the reference solver runs under
//     (MRFEnergy::SetAutomaticOrdering, cpp/trw-s/ordering.cpp:7-157): closed
//     form for H,W >= 4 (verified against the compiled
//     reference in tests/test_ordering.py), literal greedy scan otherwise;
//   * derive the dispatch schedule (longest-path levels of the orientation DAG,
//     MRFEnergy.cpp:188-219) the sweep kernels consume.
#include "trws_order.h"
#include <algorithm>
#include <numeric>
#include <cstring>
#include <cstdlib>

namespace sb {

void grid_terms(int H, int W, std::vector<int32_t> &tail, std::vector<int32_t> &head)
{
    const int64_t nV = (int64_t)(H - 1) * W, nH = (int64_t)H * (W - 1);
    tail.resize(2 * (nV + nH));
    head.resize(2 * (nV + nH));
    int64_t p = 0;
    for (int c = 0; c < W; c++) for (int r = 0; r < H - 1; r++, p++) { tail[p] = r + H * c; head[p] = r + 1 + H * c; }
    for (int c = 0; c < W; c++) for (int r = 0; r < H - 1; r++, p++) { tail[p] = r + 1 + H * c; head[p] = r + H * c; }
    for (int c = 0; c < W - 1; c++) for (int r = 0; r < H; r++, p++) { tail[p] = r + H * c; head[p] = r + H * (c + 1); }
    for (int c = 0; c < W - 1; c++) for (int r = 0; r < H; r++, p++) { tail[p] = r + H * (c + 1); head[p] = r + H * c; }
}

bool grid_from_connectivity(int64_t N, int64_t E, const uint32_t *conn, int &H, int &W)
{
    if (N <= 0 || E < 0) return false;
    if (E == 0) {
        if (N == 1) { H = W = 1; return true; }
        return false;
    }
    // The first term is (0 -> 1) when H > 1 (vertical down) and (0 -> H) = (0 -> 1)
    // when H == 1 too, so H has to come from the edge count: E = 2[(H-1)W + H(W-1)]
    // with HW = N  =>  H + W = 2N - E/2.
    if (E % 2) return false;
    const int64_t sum = 2 * N - E / 2; // H + W
    // H, W are the roots of x^2 - sum x + N = 0
    const double disc = (double)sum * (double)sum - 4.0 * (double)N;
    if (disc < -0.5) return false;
    const int64_t root = (int64_t)llround(std::sqrt(std::max(0.0, disc)));
    int64_t a = (sum - root) / 2, b = (sum + root) / 2;
    if (a <= 0 || a * b != N || a + b != sum) return false;
    // two candidates (a x b and b x a); the vertical block decides.
    for (int pass = 0; pass < 2; pass++) {
        const int64_t h = pass ? b : a, w = pass ? a : b;
        if (pass && a == b) break;
        if (h > INT32_MAX || w > INT32_MAX) continue;
        bool ok = true;
        const int64_t nV = (h - 1) * w, nH = h * (w - 1);
        if (2 * (nV + nH) != E) continue;
        int64_t p = 0;
        for (int64_t c = 0; c < w && ok; c++)
            for (int64_t r = 0; r < h - 1; r++, p++) {
                const uint32_t u = (uint32_t)(r + h * c);
                if (conn[2 * p] != u || conn[2 * p + 1] != u + 1 ||
                    conn[2 * (p + nV)] != u + 1 || conn[2 * (p + nV) + 1] != u) { ok = false; break; }
            }
        p = 2 * nV;
        for (int64_t c = 0; c < w - 1 && ok; c++)
            for (int64_t r = 0; r < h; r++, p++) {
                const uint32_t u = (uint32_t)(r + h * c), v = (uint32_t)(r + h * (c + 1));
                if (conn[2 * p] != u || conn[2 * p + 1] != v ||
                    conn[2 * (p + nH)] != v || conn[2 * (p + nH) + 1] != u) { ok = false; break; }
            }
        if (ok) { H = (int)h; W = (int)w; return true; }
    }
    return false;
}

// Literal restatement of the greedy minimum-remaining-degree scan of
// ordering.cpp:24-152 for an arbitrary term list.  Lists are intrusive doubly
// linked lists with insertion at the front, exactly as in the reference, because
// the tie-breaking ("first minimum in list order", strict >, ordering.cpp:49,75)
// depends on list order.  Adjacency is visited in the reference's order: terms
// with the node as tail, newest first (MRFEnergy.cpp:98-100), then terms with
// the node as head, newest first.
bool greedy_ordering(int64_t N, const std::vector<int32_t> &tail, const std::vector<int32_t> &head,
                     std::vector<int32_t> &ordering)
{
    const int64_t E = (int64_t)tail.size();
    std::vector<int64_t> fwd_first(N, -1), bwd_first(N, -1), fwd_next(E, -1), bwd_next(E, -1);
    for (int64_t p = 0; p < E; p++) {
        fwd_next[p] = fwd_first[tail[p]]; fwd_first[tail[p]] = p;
        bwd_next[p] = bwd_first[head[p]]; bwd_first[head[p]] = p;
    }
    // state: 0 = in `list`, 1 = in `listBoundary`, 2 = ordered
    std::vector<int64_t> deg(N, 0), prev(N), next(N);
    std::vector<char> state(N, 0);
    for (int64_t p = 0; p < E; p++) { deg[tail[p]]++; deg[head[p]]++; }
    for (int64_t u = 0; u < N; u++) { prev[u] = u - 1; next[u] = (u + 1 < N) ? u + 1 : -1; }
    int64_t list = N ? 0 : -1, boundary = -1, count = 0;
    ordering.assign(N, -1);
    auto unlink = [&](int64_t i, int64_t &first) {
        if (prev[i] >= 0) next[prev[i]] = next[i]; else first = next[i];
        if (next[i] >= 0) prev[next[i]] = prev[i];
    };
    auto touch = [&](int64_t i) {
        if (state[i] == 2) return;
        deg[i]--;
        if (state[i] == 0) {
            unlink(i, list);
            if (boundary >= 0) prev[boundary] = i;
            prev[i] = -1; next[i] = boundary; boundary = i;
            state[i] = 1;
        }
    };
    while (list >= 0) {
        int64_t dMin = N, iMin = -1;
        for (int64_t i = list; i >= 0; i = next[i])
            if (dMin > deg[i]) { dMin = deg[i]; iMin = i; }
        if (iMin < 0) return false; // the reference reads an uninitialised pointer here (2x2 grid)
        unlink(iMin, list);
        boundary = iMin; prev[iMin] = next[iMin] = -1; state[iMin] = 1;
        while (boundary >= 0) {
            dMin = N; iMin = -1;
            for (int64_t i = boundary; i >= 0; i = next[i])
                if (dMin > deg[i]) { dMin = deg[i]; iMin = i; }
            if (iMin < 0) return false;
            unlink(iMin, boundary);
            ordering[iMin] = (int32_t)count++;
            state[iMin] = 2;
            for (int64_t p = fwd_first[iMin]; p >= 0; p = fwd_next[p]) touch(head[p]);
            for (int64_t p = bwd_first[iMin]; p >= 0; p = bwd_next[p]) touch(tail[p]);
        }
    }
    return count == N;
}

bool grid_ordering(int H, int W, std::vector<int32_t> &ordering)
{
    const int64_t N = (int64_t)H * W;
    if (H >= 4 && W >= 4) {
        // SURVEY.md Appendix A.1 (first matching rule wins)
        ordering.resize(N);
        const int64_t ring = 2LL * H + 2LL * W - 4;
        const int64_t B = ring + (int64_t)(H - 4) * (W - 2);
        for (int c = 0; c < W; c++)
            for (int r = 0; r < H; r++) {
                int64_t v;
                if (c == 0) v = r;
                else if (r == H - 1) v = H - 1 + c;
                else if (c == W - 1) v = (int64_t)H + W - 2 + (H - 1 - r);
                else if (r == 0) v = 2LL * H + W - 3 + (W - 1 - c);
                else if (r <= H - 4) v = ring + (int64_t)(r - 1) * (W - 2) + (W - 2 - c);
                else if (c >= 2) v = B + 2LL * (W - 2 - c) + (r == H - 2 ? 1 : 0);
                else v = B + 2LL * (W - 3) + (r == H - 3 ? 1 : 0);
                ordering[r + (int64_t)H * c] = (int32_t)v;
            }
        return true;
    }
    if (N == 1) { ordering.assign(1, 0); return true; }
    std::vector<int32_t> tail, head;
    grid_terms(H, W, tail, head);
    return greedy_ordering(N, tail, head, ordering);
}

// Per-node incidence byte: bit d (0 up, 1 down, 2 left, 3 right) = neighbour exists,
// bit 4+d = that neighbour has a LOWER ordering (its terms are backward edges of this
// node, MRFEnergy.cpp:188-219).  gamma and the send / wait sets follow from it.
void build_node_info(int H, int W, const std::vector<int32_t> &ordering, std::vector<uint8_t> &info)
{
    const int64_t N = (int64_t)H * W;
    info.resize(N);
    for (int c = 0; c < W; c++)
        for (int r = 0; r < H; r++) {
            const int64_t u = r + (int64_t)H * c;
            const int32_t o = ordering[u];
            unsigned b = 0;
            if (r > 0) { b |= 1u; if (ordering[u - 1] < o) b |= 16u; }
            if (r < H - 1) { b |= 2u; if (ordering[u + 1] < o) b |= 32u; }
            if (c > 0) { b |= 4u; if (ordering[u - H] < o) b |= 64u; }
            if (c < W - 1) { b |= 8u; if (ordering[u + H] < o) b |= 128u; }
            info[u] = (uint8_t)b;
        }
}

// Dispatch schedule of one sweep: a list of STRIPS, each a run of nodes one warp
// processes back to back.  A node's forward sweep may run once all lower-ordered
// neighbours are done (minimize.cpp:36-62 visits nodes by m_ordering; only messages on
// incident edges are read or written), so any dispatch order in which a strip only
// waits for strips dispatched before it (plus the one documented exception below)
// reproduces the sequential sweep.
//   regular grids (H,W >= 4): strip 0 is the boundary ring in ordering sequence (a serial
//     chain: every ring node's predecessor is its neighbour), strips 1..H-2 are the
//     interior rows, swept right to left; row r only waits for the ring and row r-1,
//     except that (H-3,1) -- the last node of the ordering -- waits for (H-2,1);
//   other grids: one node per strip, sorted by longest-path level of the DAG.
// The backward sweep uses the exact reverse (strips and nodes within strips).
void build_schedule(int H, int W, const std::vector<int32_t> &ordering, Schedule &s, int world)
{
    const int64_t N = (int64_t)H * W;
    s.nodes.clear();
    s.strip_ptr.clear();
    s.owner.clear();
    s.nodes.reserve(N);
    s.world = world;
    SB_REQUIRE(world >= 1, SB_EINVAL, "build_schedule: bad world size");
    SB_REQUIRE(world == 1 || (H >= 4 && W >= 4 && world <= H / 2), SB_EUNSUP,
               "row-banded sweeps need a regular grid with at least two rows per rank");
    if (H >= 4 && W >= 4) {
        const int64_t ring = 2LL * H + 2LL * W - 4;
        std::vector<int32_t> ringnodes(ring);
        auto put = [&](int r, int c) { const int64_t u = r + (int64_t)H * c; ringnodes[ordering[u]] = (int32_t)u; };
        for (int r = 0; r < H; r++) { put(r, 0); put(r, W - 1); }
        for (int c = 1; c < W - 1; c++) { put(0, c); put(H - 1, c); }
        // the ring in ordering sequence, cut where the owning band changes
        for (int64_t k = 0; k < ring; k++) {
            const int own = band_of_row(ringnodes[k] % H, H, world);
            if (k == 0 || own != s.owner.back()) {
                s.strip_ptr.push_back((int64_t)s.nodes.size());
                s.owner.push_back(own);
            }
            s.nodes.push_back(ringnodes[k]);
        }
        for (int r = 1; r <= H - 2; r++) {
            s.strip_ptr.push_back((int64_t)s.nodes.size());
            s.owner.push_back(band_of_row(r, H, world));
            for (int c = W - 2; c >= 1; c--) s.nodes.push_back((int32_t)(r + (int64_t)H * c));
        }
        s.strip_ptr.push_back((int64_t)s.nodes.size());
        s.regular = true;
    } else {
        std::vector<int32_t> inv(N), level(N, 0);
        for (int64_t u = 0; u < N; u++) inv[ordering[u]] = (int32_t)u;
        int32_t maxl = 0;
        for (int64_t o = 0; o < N; o++) {
            const int64_t u = inv[o];
            const int r = (int)(u % H), c = (int)(u / H);
            int32_t l = 0;
            auto dep = [&](int64_t v) { if (ordering[v] < o) l = std::max(l, level[v] + 1); };
            if (r > 0) dep(u - 1);
            if (r < H - 1) dep(u + 1);
            if (c > 0) dep(u - H);
            if (c < W - 1) dep(u + H);
            level[u] = l;
            maxl = std::max(maxl, l);
        }
        std::vector<int64_t> start(maxl + 2, 0);
        for (int64_t u = 0; u < N; u++) start[level[u] + 1]++;
        for (int32_t l = 0; l <= maxl; l++) start[l + 1] += start[l];
        s.nodes.resize(N);
        for (int64_t o = 0; o < N; o++) {
            const int64_t u = inv[o];
            s.nodes[start[level[u]]++] = (int32_t)u;
        }
        s.strip_ptr.resize(N + 1);
        for (int64_t i = 0; i <= N; i++) s.strip_ptr[i] = i;
        s.owner.assign((size_t)N, 0);
        s.regular = false;
    }
}

void build_schedule_cols(int H, int W, const std::vector<int32_t> &ordering, Schedule &s, int world, int blocks)
{
    if (world <= 1) { build_schedule(H, W, ordering, s, 1); return; }
    SB_REQUIRE(blocks >= 1 && H >= 4 && W >= 4 && world * blocks <= W / 4, SB_EUNSUP,
               "column-banded sweeps need a regular grid with at least four columns per block");
    const int wb = col_block_width(W, world, blocks);
    SB_REQUIRE((W + wb - 1) / wb >= world, SB_EUNSUP, "column-banded sweeps: fewer column blocks than ranks");
    s.nodes.clear();
    s.strip_ptr.clear();
    s.owner.clear();
    s.nodes.reserve((int64_t)H * W);
    s.world = world;
    const int64_t ring = 2LL * H + 2LL * W - 4;
    std::vector<int32_t> ringnodes(ring);
    auto put = [&](int r, int c) { const int64_t u = r + (int64_t)H * c; ringnodes[ordering[u]] = (int32_t)u; };
    for (int r = 0; r < H; r++) { put(r, 0); put(r, W - 1); }
    for (int c = 1; c < W - 1; c++) { put(0, c); put(H - 1, c); }
    for (int64_t k = 0; k < ring; k++) {
        const int own = band_of_col(ringnodes[k] / H, wb, world);
        if (k == 0 || own != s.owner.back()) {
            s.strip_ptr.push_back((int64_t)s.nodes.size());
            s.owner.push_back(own);
        }
        s.nodes.push_back(ringnodes[k]);
    }
    // interior rows run right to left: the piece of the last rank first
    for (int r = 1; r <= H - 2; r++) {
        int cur = -1;
        for (int c = W - 2; c >= 1; c--) {
            const int own = band_of_col(c, wb, world);
            if (own != cur) {
                s.strip_ptr.push_back((int64_t)s.nodes.size());
                s.owner.push_back(own);
                cur = own;
            }
            s.nodes.push_back((int32_t)(r + (int64_t)H * c));
        }
    }
    s.strip_ptr.push_back((int64_t)s.nodes.size());
    s.regular = true;
}


// ---------------------------------------------------------------------------
// Segment descriptors (trws_sched.h).
namespace {

struct NodeDesc {
    int u, halves, gamma_den, use_carry, nitems;
    struct Own { long long term; int flags; } own[trws::SCHED_NCW][2];
    struct Item { long long term; int kind; int strip; int need; } item[trws::SCHED_ITEMS];
};

struct Inc { int nb; long long term; bool tail; };

// Terms between node (r, c) and its neighbour in direction d (0 up, 1 down, 2 left,
// 3 right); j picks the term of the pair.  Order of dispmap_super.m:284-294.
inline Inc incidence(int d, int j, int r, int c, int u, int H, long long nV, long long nH)
{
    Inc I;
    if (d == 0) { I.nb = u - 1; I.term = (long long)c * (H - 1) + (r - 1) + (j ? nV : 0); I.tail = (j == 1); }
    else if (d == 1) { I.nb = u + 1; I.term = (long long)c * (H - 1) + r + (j ? nV : 0); I.tail = (j == 0); }
    else if (d == 2) { I.nb = u - H; I.term = 2 * nV + (long long)(c - 1) * H + r + (j ? nH : 0); I.tail = (j == 1); }
    else { I.nb = u + H; I.term = 2 * nV + (long long)c * H + r + (j ? nH : 0); I.tail = (j == 0); }
    return I;
}

inline int direction_to(int u, int v, int H)
{
    const int d = v - u;
    return d == -1 ? 0 : d == 1 ? 1 : d == -H ? 2 : d == H ? 3 : -1;
}

} // namespace

void build_pass_plan(int H, int W, const std::vector<uint8_t> &info, const Schedule &s, int pass, PassPlan &plan,
                     int rank)
{
    using namespace trws;
    const int64_t N = (int64_t)H * W;
    const long long nV = (long long)(H - 1) * W, nH = (long long)H * (W - 1);
    const int S = (int)s.strip_ptr.size() - 1;
    // where every node sits: strip and index within the strip (forward order)
    std::vector<int32_t> strip_of((size_t)N), idx_of((size_t)N);
    for (int fs = 0; fs < S; fs++)
        for (int64_t k = s.strip_ptr[fs]; k < s.strip_ptr[fs + 1]; k++) {
            strip_of[s.nodes[k]] = fs;
            idx_of[s.nodes[k]] = (int32_t)(k - s.strip_ptr[fs]);
        }
    auto need_of = [&](int v) {
        const int fs = strip_of[v];
        const int len = (int)(s.strip_ptr[fs + 1] - s.strip_ptr[fs]);
        return pass == 0 ? idx_of[v] + 1 : len - idx_of[v];
    };
    auto describe = [&](int u, int u_prev, int u_next, NodeDesc &nd) {
        std::memset(&nd, 0, sizeof(nd));
        for (int q = 0; q < SCHED_ITEMS; q++) nd.item[q].strip = -1;
        const int r = u % H, c = u / H;
        const unsigned valid = info[u] & 15u, lower = info[u] >> 4;
        const unsigned send_mask = pass == 0 ? (valid & ~lower) : lower;
        const unsigned dep_mask = pass == 0 ? lower : (valid & ~lower);
        int d_prev = u_prev >= 0 ? direction_to(u, u_prev, H) : -1;
        if (d_prev >= 0 && !((dep_mask >> d_prev) & 1u)) d_prev = -1;
        int d_next = u_next >= 0 ? direction_to(u, u_next, H) : -1;
        if (d_next >= 0 && !((send_mask >> d_next) & 1u)) d_next = -1;
        nd.u = u;
        const int nB = 2 * __builtin_popcount(lower), nF = 2 * __builtin_popcount(valid & ~lower);
        nd.gamma_den = std::max(1, std::max(nF, nB));
        nd.use_carry = d_prev >= 0;
        const int ns = 2 * __builtin_popcount(send_mask);
        nd.halves = ns > SCHED_NCW ? 2 : 1;
        int n = 0, si = 0;
        auto add_item = [&](int kind, long long term, bool tail, int strip, int need) {
            SB_REQUIRE(n < SCHED_ITEMS, SB_EUNSUP, "trws schedule: too many rows per node");
            nd.item[n].term = term;
            nd.item[n].kind = kind | (tail ? 256 : 0);
            nd.item[n].strip = strip;
            nd.item[n].need = need;
            n++;
        };
        add_item(S_D, u, false, -1, 0);
        for (int e = 0; e < 8; e++) {
            const int d = e >> 1, j = e & 1;
            if (!((valid >> d) & 1u)) continue;
            const Inc I = incidence(d, j, r, c, u, H, nV, nH);
            if ((send_mask >> d) & 1u) {
                NodeDesc::Own &o = nd.own[si & 3][si >> 2];
                o.term = I.term;
                o.flags = OWN_HAS | (I.tail ? OWN_TAIL : 0) | (d == d_next ? OWN_TO_NEXT : 0) | (j ? OWN_J : 0);
                if (rank >= 0) {
                    const int peer = s.owner[strip_of[I.nb]];
                    if (peer == rank - 1) o.flags |= OWN_PEER_UP;
                    else if (peer == rank + 1) o.flags |= OWN_PEER_DOWN;
                    else SB_REQUIRE(peer == rank, SB_EUNSUP, "trws schedule: a term spans non-adjacent ranks");
                }
                add_item(S_SEND, I.term, false, -1, 0);
                si++;
            } else if (d != d_prev) {
                add_item(S_DYN, I.term, false, strip_of[I.nb], need_of(I.nb));
            }
        }
        if (pass == 0)
            for (int e = 0; e < 8; e++) {
                const int d = e >> 1, j = e & 1;
                if (((lower >> d) & 1u) && d != d_prev) {
                    const Inc I = incidence(d, j, r, c, u, H, nV, nH);
                    add_item(S_RND, I.term, I.tail, strip_of[I.nb], need_of(I.nb));
                }
            }
        nd.nitems = n;
        // invariant: send rows and own slots are handed out together, in the same order
        {
            int t = 0;
            for (int q = 0; q < n; q++)
                if ((nd.item[q].kind & 255) == S_SEND) {
                    const NodeDesc::Own &o = nd.own[t & 3][t >> 2];
                    SB_REQUIRE((o.flags & OWN_HAS) && o.term == nd.item[q].term, SB_EUNSUP,
                               "trws schedule: send rows and own slots out of step");
                    t++;
                }
            SB_REQUIRE(t == si, SB_EUNSUP, "trws schedule: send rows and own slots out of step");
        }
    };
    auto same_structure = [&](const NodeDesc &a, const NodeDesc &b) {
        if (a.halves != b.halves || a.gamma_den != b.gamma_den || a.use_carry != b.use_carry || a.nitems != b.nitems)
            return false;
        for (int w = 0; w < SCHED_NCW; w++)
            for (int h = 0; h < 2; h++)
                if (a.own[w][h].flags != b.own[w][h].flags) return false;
        for (int q = 0; q < SCHED_ITEMS; q++)
            if (a.item[q].kind != b.item[q].kind || a.item[q].strip != b.item[q].strip) return false;
        return true;
    };

    plan.segs.clear();
    plan.seg_ptr.clear();
    plan.strips.clear();
    plan.strip_len.clear();
    NodeDesc first, nd;
    long long d_u = 0, d_own[SCHED_NCW][2], d_term[SCHED_ITEMS], d_need[SCHED_ITEMS];
    int seg_n = 0;
    auto flush = [&]() {
        if (!seg_n) return;
        Segment g;
        std::memset(&g, 0, sizeof(g));
        g.u0 = first.u; g.du = (int)d_u; g.n = seg_n;
        g.halves = first.halves; g.gamma_den = first.gamma_den; g.use_carry = first.use_carry; g.nitems = first.nitems;
        for (int w = 0; w < SCHED_NCW; w++)
            for (int h = 0; h < 2; h++) {
                g.own[w][h].term0 = first.own[w][h].term;
                g.own[w][h].tstride = (int)d_own[w][h];
                g.own[w][h].flags = first.own[w][h].flags;
            }
        for (int q = 0; q < SCHED_ITEMS; q++) {
            g.item[q].term0 = first.item[q].term;
            g.item[q].tstride = (int)d_term[q];
            g.item[q].kind = first.item[q].kind;
            g.item[q].strip = first.item[q].strip;
            g.item[q].need0 = first.item[q].need;
            g.item[q].dneed = (int)d_need[q];
        }
        plan.segs.push_back(g);
        seg_n = 0;
    };
    for (int fs = 0; fs < S; fs++) {
        if (rank >= 0 && s.owner[fs] != rank) continue;
        plan.seg_ptr.push_back((int32_t)plan.segs.size());
        plan.strips.push_back(fs);
        const int64_t sb = s.strip_ptr[fs], se = s.strip_ptr[fs + 1];
        const int64_t len = se - sb;
        plan.strip_len.push_back((int32_t)len);
        auto node_at = [&](int64_t i) { return (int)s.nodes[pass == 0 ? sb + i : se - 1 - i]; };
        for (int64_t i = 0; i < len; i++) {
            describe(node_at(i), i > 0 ? node_at(i - 1) : -1, i + 1 < len ? node_at(i + 1) : -1, nd);
            bool extend = false;
            if (seg_n >= 1 && same_structure(first, nd)) {
                extend = true;
                if (seg_n == 1) {
                    d_u = (long long)nd.u - first.u;
                    for (int w = 0; w < SCHED_NCW; w++)
                        for (int h = 0; h < 2; h++) {
                            d_own[w][h] = nd.own[w][h].term - first.own[w][h].term;
                            if (std::llabs(d_own[w][h]) > INT32_MAX) extend = false;   // strides are 32-bit
                        }
                    for (int q = 0; q < SCHED_ITEMS; q++) {
                        d_term[q] = nd.item[q].term - first.item[q].term;
                        d_need[q] = (long long)nd.item[q].need - first.item[q].need;
                        if (std::llabs(d_term[q]) > INT32_MAX) extend = false;
                    }
                } else {
                    const long long k = seg_n;
                    if ((long long)nd.u != first.u + k * d_u) extend = false;
                    for (int w = 0; w < SCHED_NCW && extend; w++)
                        for (int h = 0; h < 2; h++)
                            if (nd.own[w][h].term != first.own[w][h].term + k * d_own[w][h]) extend = false;
                    for (int q = 0; q < SCHED_ITEMS && extend; q++)
                        if (nd.item[q].term != first.item[q].term + k * d_term[q] ||
                            nd.item[q].need != first.item[q].need + k * d_need[q]) extend = false;
                }
            }
            if (extend) {
                seg_n++;
            } else {
                flush();
                first = nd;
                d_u = 0;
                std::memset(d_own, 0, sizeof(d_own));
                std::memset(d_term, 0, sizeof(d_term));
                std::memset(d_need, 0, sizeof(d_need));
                seg_n = 1;
            }
        }
        flush();
    }
    plan.seg_ptr.push_back((int32_t)plan.segs.size());
}

} // namespace sb

extern "C" {

int sb_trws_grid_ordering(int H, int W, int32_t *ordering)
{
    return sb::guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && ordering, SB_EINVAL, "sb_trws_grid_ordering: bad arguments");
        std::vector<int32_t> o;
        SB_REQUIRE(sb::grid_ordering(H, W, o), SB_EINVAL,
                   "sb_trws_grid_ordering: %dx%d grid has no valid automatic ordering "
                   "(every node's degree >= node count; the reference crashes here)", H, W);
        std::copy(o.begin(), o.end(), ordering);
    });
}

int sb_trws_plan_stats(int H, int W, int rank, int world, int64_t *stats)
{
    return sb::guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && stats && world >= 1 && rank < world, SB_EINVAL, "sb_trws_plan_stats: bad arguments");
        std::vector<int32_t> order;
        SB_REQUIRE(sb::grid_ordering(H, W, order), SB_EINVAL, "sb_trws_plan_stats: no valid ordering");
        std::vector<uint8_t> info;
        sb::build_node_info(H, W, order, info);
        sb::Schedule sched;
        sb::build_schedule(H, W, order, sched, world);
        stats[0] = (int64_t)sched.strip_ptr.size() - 1;
        for (int pass = 0; pass < 2; pass++) {
            sb::PassPlan plan;
            sb::build_pass_plan(H, W, info, sched, pass, plan, world > 1 ? rank : -1);
            int64_t nodes = 0, items = 0, two_half = 0, up = 0, down = 0;
            for (const auto &g : plan.segs) {
                nodes += g.n;
                items += (int64_t)g.n * g.nitems;
                if (g.halves == 2) two_half += g.n;
                for (int w = 0; w < sb::trws::SCHED_NCW; w++)
                    for (int h = 0; h < 2; h++) {
                        if (g.own[w][h].flags & sb::trws::OWN_PEER_UP) up += g.n;
                        if (g.own[w][h].flags & sb::trws::OWN_PEER_DOWN) down += g.n;
                    }
            }
            stats[1 + 6 * pass] = (int64_t)plan.segs.size();
            stats[2 + 6 * pass] = nodes;
            stats[3 + 6 * pass] = items;
            stats[4 + 6 * pass] = two_half;
            stats[5 + 6 * pass] = up;
            stats[6 + 6 * pass] = down;
        }
    });
}

int sb_grid_from_connectivity(int64_t N, int64_t E, const uint32_t *conn, int *H, int *W)
{
    return sb::guarded([&] {
        SB_REQUIRE(conn || E == 0, SB_EINVAL, "sb_grid_from_connectivity: null connectivity");
        SB_REQUIRE(H && W, SB_EINVAL, "sb_grid_from_connectivity: null output");
        int h = 0, w = 0;
        SB_REQUIRE(sb::grid_from_connectivity(N, E, conn, h, w), SB_ENOTGRID,
                   "connectivity (N=%lld, E=%lld) is not the 4-connected dispmap_super grid",
                   (long long)N, (long long)E);
        *H = h; *W = w;
    });
}

} // extern "C"
