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
void build_schedule(int H, int W, const std::vector<int32_t> &ordering, Schedule &s)
{
    const int64_t N = (int64_t)H * W;
    s.nodes.clear();
    s.strip_ptr.clear();
    s.nodes.reserve(N);
    if (H >= 4 && W >= 4) {
        const int64_t ring = 2LL * H + 2LL * W - 4;
        std::vector<int32_t> ringnodes(ring);
        auto put = [&](int r, int c) { const int64_t u = r + (int64_t)H * c; ringnodes[ordering[u]] = (int32_t)u; };
        for (int r = 0; r < H; r++) { put(r, 0); put(r, W - 1); }
        for (int c = 1; c < W - 1; c++) { put(0, c); put(H - 1, c); }
        s.strip_ptr.push_back(0);
        s.nodes.insert(s.nodes.end(), ringnodes.begin(), ringnodes.end());
        for (int r = 1; r <= H - 2; r++) {
            s.strip_ptr.push_back((int64_t)s.nodes.size());
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
        s.regular = false;
    }
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
