// gtrws_plan.cpp -- host-side schedule of the grid-native TRW-S sweep (no CUDA).
// Reuses the ordering / strip schedule of trws_order.cpp (the reference's
// SetAutomaticOrdering, cpp/trw-s/ordering.cpp:7-157, and the edge orientation of
// MRFEnergy.cpp:188-219) and re-expresses every node step in the role-per-direction form of
// gtrws_plan.h.
#include "gtrws_plan.h"
#include "trws_order.h"
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace sb {
namespace gtrws {

int default_col_blocks(int W, int world)
{
    if (world <= 1) return 1;
    if (const char *e = getenv("SB_GTRWS_BLOCKS")) {
        const int v = atoi(e);
        if (v >= 1) return v;
    }
    // One block per rank (contiguous bands).  Measured on 8 B200s at 1980 x 2880 x 192: four blocks per rank are SLOWER
    // (30.1 vs 27.3 ms per pass) -- at that point the pass already runs at the critical path of the reference's DAG
    // (ring + one row lag per image row + one traversal of the last row), which more blocks only lengthen by their
    // extra NVLink hand-overs; with two ranks two blocks gain 3 %.
    (void)W;
    return 1;
}

Band band_window(int H, int W, int rank, int world, int blocks)
{
    Band b;
    b.H = H; b.W = W; b.rank = rank < 0 ? 0 : rank; b.world = (rank < 0 || world <= 1) ? 1 : world;
    if (b.world == 1) {
        b.wb = W; b.Wl = W; b.nblocks = 1; b.NB = 1;
        return b;
    }
    if (blocks <= 0) blocks = default_col_blocks(W, world);
    b.wb = col_block_width(W, world, blocks);
    b.Wl = b.wb + 2;
    b.NB = (W + b.wb - 1) / b.wb;
    b.nblocks = b.NB > b.rank ? (b.NB - b.rank + world - 1) / world : 0;
    return b;
}

namespace {

struct Step {
    int u;            // local node id
    uint32_t roles;
    int gamma_den, next_dir, flags, save;
    uint8_t peer[4];
};

inline bool same_shape(const Step &a, const Step &b)
{
    return a.roles == b.roles && a.gamma_den == b.gamma_den && a.next_dir == b.next_dir && a.flags == b.flags &&
           std::memcmp(a.peer, b.peer, 4) == 0;
}

} // namespace

void build_gpass_plan(int H, int W, int pass, int rank, int world, int blocks, GPassPlan &plan)
{
    SB_REQUIRE(H >= 4 && W >= 4, SB_EUNSUP, "sb_trws_grid: the grid-native path needs H, W >= 4 (got %d x %d)", H, W);
    if (rank < 0) world = 1;
    std::vector<int32_t> order;
    SB_REQUIRE(grid_ordering(H, W, order), SB_EINVAL, "sb_trws_grid: no valid ordering");
    std::vector<uint8_t> info;
    build_node_info(H, W, order, info);
    Schedule s;
    if (blocks <= 0) blocks = default_col_blocks(W, world);
    build_schedule_cols(H, W, order, s, world, blocks);
    const Band band = band_window(H, W, rank, world, blocks);
    const int Wl = band.Wl;
    const int S = (int)s.strip_ptr.size() - 1;
    const int64_t N = (int64_t)H * W;
    std::vector<int32_t> strip_of((size_t)N);
    for (int fs = 0; fs < S; fs++)
        for (int64_t k = s.strip_ptr[fs]; k < s.strip_ptr[fs + 1]; k++) strip_of[s.nodes[k]] = fs;

    // reference node id (r + H c) -> neighbour in direction d
    auto nb_ref = [&](int u, int d) { return d == DIR_UP ? u - 1 : d == DIR_DOWN ? u + 1 : d == DIR_LEFT ? u - H : u + H; };
    auto dir_to = [&](int u, int v) {
        const int dd = v - u;
        return dd == -1 ? DIR_UP : dd == 1 ? DIR_DOWN : dd == -H ? DIR_LEFT : dd == H ? DIR_RIGHT : -1;
    };
    auto local_id = [&](int u) {
        const int r = u % H, c = u / H;
        if (band.world == 1) return r * Wl + c;
        const int B = c / band.wb;
        SB_REQUIRE(B % band.world == band.rank, SB_EUNSUP, "sb_trws_grid: node outside the rank's blocks");
        return (B / band.world) * (H * Wl) + r * Wl + (c - B * band.wb + 1);
    };

    plan.segs.clear();
    plan.seg_ptr.clear();
    plan.strip_len.clear();
    plan.is_ring.clear();
    plan.save_slots = 0;
    auto emit = [&](const std::vector<Step> &steps, bool ring) {
        plan.seg_ptr.push_back((int32_t)plan.segs.size());
        plan.strip_len.push_back((int32_t)steps.size());
        plan.is_ring.push_back(ring ? 1 : 0);
        // fold runs of identical shape, constant node stride and consecutive scratch slots
        size_t i = 0;
        while (i < steps.size()) {
            GSeg g;
            std::memset(&g, 0, sizeof(g));
            g.u0 = steps[i].u; g.du = 0; g.n = 1;
            g.roles = steps[i].roles; g.gamma_den = (int16_t)steps[i].gamma_den; g.next_dir = (int8_t)steps[i].next_dir;
            g.flags = (uint8_t)steps[i].flags; g.save0 = steps[i].save;
            std::memcpy(g.peer, steps[i].peer, 4);
            size_t j = i + 1;
            if (j < steps.size() && same_shape(steps[i], steps[j]) && (!steps[i].flags || steps[j].save == steps[i].save + 1)) {
                g.du = steps[j].u - steps[i].u;
                while (j < steps.size() && same_shape(steps[i], steps[j]) &&
                       steps[j].u == g.u0 + (int64_t)(j - i) * g.du &&
                       (!steps[i].flags || steps[j].save == steps[i].save + (int)(j - i))) j++;
                g.n = (int32_t)(j - i);
            }
            plan.segs.push_back(g);
            i = j;
        }
    };
    std::vector<Step> steps, aux;
    // The rank's strips in the order its CTAs take them.  One block per rank: schedule order.  Several (block-cyclic):
    // the pieces of ONE row on a rank depend on each other through the other ranks, a piece-time apart, so in schedule
    // order (row by row) the walkers would sit on pieces that cannot start yet and the wavefront would be
    // blocks-per-rank times shallower.  They are sorted by the time they become ready instead -- row + (blocks the row
    // has to cross first) x (piece time in row staggers) -- which is still a linear extension of the dependency
    // order on every rank (both dependencies of a piece, (r, B + 1) and (r - 1, B), have a smaller key), so walkers
    // that take strips in list order and wait cannot deadlock.
    std::vector<int> mine;
    for (int fs = 0; fs < S; fs++)
        if (rank < 0 || s.owner[fs] == rank) mine.push_back(fs);
    if (band.world > 1 && band.nblocks > 1) {
        double rho = (double)band.wb / 3.0;       // a row follows the row above about three node steps behind
        if (const char *e = getenv("SB_GTRWS_RHO")) rho = atof(e);
        auto key = [&](int fs) {
            const int u = s.nodes[s.strip_ptr[fs]];
            const int r = u % H, c = u / H;
            if (r == 0 || r == H - 1 || c == 0 || c == W - 1) return -1.0;          // ring pieces first, in schedule order
            return (double)r + (double)(band.NB - 1 - c / band.wb) * rho;
        };
        std::stable_sort(mine.begin(), mine.end(), [&](int a, int b) { return key(a) < key(b); });
    }
    if (pass == 1) std::reverse(mine.begin(), mine.end());
    for (size_t k = 0; k < mine.size(); k++) {
        const int fs = mine[k];     // processing order of this pass
        const int64_t sb = s.strip_ptr[fs], se = s.strip_ptr[fs + 1], len = se - sb;
        auto node_at = [&](int64_t i) { return (int)s.nodes[pass == 0 ? sb + i : se - 1 - i]; };
        steps.clear();
        aux.clear();
        for (int64_t i = 0; i < len; i++) {
            const int u = node_at(i);
            const int u_prev = i > 0 ? node_at(i - 1) : -1, u_next = i + 1 < len ? node_at(i + 1) : -1;
            const unsigned valid = info[u] & 15u, lower = info[u] >> 4;
            const unsigned send_mask = pass == 0 ? (valid & ~lower) : lower;
            const unsigned dep_mask = pass == 0 ? lower : (valid & ~lower);
            int d_prev = u_prev >= 0 ? dir_to(u, u_prev) : -1;
            if (d_prev >= 0 && !((dep_mask >> d_prev) & 1u)) d_prev = -1;
            int d_next = u_next >= 0 ? dir_to(u, u_next) : -1;
            if (d_next >= 0 && !((send_mask >> d_next) & 1u)) d_next = -1;
            const int nB = 2 * __builtin_popcount(lower), nF = 2 * __builtin_popcount(valid & ~lower);
            Step base;
            std::memset(&base, 0, sizeof(base));
            base.u = local_id(u);
            base.gamma_den = std::max(1, std::max(nF, nB));
            base.next_dir = -1;
            // send directions by urgency: the next strip node first, then the receivers in the order this pass
            // reaches them (ascending node order forward, descending backward)
            int sd[4], ns = 0;
            if (d_next >= 0) sd[ns++] = d_next;
            {
                int cand[4], nc = 0;
                for (int d = 0; d < 4; d++)
                    if (((send_mask >> d) & 1u) && d != d_next) cand[nc++] = d;
                std::sort(cand, cand + nc, [&](int a, int b) {
                    const int32_t oa = order[nb_ref(u, a)], ob = order[nb_ref(u, b)];
                    return pass == 0 ? oa < ob : oa > ob;
                });
                for (int q = 0; q < nc; q++) sd[ns++] = cand[q];
            }
            for (int d = 0; d < 4; d++)
                if ((send_mask >> d) & 1u) {
                    if (rank >= 0) {
                        // only horizontal terms cross a block boundary; the direction names the neighbour (with two
                        // ranks the blocks to the left and to the right belong to the same one)
                        const int peer = s.owner[strip_of[nb_ref(u, d)]];
                        if (peer != rank) {
                            SB_REQUIRE(d == DIR_LEFT || d == DIR_RIGHT, SB_EUNSUP, "sb_trws_grid: a vertical term spans two ranks");
                            base.peer[d] = d == DIR_LEFT ? 1 : 2;
                        }
                    }
                }
            auto role_set = [](uint32_t &roles, int d, int role) { roles |= (uint32_t)role << (4 * d); };
            Step st = base;
            for (int d = 0; d < 4; d++)
                if ((dep_mask >> d) & 1u) role_set(st.roles, d, d == d_prev ? ROLE_CARRY : ROLE_POLL);
            // the to-next pair goes into send slot 1 (its warps also hand the carry rows over), the other into slot 0
            auto is_ring_node = [&](int v) { const int r = v % H, c = v / H; return r == 0 || r == H - 1 || c == 0 || c == W - 1; };
            bool urgent_extra = false;   // a send beyond the first two goes to an INTERIOR node: the wavefront waits for it
            for (int q = 2; q < ns; q++)
                if (!is_ring_node(nb_ref(u, sd[q]))) urgent_extra = true;
            if (ns > 2 && urgent_extra) {
                // split in place: this step sends on the two most urgent pairs that do not lead to the next strip
                // node and saves the node total; the step right behind it sends on the rest, the to-next pair included
                // (so the carry rows are still written by the step before the next node)
                SB_REQUIRE(pass == 1, SB_EUNSUP, "sb_trws_grid: a forward node sends on more than two pairs");
                int oth[4], no = 0;
                for (int q = 0; q < ns; q++)
                    if (sd[q] != d_next) oth[no++] = sd[q];
                st.flags = GF_SAVE;
                st.save = plan.save_slots++;
                role_set(st.roles, oth[0], ROLE_SEND0);
                role_set(st.roles, oth[1], ROLE_SEND1);
                Step ax = base;
                ax.flags = GF_DEFERRED;
                ax.save = st.save;
                int slot = 0;
                for (int q = 2; q < no; q++) { role_set(st.roles, oth[q], ROLE_ADD); role_set(ax.roles, oth[q], ROLE_SEND0 + slot++); }
                if (d_next >= 0) {
                    role_set(st.roles, d_next, ROLE_ADD);
                    role_set(ax.roles, d_next, slot == 0 ? ROLE_SEND0 : ROLE_SEND1);
                    ax.next_dir = d_next;
                }
                steps.push_back(st);
                steps.push_back(ax);
                continue;
            }
            const int kept = std::min(ns, 2);
            if (kept == 2) { role_set(st.roles, sd[1], ROLE_SEND0); role_set(st.roles, sd[0], ROLE_SEND1); }
            else if (kept == 1) role_set(st.roles, sd[0], ROLE_SEND0);
            st.next_dir = d_next;
            if (ns > 2) {
                // the sends beyond two all go to ring nodes, which only the far end of the pass needs: they run at
                // the end of the strip, from the saved node total
                SB_REQUIRE(pass == 1, SB_EUNSUP, "sb_trws_grid: a forward node sends on more than two pairs");
                st.flags = GF_SAVE;
                st.save = plan.save_slots++;
                for (int q = 2; q < ns; q++) role_set(st.roles, sd[q], ROLE_ADD);
                Step ax = base;
                ax.flags = GF_DEFERRED;
                ax.save = st.save;
                for (int q = 2; q < ns; q++) role_set(ax.roles, sd[q], ROLE_SEND0 + (q - 2));
                aux.push_back(ax);
            }
            steps.push_back(st);
        }
        // the deferred sends run at the END of the same strip, by the same CTA: nothing but the far end of the pass
        // (the boundary ring) waits for them, and a strip of their own would only park a CTA until its row starts
        steps.insert(steps.end(), aux.begin(), aux.end());
        {
            const int u0 = node_at(0), r0 = u0 % H, c0 = u0 / H;
            emit(steps, r0 == 0 || r0 == H - 1 || c0 == 0 || c0 == W - 1);
        }
    }
    plan.seg_ptr.push_back((int32_t)plan.segs.size());
}

} // namespace gtrws
} // namespace sb
