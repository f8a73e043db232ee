// gtrws_plan.cpp -- host-side schedule of the grid-native TRW-S sweep (no CUDA).
// Reuses the ordering / strip schedule of trws_order.cpp (the reference's
// SetAutomaticOrdering, cpp/trw-s/ordering.cpp:7-157, and the edge orientation of
// MRFEnergy.cpp:188-219) and re-expresses every node step in the role-per-direction form of
// gtrws_plan.h.
#include "gtrws_plan.h"
#include "trws_order.h"
#include <cstring>
#include <algorithm>

namespace sb {
namespace gtrws {

Band band_rows(int H, int rank, int world)
{
    Band b;
    if (rank < 0 || world <= 1) {
        b.r_lo = 0; b.r_hi = H; b.r_base = 0; b.r_top = H;
        return b;
    }
    // rows r with band_of_row(r) == rank, band_of_row(r) = floor(r * world / H)
    auto first_row = [&](int k) { return (int)(((long long)k * H + world - 1) / world); };
    b.r_lo = first_row(rank);
    b.r_hi = first_row(rank + 1);
    b.r_base = std::max(0, b.r_lo - 1);
    b.r_top = std::min(H, b.r_hi + 1);
    return b;
}

namespace {

struct Step {
    int u;            // local node id
    uint32_t roles;
    int gamma_den, next_dir, flags;
    uint8_t peer[4];
};

inline bool same_shape(const Step &a, const Step &b)
{
    return a.roles == b.roles && a.gamma_den == b.gamma_den && a.next_dir == b.next_dir && a.flags == b.flags &&
           std::memcmp(a.peer, b.peer, 4) == 0;
}

} // namespace

void build_gpass_plan(int H, int W, int pass, int rank, int world, GPassPlan &plan)
{
    SB_REQUIRE(H >= 4 && W >= 4, SB_EUNSUP, "sb_trws_grid: the grid-native path needs H, W >= 4 (got %d x %d)", H, W);
    if (rank < 0) world = 1;
    std::vector<int32_t> order;
    SB_REQUIRE(grid_ordering(H, W, order), SB_EINVAL, "sb_trws_grid: no valid ordering");
    std::vector<uint8_t> info;
    build_node_info(H, W, order, info);
    Schedule s;
    build_schedule(H, W, order, s, world);
    const Band band = band_rows(H, rank, world);
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
        SB_REQUIRE(r >= band.r_base && r < band.r_top, SB_EUNSUP, "sb_trws_grid: node outside the band's storage");
        return (r - band.r_base) * W + c;
    };

    plan.segs.clear();
    plan.seg_ptr.clear();
    plan.strip_len.clear();
    std::vector<Step> steps;
    for (int fs = 0; fs < S; fs++) {
        if (rank >= 0 && s.owner[fs] != rank) continue;
        const int64_t sb = s.strip_ptr[fs], se = s.strip_ptr[fs + 1], len = se - sb;
        auto node_at = [&](int64_t i) { return (int)s.nodes[pass == 0 ? sb + i : se - 1 - i]; };
        steps.clear();
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
            // send directions, the one towards the next strip node LAST (its messages are handed over
            // through shared memory by the step right before that node)
            int sd[4], ns = 0;
            for (int d = 0; d < 4; d++)
                if (((send_mask >> d) & 1u) && d != d_next) sd[ns++] = d;
            if (d_next >= 0) sd[ns++] = d_next;
            for (int d = 0; d < 4; d++)
                if ((send_mask >> d) & 1u) {
                    if (rank >= 0) {
                        const int peer = s.owner[strip_of[nb_ref(u, d)]];
                        if (peer == rank - 1) base.peer[d] = 1;
                        else if (peer == rank + 1) base.peer[d] = 2;
                        else SB_REQUIRE(peer == rank, SB_EUNSUP, "sb_trws_grid: a term spans non-adjacent ranks");
                    }
                }
            auto role_set = [](uint32_t &roles, int d, int role) { roles |= (uint32_t)role << (4 * d); };
            uint32_t dep_roles = 0;
            for (int d = 0; d < 4; d++)
                if ((dep_mask >> d) & 1u) role_set(dep_roles, d, d == d_prev ? ROLE_CARRY : ROLE_POLL);
            if (ns <= 2) {
                Step st = base;
                st.roles = dep_roles;
                for (int q = 0; q < ns; q++) role_set(st.roles, sd[q], ROLE_SEND0 + q);
                st.next_dir = d_next;
                steps.push_back(st);
            } else {
                Step a = base, b = base;
                a.roles = dep_roles;
                a.flags = GF_FIRST;
                role_set(a.roles, sd[0], ROLE_SEND0);
                role_set(a.roles, sd[1], ROLE_SEND1);
                for (int q = 2; q < ns; q++) role_set(a.roles, sd[q], ROLE_ADD);
                b.flags = GF_SECOND;
                for (int q = 2; q < ns; q++) role_set(b.roles, sd[q], ROLE_SEND0 + (q - 2));
                b.next_dir = d_next;
                steps.push_back(a);
                steps.push_back(b);
            }
        }
        plan.seg_ptr.push_back((int32_t)plan.segs.size());
        plan.strip_len.push_back((int32_t)steps.size());
        // fold runs of identical shape and constant stride
        size_t i = 0;
        while (i < steps.size()) {
            GSeg g;
            std::memset(&g, 0, sizeof(g));
            g.u0 = steps[i].u; g.du = 0; g.n = 1;
            g.roles = steps[i].roles; g.gamma_den = (int16_t)steps[i].gamma_den; g.next_dir = (int8_t)steps[i].next_dir;
            g.flags = (uint8_t)steps[i].flags;
            std::memcpy(g.peer, steps[i].peer, 4);
            size_t j = i + 1;
            // two-step nodes stay segments of their own (their second step repeats the node id)
            if (!steps[i].flags && j < steps.size() && same_shape(steps[i], steps[j])) {
                g.du = steps[j].u - steps[i].u;
                while (j < steps.size() && same_shape(steps[i], steps[j]) &&
                       steps[j].u == g.u0 + (int64_t)(j - i) * g.du) j++;
                g.n = (int32_t)(j - i);
            }
            plan.segs.push_back(g);
            i = j;
        }
    }
    plan.seg_ptr.push_back((int32_t)plan.segs.size());
}

} // namespace gtrws
} // namespace sb
