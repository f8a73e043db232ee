// gtrws_plan.h -- sweep schedule of the GRID-NATIVE TRW-S path (sb_trws_grid_*, gtrws_*.cu).
//
// Same node order / orientation DAG / strips as the MATLAB-layout path (trws_order.cpp:
// SetAutomaticOrdering of cpp/trw-s/ordering.cpp:7-157, edge orientation of
// MRFEnergy.cpp:188-219), but addressed for the plane-native, row-major, band-local layout of
// gtrws_kernels.cuh: a node step is described by its local node id and a ROLE per grid direction;
// every address (neighbour node, pair record, term slot) follows from those by arithmetic.
#pragma once
#include <stdint.h>
#include <vector>

namespace sb {
namespace gtrws {

enum { DIR_UP = 0, DIR_DOWN = 1, DIR_LEFT = 2, DIR_RIGHT = 3 };
// what a node does with the neighbour pair in one direction during a pass
enum { ROLE_NONE = 0,
       ROLE_SEND0 = 1,   // sends both terms of the pair (staged in send slot 0)
       ROLE_SEND1 = 2,   // ... send slot 1
       ROLE_ADD = 3,     // a send pair handled by the node's second step: only summed into the node total here
       ROLE_POLL = 4,    // receives: waits for the neighbour's two messages of this pass (tagged words)
       ROLE_CARRY = 5 }; // receives from the previous node of the strip through shared memory

enum { GF_SAVE = 1,      // the node sends on more than two pairs: the extra pairs are only summed here (ROLE_ADD), the node
                         // total is saved to a scratch slot and the extra sends run at the end of the strip
       GF_DEFERRED = 2 };// such a deferred step: node total taken from the scratch slot, no dependencies

// A run of node steps with identical roles and a constant node stride (32 bytes).
struct GSeg {
    int32_t u0, du, n;       // local node ids u0 + i * du
    uint32_t roles;          // 4 bits per direction
    int16_t gamma_den;       // max(nF, nB): gamma = 1 / gamma_den (treeProbabilities.cpp:28-45)
    int8_t next_dir;         // direction of the next strip node when it is a send target, else -1
    uint8_t flags;           // GF_*
    uint8_t peer[4];         // per direction: 0 receiver local, 1 on the rank that owns the block to the left, 2 ... to the right
    int32_t save0;           // GF_SAVE / GF_DEFERRED: scratch slot of the segment's first node (slots advance by 1)
    int32_t pad;
};
static_assert(sizeof(GSeg) == 32, "GSeg layout");

struct GPassPlan {
    std::vector<GSeg> segs;
    std::vector<int32_t> seg_ptr;    // per strip of this rank (schedule order)
    std::vector<int32_t> strip_len;  // node steps per strip
    std::vector<int32_t> is_ring;    // the strip is (a piece of) the boundary ring
    int32_t save_slots = 0;          // scratch slots for saved node totals
};

// Storage geometry of rank `rank` of `world`.  The grid is split into COLUMN blocks of `wb` columns dealt round
// robin to the ranks (trws_order.h, build_schedule_cols): the strips of the sweep run along the image rows, so the
// ranks are the stages of a pipeline.  A rank stores each of its blocks as a grid of H rows x Wl = wb + 2 columns
// (one halo column per side, unused at the image border), block after block:
//   local node id = b * H * Wl + r * Wl + (c - B * wb + 1),   B = c / wb = rank + b * world.
// world == 1: one block = the whole grid, no halo (Wl = W).
struct Band {
    int H, W, rank, world;
    int wb, Wl, nblocks, NB;
    int block_id(int b) const { return world <= 1 ? 0 : rank + b * world; }
    int c_base(int b) const { return world <= 1 ? 0 : block_id(b) * wb - 1; }                  // first STORED column (may be -1)
    int c_lo(int b) const { return world <= 1 ? 0 : block_id(b) * wb; }                        // first owned column
    int c_hi(int b) const { const int e = world <= 1 ? W : (block_id(b) + 1) * wb; return e < W ? e : W; }
    long long nodes() const { return (long long)nblocks * H * Wl; }
};
// blocks <= 0: the default (1 for one rank; up to 4 per rank when the blocks stay at least 16 columns wide)
int default_col_blocks(int W, int world);
Band band_window(int H, int W, int rank, int world, int blocks);

// pass 0 forward, 1 backward.  rank < 0: whole grid on one GPU.  Strips are listed in PROCESSING order of the pass
// (the backward sweep runs the forward schedule in reverse); the deferred sends of a strip's nodes (GF_DEFERRED
// steps) follow its regular steps.
void build_gpass_plan(int H, int W, int pass, int rank, int world, int blocks, GPassPlan &plan);

} // namespace gtrws
} // namespace sb
