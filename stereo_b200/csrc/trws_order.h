// trws_order.h -- host-side grid / ordering / schedule helpers (trws_order.cpp).
#pragma once
#include <stdint.h>
#include <cmath>
#include <vector>
#include "sb_common.h"
#include "trws_sched.h"

namespace sb {

// Pairwise terms of the H x W grid in dispmap_super.construct_neighborhood order.
void grid_terms(int H, int W, std::vector<int32_t> &tail, std::vector<int32_t> &head);
// True (and H, W set) iff conn is exactly that grid.
bool grid_from_connectivity(int64_t N, int64_t E, const uint32_t *conn, int &H, int &W);
// ordering.cpp:24-152 on an arbitrary term list; false where the reference would crash.
bool greedy_ordering(int64_t N, const std::vector<int32_t> &tail, const std::vector<int32_t> &head,
                     std::vector<int32_t> &ordering);
// SetAutomaticOrdering on the grid (closed form for H,W >= 4).
bool grid_ordering(int H, int W, std::vector<int32_t> &ordering);
// Per-node incidence byte (valid mask | lower mask << 4).
void build_node_info(int H, int W, const std::vector<int32_t> &ordering, std::vector<uint8_t> &info);
// Strip schedule of the forward sweep (the backward sweep runs it in reverse).
struct Schedule {
    std::vector<int32_t> nodes;      // concatenated strips
    std::vector<int64_t> strip_ptr;  // strip s = nodes[strip_ptr[s] .. strip_ptr[s+1])
    std::vector<int32_t> owner;      // rank that sweeps strip s (row-banded multi-GPU; all 0 for one GPU)
    bool regular = false;            // ring + interior rows (H,W >= 4)
    int world = 1;
};
// Rank that owns image row r when the rows are split into `world` contiguous bands.
inline int band_of_row(int r, int H, int world) { return (int)(((long long)r * world) / H); }
// world > 1 (regular grids only): the ring strip is cut at the band boundaries and every strip
// gets an owner.
void build_schedule(int H, int W, const std::vector<int32_t> &ordering, Schedule &s, int world = 1);
// Column blocks of the grid-native multi-GPU sweep: blocks of `wb` columns dealt round robin to the ranks (block B
// belongs to rank B % world).  One block per rank = contiguous bands; several = block-cyclic, which shortens the
// pipeline fill (a rank waits for (world - 1) BLOCKS of the first row, not for (world - 1) / world of it).
inline int col_block_width(int W, int world, int blocks) { return (W + world * blocks - 1) / (world * blocks); }
inline int band_of_col(int c, int wb, int world) { return (c / wb) % world; }
// Column-banded variant (grid-native path, gtrws_plan.cpp): the strips run ALONG the image rows, so
// cutting every strip at the column-band boundaries turns the ranks into the stages of a pipeline -- rank g
// works on its piece of row r while rank g + 1 already works on row r + 1 -- where row bands would make
// the ranks take turns (row r + 1 waits for row r).  The ring is cut wherever its owner changes.
void build_schedule_cols(int H, int W, const std::vector<int32_t> &ordering, Schedule &s, int world, int blocks = 1);

// Segment descriptors of one pass (trws_sched.h): segment range of forward strip fs =
// [seg_ptr[fs], seg_ptr[fs + 1]).  pass 0 = forward sweep,
// 1 = backward sweep (strips visited in reverse, nodes within a strip in reverse).
struct PassPlan {
    std::vector<trws::Segment> segs;
    std::vector<int32_t> seg_ptr;     // over the strips of `strips` (the rank's own, in schedule order)
    std::vector<int32_t> strips;      // global strip ids
    std::vector<int32_t> strip_len;   // their node counts
};
// rank < 0: every strip (single GPU).  Otherwise only the strips `rank` owns, with the send terms
// whose receiver lives on a neighbouring rank flagged OWN_PEER_UP / OWN_PEER_DOWN.
void build_pass_plan(int H, int W, const std::vector<uint8_t> &info, const Schedule &s, int pass, PassPlan &plan,
                     int rank = -1);

} // namespace sb
