// trws_sched.h -- the data-driven sweep schedule shared by the host builder
// (trws_order.cpp) and the sweep kernel (trws_kernels.cuh).
//
// A sweep visits the nodes in STRIPS (trws_order.cpp: the boundary ring, then one
// strip per interior row).  Within a strip, runs of nodes that look alike -- same
// incident-term structure, same dependencies, addresses advancing by constant
// strides -- are folded into SEGMENTS, and everything a warp needs to process the
// i-th node of a segment is `base + i * stride` of a few descriptor fields.  The
// kernel therefore does no index arithmetic on the grid at all: which terms a node
// sends on (MRFEnergy.cpp:188-219 orientation), which warp owns which term, which
// rows have to be fetched and which progress counter guards them are all decided
// here, once per problem.
#pragma once
#include <stdint.h>

namespace sb {
namespace trws {

enum { SCHED_NCW = 4, SCHED_ITEMS = 22 };

// kinds of rows the helper warps fetch for a node
enum { S_NONE = 0, S_D = 1, S_SEND = 2, S_DYN = 3, S_RND = 4 };

// own-term flags
enum { OWN_HAS = 1, OWN_TAIL = 2, OWN_TO_NEXT = 4, OWN_J = 8,
       OWN_PEER_UP = 16,     // row-banded multi-GPU: the receiving node lives on rank - 1 ...
       OWN_PEER_DOWN = 32 }; // ... or on rank + 1: the message is pushed into that GPU's memory

struct SegOwn {            // the send term a term warp owns in one half of a node (16 bytes)
    long long term0;       // term of the segment's first node
    int tstride;           // term increment per node
    int flags;             // OWN_*
};

struct SegItem {           // one row a helper warp fetches per node (32 bytes)
    long long term0;       // term (S_D: node) of the segment's first node
    int tstride;           // increment per node
    int kind;              // S_*  | tail << 8 (S_RND: am I the tail of the term)
    int strip;             // S_DYN / S_RND: progress counter that guards the row, else -1
    int need0, dneed;      // ... and the count it must have reached: need0 + i * dneed
    int pad;
};

struct Segment {           // 864 bytes, 16-byte aligned
    int u0, du, n;         // nodes u0 + i * du, i in [0, n)
    int halves;            // 2 when the nodes send on more than SCHED_NCW terms
    int gamma_den;         // max(nF, nB): gamma = 1 / gamma_den (treeProbabilities.cpp:28-45)
    int use_carry;         // the nodes receive the two messages of the previous strip node through shared memory
    int nitems;
    int pad;
    SegOwn own[SCHED_NCW][2];
    SegItem item[SCHED_ITEMS];
};

static_assert(sizeof(Segment) == 864, "Segment layout");

} // namespace trws
} // namespace sb
