// trws_kernels.cuh -- sm_100a device code of the TRW-S message sweep (K5) and
// its setup kernels.  Replaces, on the GPU:
//   Minimize_TRW_S forward / backward sweeps      cpp/trw-s/minimize.cpp:31-95
//   ComputeSolutionAndEnergy (primal rounding)    cpp/trw-s/minimize.cpp:223-264
//   TypeStereoLinear::Edge::UpdateMessage         cpp/trw-s/typeStereoLinear.h:329-487
//   TypeStereoQuadratic::Edge::UpdateMessage      cpp/trw-s/typeStereoQuadratic.h:329-501
//   Edge::AddColumn / Smooth                      typeStereoLinear.h:324-327,491-518
//   the per-edge argsort of trws_mex.cpp:84-119   (rank / merge-count tables)
//
// Execution model (DESIGN.md "K5"; details above sweep_kernel):
//   * the L labels of every per-node vector are blocked over the 32 lanes of a warp
//     (label = lane*K + k, K = LP/32 values per lane in registers);
//   * work is dispatched in STRIPS (trws_order.cpp: the boundary ring, then one strip per
//     interior row) through an atomic ticket; one CTA walks a strip node by node, its four
//     term warps each owning one send term of the node, helper warps assembling everything
//     else one node ahead.  A whole sweep is ONE persistent launch with no grid barriers, and
//     because it honours the reference's orientation DAG it computes what the sequential
//     sweep computes;
//   * messages to the next node of the strip stay in shared memory; messages to other strips
//     (other SMs, other GPUs) travel as self-validating 64-bit words (value | launch epoch)
//     the receiver polls -- no flag, no fence on the dependent chain (fp32 path);
//   * the min-plus update is O(L): labels are visited in the order of their (irregular)
//     positions through iteration-invariant uint8 rank tables, the two directional distance
//     transforms are warp-shuffle scans over (offset, value) pairs -- no h - alpha*x
//     cancellation -- and each destination label looks its two bracketing sources up through
//     a precomputed merge count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "trws_sched.h"

// Timing experiments: SB_TRWS_PROFILE=1 (environment, any build) prints per-phase cycle counters of
// the sweep; the work-skipping bisection switches of SB_TRWS_DEBUG need a build with
// make EXTRA=-DSB_TRWS_INSTRUMENT=1.
#ifndef SB_TRWS_INSTRUMENT
#define SB_TRWS_INSTRUMENT 0
#endif

namespace sb {
namespace trws {

enum { PASS_FWD = 0, PASS_BWD = 1 };
enum { MODE_SEND = 1, MODE_ROUND = 2 };

// Phase counters (SB_TRWS_PROFILE) are compiled in only by  make EXTRA=-DSB_TRWS_DIAG=1 : switched off at run time
// they still issue (predicated-off instructions take issue slots) -- 10-16 % per pass on the grid sweep, measured.
#ifndef SB_TRWS_DIAG
#define SB_TRWS_DIAG 0
#endif
constexpr bool TDIAG = SB_TRWS_DIAG != 0;

template <typename REAL> struct Lim;
template <> struct Lim<float> { static __host__ __device__ float big() { return 1e30f; } };
template <> struct Lim<double> { static __host__ __device__ double big() { return 1e300; } };

template <typename REAL> struct alignas(2 * sizeof(REAL)) Pair { REAL a, b; };

template <typename REAL>
struct Problem {
    int H, W, L, LP;
    long long N, E, nV, nH;
    const REAL *D;          // [N][LP]  unary (node data, MRFEnergy.cpp:54-57)
    REAL *msg;              // [E][LP]  one message per term (MRFEnergy.h:197-200)
    const REAL *posq;       // [E][LP]  head positions   (q,     trws_mex.cpp:101-105)
    const REAL *posqp;      // [E][LP]  tail positions   (qprim, trws_mex.cpp:107-111)
    const uint8_t *rank_q;  // [E][LP]  rank of each label in sorted q
    const uint8_t *rank_qp; // [E][LP]  rank of each label in sorted qprim
    const uint8_t *cnt_q;   // [E][LP]  #{qprim <= q[l]}   clamped to 255
    const uint8_t *cnt_qp;  // [E][LP]  #{q <= qprim[l]}   clamped to 255
    const REAL *alpha;      // [E]
    REAL lambda;
    const Segment *segs;    // segment descriptors of THIS pass (trws_sched.h)
    const int32_t *seg_ptr; // [S+1] segments of each of this rank's strips (schedule order)
    const int32_t *strip_len; // [S] their node counts
    const int32_t *strip_gid; // [S] their ids in the whole schedule (index of the progress counter)
    int S;
    int world;              // > 1: row-banded over several GPUs (fp32 mailbox path only)
    REAL *peer_msg[2];      // message arrays of rank - 1 / rank + 1 (peer-mapped), mirrors of boundary terms
    unsigned long long *peer_mbox[2], *peer_selbox[2];
    int32_t *progress;      // [S] nodes of each strip completed in this pass (zeroed per launch)
    int32_t *sol;           // [N] rounded labels (0-based)
    REAL *selpos;           // [E] position of the sender's rounded label on each forward term
    unsigned long long *mbox;   // fp32 only: [E][LP] (value bits | epoch << 32) self-validating copy of
                                // every message sent across strips in the current pass
    unsigned long long *selbox; // fp32 only: [E] (selected position bits | epoch << 32)
    unsigned epoch;             // launch counter (> 0): the tag that validates mailbox words
    long long *prof;        // optional [2][8] cycle counters (SB_TRWS_PROFILE), else null
    int *ticket;            // dispatch counter (zeroed before each launch)
    double *acc;            // [0] energy  [1] lower bound (zeroed before each launch)
    int mode;               // PASS_FWD only: MODE_SEND | MODE_ROUND
    int debug;              // SB_TRWS_DEBUG bisection switches (timing experiments only; 0 in production)
    int *rec;               // SB_TRWS_RECORD: host-mapped flight recorder, [grid][8 warps][4] ints, else null
};

// ---------------------------------------------------------------- helpers


template <typename T> __device__ __forceinline__ T warp_min(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sm_100a: one-instruction warp minimum for fp32 (CREDUX.MIN.F32)
template <> __device__ __forceinline__ float warp_min<float>(float v)
{
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ int warp_min_s32(int v)
{
    int r;
    asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}

// Register <-> 32-bit word views of REAL (no address-taken locals: everything stays in registers).
template <typename REAL> struct Words;
template <> struct Words<float> {
    static constexpr int N = 1;
    static __device__ __forceinline__ void put(float &r, const int *w) { r = __int_as_float(w[0]); }
    static __device__ __forceinline__ void get(int *w, float r) { w[0] = __float_as_int(r); }
};
template <> struct Words<double> {
    static constexpr int N = 2;
    static __device__ __forceinline__ void put(double &r, const int *w) { r = __hiloint2double(w[1], w[0]); }
    static __device__ __forceinline__ void get(int *w, double r) { w[0] = __double2loint(r); w[1] = __double2hiint(r); }
};

enum { LD_CG = 0, LD_CS = 1 };

// K consecutive REALs per lane, widest aligned vector access.
template <typename REAL, int K> struct VecIO {
    static constexpr int BYTES = K * (int)sizeof(REAL);
    static constexpr int VB = (BYTES % 16 == 0) ? 16 : (BYTES % 8 == 0) ? 8 : 4;
    static constexpr int NV = BYTES / VB;   // vector accesses per lane
    static constexpr int WV = VB / 4;       // 32-bit words per access
    static constexpr int WR = Words<REAL>::N;

    template <int HOW> static __device__ __forceinline__ void load(REAL (&r)[K], const REAL *p)
    {
        int w[K * WR];
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if constexpr (VB == 16) {
                const int4 *q = reinterpret_cast<const int4 *>(p) + i;
                const int4 v = (HOW == LD_CG) ? __ldcg(q) : __ldcs(q);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            } else if constexpr (VB == 8) {
                const int2 *q = reinterpret_cast<const int2 *>(p) + i;
                const int2 v = (HOW == LD_CG) ? __ldcg(q) : __ldcs(q);
                w[2 * i] = v.x; w[2 * i + 1] = v.y;
            } else {
                const int *q = reinterpret_cast<const int *>(p) + i;
                w[i] = (HOW == LD_CG) ? __ldcg(q) : __ldcs(q);
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) Words<REAL>::put(r[k], w + k * WR);
    }
    // streaming (L2-coherent) load: messages written by other SMs in this launch
    static __device__ __forceinline__ void load_cg(REAL (&r)[K], const REAL *p) { load<LD_CG>(r, p); }
    // data never written during a sweep: evict-first streaming load
    static __device__ __forceinline__ void load_ro(REAL (&r)[K], const REAL *p) { load<LD_CS>(r, p); }

    static __device__ __forceinline__ void store(REAL *p, const REAL (&r)[K])
    {
        int w[K * WR];
#pragma unroll
        for (int k = 0; k < K; k++) Words<REAL>::get(w + k * WR, r[k]);
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if constexpr (VB == 16)
                __stcg(reinterpret_cast<int4 *>(p) + i, make_int4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
            else if constexpr (VB == 8)
                __stcg(reinterpret_cast<int2 *>(p) + i, make_int2(w[2 * i], w[2 * i + 1]));
            else
                __stcg(reinterpret_cast<int *>(p) + i, w[i]);
        }
    }
};

// K consecutive bytes per lane.
template <int K> struct ByteIO {
    static constexpr int VB = (K % 8 == 0) ? 8 : (K % 4 == 0) ? 4 : (K % 2 == 0) ? 2 : 1;
    static constexpr int NV = K / VB;
    static __device__ __forceinline__ void load(uint8_t (&r)[K], const uint8_t *p)
    {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if constexpr (VB == 8) {
                const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(p) + i);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    r[8 * i + j] = (uint8_t)(v.x >> (8 * j));
                    r[8 * i + 4 + j] = (uint8_t)(v.y >> (8 * j));
                }
            } else if constexpr (VB == 4) {
                const unsigned v = __ldcs(reinterpret_cast<const unsigned *>(p) + i);
#pragma unroll
                for (int j = 0; j < 4; j++) r[4 * i + j] = (uint8_t)(v >> (8 * j));
            } else if constexpr (VB == 2) {
                const unsigned short v = __ldcs(reinterpret_cast<const unsigned short *>(p) + i);
                r[2 * i] = (uint8_t)(v & 0xff);
                r[2 * i + 1] = (uint8_t)(v >> 8);
            } else {
                r[i] = __ldcs(p + i);
            }
        }
    }
    static __device__ __forceinline__ void store(uint8_t *p, const uint8_t (&r)[K])
    {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if constexpr (VB == 8) {
                uint2 v = make_uint2(0, 0);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v.x |= (unsigned)r[8 * i + j] << (8 * j);
                    v.y |= (unsigned)r[8 * i + 4 + j] << (8 * j);
                }
                reinterpret_cast<uint2 *>(p)[i] = v;
            } else if constexpr (VB == 4) {
                unsigned v = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) v |= (unsigned)r[4 * i + j] << (8 * j);
                reinterpret_cast<unsigned *>(p)[i] = v;
            } else if constexpr (VB == 2) {
                reinterpret_cast<unsigned short *>(p)[i] = (unsigned short)(r[2 * i] | (r[2 * i + 1] << 8));
            } else {
                p[i] = r[i];
            }
        }
    }
};

// K consecutive bytes per lane kept PACKED as loaded: unpacking at load time would make the warp
// wait for the load right after issuing it (measured: the whole L2 / HBM latency of the operand
// prefetch landed on the chain).  unpack() runs when the update needs the bytes.
template <int K> struct PackedBytes {
    static constexpr int VB = (K % 8 == 0) ? 8 : (K % 4 == 0) ? 4 : (K % 2 == 0) ? 2 : 1;
    static constexpr int NV = K / VB;
    static constexpr int NW = (VB == 8) ? 2 * NV : NV;
    unsigned w[NW];
    __device__ __forceinline__ void load(const uint8_t *p)
    {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if constexpr (VB == 8) {
                const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(p) + i);
                w[2 * i] = v.x;
                w[2 * i + 1] = v.y;
            } else if constexpr (VB == 4) {
                w[i] = __ldcs(reinterpret_cast<const unsigned *>(p) + i);
            } else if constexpr (VB == 2) {
                w[i] = __ldcs(reinterpret_cast<const unsigned short *>(p) + i);
            } else {
                w[i] = __ldcs(p + i);
            }
        }
    }
    __device__ __forceinline__ void load_shared(const uint8_t *p)
    {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            if constexpr (VB == 8) {
                const uint2 v = reinterpret_cast<const uint2 *>(p)[i];
                w[2 * i] = v.x;
                w[2 * i + 1] = v.y;
            } else if constexpr (VB == 4) {
                w[i] = reinterpret_cast<const unsigned *>(p)[i];
            } else if constexpr (VB == 2) {
                w[i] = reinterpret_cast<const unsigned short *>(p)[i];
            } else {
                w[i] = p[i];
            }
        }
    }
    __device__ __forceinline__ void unpack(uint8_t (&r)[K]) const
    {
#pragma unroll
        for (int k = 0; k < K; k++) {
            if constexpr (VB == 8) r[k] = (uint8_t)(w[k / 4] >> (8 * (k % 4)));
            else r[k] = (uint8_t)(w[k / VB] >> (8 * (k % VB)));
        }
    }
};

// Shared scratch of one warp: sorted-domain (value, position) pairs.
// Logical sorted index i in [0, LP) lives at phys(i); slot 0 and slot
// phys(LP) are the (+big, 0) sentinels for "no source on this side".
// Even K gets one pad slot per K so the blocked per-lane reads are
// bank-conflict free.
// Only K = 2, 4, 8 are padded (there the pad index i / K is a shift).  K = 6 would need a division per access -- 24 of
// them per update, ~100 integer instructions of the ~1100 a node step issues, measured -- to save a 2-way bank conflict
// on the 12 blocked accesses, so it goes unpadded like the odd K.
template <int K> __host__ __device__ constexpr bool scratch_padded() { return K % 2 == 0 && (K & (K - 1)) == 0; }
template <int K> __device__ __forceinline__ int phys(int i)
{
    if constexpr (scratch_padded<K>()) return 1 + i + i / K;
    else return 1 + i;
}
template <int K> __host__ __device__ constexpr int scratch_pairs() { return 2 + 32 * K + (scratch_padded<K>() ? 32 : 0); }

// ---------------------------------------------------------------- min-plus updates
//
// Both return vMin and leave the min-normalised new message in m[] (label
// domain).  Di is the (gamma-unscaled) node sum, m the old message on this term,
// s/rk the sender's positions and their ranks, x/cn the receiver's positions
// and merge counts.

// Truncated linear: msg[j] = min(vTrunc, min_i H_i + alpha |x_j - s_i|)
// (typeStereoLinear.h:375-480).  Equal to the reference's cone envelope except on exact ties
// h_j - h_k == alpha (q_k - q_j), where the reference drops cone k and returns values above the exact
// minimum (include/stereo_b200.h, sb_trws_update_message; tests/test_update_message_gpu.py).
// `valid`: bit k set iff this lane's k-th label is a real label (< L); the caller owns the lane <-> label map.
// FULL: every label slot of the warp is a real label (L == 32 K): the `valid` tests fold away.
template <typename REAL, int K, bool FULL = false>
__device__ __forceinline__ REAL update_linear(REAL gamma, REAL alpha, REAL lambda, unsigned valid, int lane,
                                              const REAL (&Di)[K], REAL (&m)[K], const REAL (&s)[K],
                                              const uint8_t (&rk)[K], const REAL (&x)[K],
                                              const uint8_t (&cn)[K], Pair<REAL> *P)
{
    const REAL BIG = Lim<REAL>::big();
    REAL h[K];
    REAL hmin = BIG;
#pragma unroll
    for (int k = 0; k < K; k++) {
        h[k] = (FULL || ((valid >> k) & 1u)) ? gamma * Di[k] - m[k] : BIG;
        hmin = min(hmin, h[k]);
    }
    hmin = warp_min(hmin);
    if (alpha == REAL(0)) { // typeStereoLinear.h:390-395
#pragma unroll
        for (int k = 0; k < K; k++) m[k] = REAL(0);
        return hmin;
    }
    const REAL vTrunc = hmin + alpha * lambda;
#pragma unroll
    for (int k = 0; k < K; k++) {
        Pair<REAL> t;
        t.a = h[k];
        t.b = s[k];
        P[phys<K>(rk[k])] = t;
    }
    __syncwarp();
    REAL hs[K], ss[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        Pair<REAL> t = P[phys<K>(lane * K + k)];
        hs[k] = t.a;
        ss[k] = t.b;
    }
    // Two directional distance transforms over the sorted sources, as (offset, value) scans:
    //   left-to-right  G_i = min(G_{i-1} + alpha (s_i - s_{i-1}), h_i), right-to-left likewise.
    // The two are independent; their shuffle rounds are issued together so the latencies overlap.
    REAL gl[K], cl[K], gr[K], cr[K];
    {
        REAL sprev = __shfl_up_sync(0xffffffffu, ss[K - 1], 1);
        REAL snext = __shfl_down_sync(0xffffffffu, ss[0], 1);
        if (lane == 0) sprev = ss[0];
        if (lane == 31) snext = ss[K - 1];
        REAL g = BIG, cum = REAL(0);
#pragma unroll
        for (int k = 0; k < K; k++) {
            const REAL d = alpha * (ss[k] - (k ? ss[k - 1] : sprev));
            g = min(g + d, hs[k]);
            cum += d;
            gl[k] = g;
            cl[k] = cum;
        }
        REAL DlL = cum, CL = g;
        g = BIG;
        cum = REAL(0);
#pragma unroll
        for (int k = K - 1; k >= 0; k--) {
            const REAL d = alpha * ((k < K - 1 ? ss[k + 1] : snext) - ss[k]);
            g = min(g + d, hs[k]);
            cum += d;
            gr[k] = g;
            cr[k] = cum;
        }
        REAL DlR = cum, CR = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const REAL DpL = __shfl_up_sync(0xffffffffu, DlL, o);
            const REAL CpL = __shfl_up_sync(0xffffffffu, CL, o);
            const REAL DpR = __shfl_down_sync(0xffffffffu, DlR, o);
            const REAL CpR = __shfl_down_sync(0xffffffffu, CR, o);
            if (lane >= o) {
                CL = min(CpL + DlL, CL);
                DlL = DpL + DlL;
            }
            if (lane + o < 32) {
                CR = min(CpR + DlR, CR);
                DlR = DpR + DlR;
            }
        }
        REAL VinL = __shfl_up_sync(0xffffffffu, CL, 1);
        REAL VinR = __shfl_down_sync(0xffffffffu, CR, 1);
        if (lane == 0) VinL = BIG;
        if (lane == 31) VinR = BIG;
#pragma unroll
        for (int k = 0; k < K; k++) {
            Pair<REAL> t;
            t.a = min(min(VinL + cl[k], gl[k]), min(VinR + cr[k], gr[k]));
            t.b = ss[k];
            P[phys<K>(lane * K + k)] = t;
        }
    }
    __syncwarp();
    REAL vmin = BIG;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int c = cn[k];
        Pair<REAL> lo, hi;
        if constexpr (scratch_padded<K>()) {
            lo = P[c ? phys<K>(c - 1) : 0];
            hi = P[phys<K>(c)];
        } else {        // unpadded: the sentinel in slot 0 is the element before sorted index 0
            lo = P[c];
            hi = P[c + 1];
        }
        const REAL v = min(vTrunc, min(lo.a + alpha * fabs(x[k] - lo.b), hi.a + alpha * fabs(x[k] - hi.b)));
        m[k] = v;
        if (FULL || ((valid >> k) & 1u)) vmin = min(vmin, v);
    }
    vmin = warp_min(vmin);
#pragma unroll
    for (int k = 0; k < K; k++) m[k] -= vmin;
    __syncwarp();
    return vmin;
}

// Truncated quadratic: msg[j] = min(vTrunc, min_i H_i + alpha (x_j - s_i)^2)
// (typeStereoQuadratic.h:392-496).  Sources further than sqrt(lambda) from x_j
// cost at least vTrunc, so each destination scans outward from its merge
// position in the sorted sources and stops at the truncation radius.
template <typename REAL, int K>
__device__ __forceinline__ REAL update_quadratic(REAL gamma, REAL alpha, REAL lambda, unsigned valid, int L, int lane,
                                                 const REAL (&Di)[K], REAL (&m)[K], const REAL (&s)[K],
                                                 const uint8_t (&rk)[K], const REAL (&x)[K],
                                                 const uint8_t (&cn)[K], Pair<REAL> *P)
{
    const REAL BIG = Lim<REAL>::big();
    REAL hmin = BIG;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const REAL h = ((valid >> k) & 1u) ? gamma * Di[k] - m[k] : BIG;
        hmin = min(hmin, h);
        Pair<REAL> t;
        t.a = h;
        t.b = s[k];
        P[phys<K>(rk[k])] = t;
    }
    hmin = warp_min(hmin);
    if (alpha == REAL(0)) { // typeStereoQuadratic.h:396-403
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; k++) m[k] = REAL(0);
        return hmin;
    }
    const REAL vTrunc = hmin + alpha * lambda;
    __syncwarp();
    REAL vmin = BIG;
#pragma unroll
    for (int k = 0; k < K; k++) {
        REAL best = vTrunc;
        if ((valid >> k) & 1u) {
            const int c = cn[k];
            const REAL xk = x[k];
            for (int j = c - 1; j >= 0; j--) {
                const Pair<REAL> t = P[phys<K>(j)];
                const REAL d = xk - t.b;
                const REAL dd = d * d;
                if (dd >= lambda) break;
                best = min(best, t.a + alpha * dd);
            }
            for (int j = c; j < L; j++) {
                const Pair<REAL> t = P[phys<K>(j)];
                const REAL d = t.b - xk;
                const REAL dd = d * d;
                if (dd >= lambda) break;
                best = min(best, t.a + alpha * dd);
            }
            vmin = min(vmin, best);
        }
        m[k] = best;
    }
    vmin = warp_min(vmin);
#pragma unroll
    for (int k = 0; k < K; k++) m[k] -= vmin;
    __syncwarp();
    return vmin;
}

// Smooth() of typeStereoLinear.h:324-327 / typeStereoQuadratic.h:324-327 without alpha.
template <typename REAL, int KERN> __device__ __forceinline__ REAL smooth(REAL d, REAL lambda)
{
    if constexpr (KERN == 1) return min(fabs(d), lambda);
    else return min(d * d, lambda);
}

// ---------------------------------------------------------------- small PTX helpers

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// Progress watermarks: the producer fences then stores its strip's count, consumers poll
// with a relaxed (L2) load and read the published data with L2-coherent accesses only
// (ld.cg / cp.async.cg).
__device__ __forceinline__ int ld_flag(const int32_t *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void publish_flag(int32_t *p, int v)
{
    asm volatile("fence.acq_rel.gpu;\n\tst.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Mailbox words: one naturally aligned 64-bit access = value + tag, single-copy atomic, so a
// matching tag proves the value without any fence or flag.
__device__ __forceinline__ unsigned long long ld_mbox(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_mbox(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// system-scope variants: the word is written by one GPU into another GPU's memory over NVLink
__device__ __forceinline__ unsigned long long ld_mbox_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_mbox_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ int ld_acquire_cta(int *smem)
{
    const int v = atomicAdd(smem, 0);   // atomic read of the completion counter ...
    __threadfence_block();              // ... ordered before what follows (acquire at CTA scope)
    return v;
}

// One LP-row global -> shared, 16 bytes per lane per round (LDGSTS, L2-coherent).
template <typename REAL, int K> __device__ __forceinline__ void row_async(REAL *dst, const REAL *src, int lane)
{
    constexpr int CH = 32 * K * (int)sizeof(REAL) / 16;
#pragma unroll
    for (int ch = lane; ch < CH; ch += 32)
        cp_async16(reinterpret_cast<char *>(dst) + ch * 16, reinterpret_cast<const char *>(src) + ch * 16);
}

// K consecutive REALs per lane from / to a shared-memory row.
template <typename REAL, int K> __device__ __forceinline__ void row_lds(REAL (&r)[K], const REAL *row, int lane)
{
    constexpr int BYTES = K * (int)sizeof(REAL);
    if constexpr (BYTES % 16 == 0) {
        const float4 *q = reinterpret_cast<const float4 *>(row + lane * K);
#pragma unroll
        for (int i = 0; i < BYTES / 16; i++) {
            const float4 v = q[i];
            if constexpr (sizeof(REAL) == 4) {
                r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
            } else {
                r[2 * i] = __hiloint2double(__float_as_int(v.y), __float_as_int(v.x));
                r[2 * i + 1] = __hiloint2double(__float_as_int(v.w), __float_as_int(v.z));
            }
        }
    } else if constexpr (BYTES % 8 == 0) {
        const float2 *q = reinterpret_cast<const float2 *>(row + lane * K);
#pragma unroll
        for (int i = 0; i < BYTES / 8; i++) {
            const float2 v = q[i];
            if constexpr (sizeof(REAL) == 4) {
                r[2 * i] = v.x; r[2 * i + 1] = v.y;
            } else {
                r[i] = __hiloint2double(__float_as_int(v.y), __float_as_int(v.x));
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; k++) r[k] = row[lane * K + k];
    }
}
template <typename REAL, int K> __device__ __forceinline__ void row_sts(REAL *row, const REAL (&r)[K], int lane)
{
    constexpr int BYTES = K * (int)sizeof(REAL);
    if constexpr (BYTES % 16 == 0) {
        float4 *q = reinterpret_cast<float4 *>(row + lane * K);
#pragma unroll
        for (int i = 0; i < BYTES / 16; i++) {
            float4 v;
            if constexpr (sizeof(REAL) == 4) {
                v.x = r[4 * i]; v.y = r[4 * i + 1]; v.z = r[4 * i + 2]; v.w = r[4 * i + 3];
            } else {
                v.x = __int_as_float(__double2loint(r[2 * i])); v.y = __int_as_float(__double2hiint(r[2 * i]));
                v.z = __int_as_float(__double2loint(r[2 * i + 1])); v.w = __int_as_float(__double2hiint(r[2 * i + 1]));
            }
            q[i] = v;
        }
    } else if constexpr (BYTES % 8 == 0) {
        float2 *q = reinterpret_cast<float2 *>(row + lane * K);
#pragma unroll
        for (int i = 0; i < BYTES / 8; i++) {
            float2 v;
            if constexpr (sizeof(REAL) == 4) {
                v.x = r[2 * i]; v.y = r[2 * i + 1];
            } else {
                v.x = __int_as_float(__double2loint(r[i])); v.y = __int_as_float(__double2hiint(r[i]));
            }
            q[i] = v;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; k++) row[lane * K + k] = r[k];
    }
}

// ---------------------------------------------------------------- the sweep
//
// CTA = 4 term warps + 2 helper warps + publisher warp + prefetch warp; a CTA walks one strip at a time,
// driven by the segment descriptors of trws_sched.h (no grid arithmetic here).
//
//   helper warps (alternating nodes, one node ahead of the term warps) fetch every row a
//     node needs besides the term warps' own operands -- the unary row, the old messages of
//     its send terms, the dependency messages published by other strips (after the strip's
//     progress watermark says so), the position rows and selected positions for the primal
//     rounding -- with cp.async into landing buffers, and reduce them to three rows in shared
//     memory: BASE (D + every incident message except the carried pair), DIB0 (D + pairwise
//     columns of the rounded lower neighbours) and RMS (sum of the forward messages);
//   term warps: warp w owns send term w of the node (operands loaded into registers one node
//     ahead).  Per node the dependent chain is only: Di = BASE + the two messages the CTA sent
//     to this node in the previous step (carry rows in shared memory; minimize.cpp:38-46 /
//     69-77), the primal rounding when the pass carries it (minimize.cpp:240-260), the
//     min-plus update of the own term, its store, and the carry rows for the next node;
//   the publisher warp watches the term warps' completion counters and publishes the strip's
//     progress watermark (gpu-scope fence + store: a thousand cycles or more, kept off the chain);
//   the prefetch warp pulls the operands of the nodes ahead into L2.
//
// Rows are double buffered by node parity.  Named barriers: FULL[par] (helper arrives, term
// warps sync: rows of the node ready and all term warps done with the previous node) and
// EMPTY[par] (term warps arrive after reading, helper syncs before overwriting).

constexpr int NCW = SCHED_NCW;      // term warps
// Helper warps in the CTA (node i is prepared by helper i % NHW).  Four helpers were tried (a
// helper needs ~4.4 k cycles per node) and measured SLOWER: the term warps' own chain, not the
// helpers, bounds the step, and the extra named barriers cost a CTA per SM.
constexpr int NHW_MAX = 2;
__host__ __device__ constexpr int cta_threads(int nhw) { return (NCW + nhw + 2) * 32; }   // + publisher warp + prefetch warp
constexpr int PF_DIST = 6;

// shared rows: NHW sets of {BASE, DIB0, RMS} (by node % NHW), then 2 sets of {CM[2], CC[2]} (by
// node parity), then NHW x SCHED_ITEMS landing rows
enum { R_BASE = 0, R_DIB0 = 1, R_RMS = 2, R_CM = 0, R_CC = 2 };
constexpr int ROWS_CARRY = 8;
enum { BAR_FULL = 1, BAR_EMPTY = 1 + NHW_MAX };   // + node % NHW

template <typename REAL, int K, int NHW> __host__ __device__ constexpr size_t sweep_smem_bytes()
{
    return (size_t)(3 * NHW + ROWS_CARRY + NHW * SCHED_ITEMS) * 32 * K * sizeof(REAL) +
           (size_t)NCW * scratch_pairs<K>() * sizeof(Pair<REAL>);
}

// Named barriers with compile-time ids (a register id would make ptxas reserve all 16 hardware
// barriers and pin the kernel to one CTA per SM): 0 = __syncthreads, 1..2 FULL, 3..4 EMPTY.
template <int ID> __device__ __forceinline__ void named_sync_id()
{
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"((NCW + 1) * 32) : "memory");
}
template <int ID> __device__ __forceinline__ void named_arrive_id()
{
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"((NCW + 1) * 32) : "memory");
}
static_assert(NHW_MAX == 2, "barrier dispatch below assumes two helper warps");
__device__ __forceinline__ void named_sync(int id)
{
    if (id == 1) named_sync_id<1>();
    else if (id == 2) named_sync_id<2>();
    else if (id == 3) named_sync_id<3>();
    else named_sync_id<4>();
}
__device__ __forceinline__ void named_arrive(int id)
{
    if (id == 1) named_arrive_id<1>();
    else if (id == 2) named_arrive_id<2>();
    else if (id == 3) named_arrive_id<3>();
    else named_arrive_id<4>();
}

template <typename REAL, int K> struct OwnTerm {
    REAL m[K], s[K], x[K];
    PackedBytes<K> rkp, cnp;   // rank / merge-count bytes, unpacked when the update runs
    REAL alpha;
    long long term;
    int flags;       // OWN_*
};

// fp32, up to 128 labels: capped at 102 registers so that two CTAs share an SM (large grids are
// bound by the number of strips in flight)
template <typename REAL, int K, int KERN, int PASS, int NHW>
__global__ void __launch_bounds__(cta_threads(NHW), (sizeof(REAL) == 4 && K <= 4) ? 2 : 1) sweep_kernel(const Problem<REAL> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    __shared__ int s_wdone[NCW]; // nodes of the current strip each term warp has completed
    __shared__ __align__(16) Segment s_seg[NHW];
    constexpr int LP = 32 * K;
    // fp32: cross-strip messages travel through self-validating mailbox words; fp64 (parity
    // instantiation): progress watermarks published behind a gpu-scope fence
    constexpr bool MBOX = (sizeof(REAL) == 4);
    constexpr int ROWS_SETS = 3 * NHW;
    constexpr int ROWS_FIXED = ROWS_SETS + ROWS_CARRY;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool is_term = warp < NCW;
    const bool is_helper = warp >= NCW && warp < NCW + NHW;
    REAL *rows = reinterpret_cast<REAL *>(smem_raw);
    const REAL BIG = Lim<REAL>::big();
    const bool do_send = (PASS == PASS_BWD) || (p.mode & MODE_SEND);
    const bool do_round = (PASS == PASS_FWD) && (p.mode & MODE_ROUND);
    const Segment *const segs = p.segs;
    double acc_energy = 0.0, acc_lb = 0.0;
    auto set_ptr = [&](int set, int r) -> REAL * { return rows + (size_t)(set * 3 + r) * LP; };
    auto carry_ptr = [&](int par, int r) -> REAL * { return rows + (size_t)(ROWS_SETS + par * 4 + r) * LP; };
    Pair<REAL> *P = reinterpret_cast<Pair<REAL> *>(rows + (size_t)(ROWS_FIXED + NHW * SCHED_ITEMS) * LP) +
                    (size_t)(is_term ? warp : 0) * scratch_pairs<K>();
    if (is_term && lane == 0) {
        Pair<REAL> t;
        t.a = BIG;
        t.b = REAL(0);
        P[0] = t;
        P[phys<K>(LP)] = t;
    }

    for (;;) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1);
        __syncthreads(); // everybody is done with the previous strip (the auxiliary warp has published it)
        const int ts = s_ticket;
        if (threadIdx.x < NCW) s_wdone[threadIdx.x] = 0;
        __syncthreads();
        if (ts >= p.S) break;
        // the backward sweep runs the forward schedule in reverse
        const int fs = (PASS == PASS_BWD) ? p.S - 1 - ts : ts;
        const int sg0 = __ldg(p.seg_ptr + fs), sg1 = __ldg(p.seg_ptr + fs + 1);
        const int n_nodes = __ldg(p.strip_len + fs);
        const int gid = __ldg(p.strip_gid + fs);
        if (sg1 <= sg0 || n_nodes <= 0) continue;

        if (is_term) {
            // ============================================================ term warps
            const int w = warp;
            int sg = sg0;
            const Segment *g = segs + sg;
            int seg_n = __ldg(&g->n), seg_i = 0;
            int u0 = __ldg(&g->u0), du = __ldg(&g->du), halves = __ldg(&g->halves), use_carry = __ldg(&g->use_carry);
            REAL gamma = REAL(1) / REAL(__ldg(&g->gamma_den));
            SegOwn so0 = g->own[w][0], so1 = g->own[w][1];

            auto load_own = [&](const SegOwn &so, int i, OwnTerm<REAL, K> &o) {
                o.flags = so.flags;
                if (so.flags & OWN_HAS) {
                    o.term = so.term0 + (long long)i * so.tstride;
                    const long long off = o.term * LP + lane * K;
                    const bool tail = (so.flags & OWN_TAIL) != 0;
                    // sender's positions: qprim if I am the tail of the term, else q
                    // (typeStereoLinear.h:343-357 with Swap(), MRFEnergy.cpp:200-203)
                    VecIO<REAL, K>::load_cg(o.m, p.msg + off);
                    VecIO<REAL, K>::load_ro(o.s, (tail ? p.posqp : p.posq) + off);
                    VecIO<REAL, K>::load_ro(o.x, (tail ? p.posq : p.posqp) + off);
                    o.rkp.load((tail ? p.rank_qp : p.rank_q) + off);
                    o.cnp.load((tail ? p.cnt_q : p.cnt_qp) + off);
                    o.alpha = __ldg(p.alpha + o.term);
                }
            };
            // one min-plus update + stores (phase B5 of the design notes)
            auto send = [&](OwnTerm<REAL, K> &o, const REAL (&Di)[K], int xs, int par, REAL gamma) {
                if (!(o.flags & OWN_HAS)) return;
                const bool to_next = (o.flags & OWN_TO_NEXT) != 0;
                const int oj = (o.flags & OWN_J) ? 1 : 0;
                // receiver on a neighbouring GPU: 0 = rank - 1, 1 = rank + 1, -1 = local
                const int peer = (o.flags & OWN_PEER_UP) ? 0 : (o.flags & OWN_PEER_DOWN) ? 1 : -1;
                if (do_round) {
                    // position of the rounded label on this term, for the receiver's rounding
                    REAL sv = o.s[0];
#pragma unroll
                    for (int k = 1; k < K; k++)
                        if (k == xs % K) sv = o.s[k];
                    sv = __shfl_sync(0xffffffffu, sv, xs / K);
                    if constexpr (MBOX) {
                        if (lane == 0 && !to_next) {
                            const unsigned long long wv = (unsigned long long)(unsigned)__float_as_int((float)sv) | ((unsigned long long)p.epoch << 32);
                            if (peer >= 0) st_mbox_sys(p.peer_selbox[peer] + o.term, wv);
                            else st_mbox(p.selbox + o.term, wv);
                        }
                    } else {
                        if (lane == 0) __stcg(p.selpos + o.term, sv);
                    }
                    if (to_next) {
                        REAL cc[K];
#pragma unroll
                        for (int k = 0; k < K; k++) cc[k] = o.alpha * smooth<REAL, KERN>(o.x[k] - sv, p.lambda);
                        row_sts<REAL, K>(carry_ptr(par ^ 1, R_CC + oj), cc, lane);
                    }
                }
                if (do_send) {
                    REAL vmin;
                    uint8_t rk[K], cn[K];
                    o.rkp.unpack(rk);
                    o.cnp.unpack(cn);
                    unsigned valid = 0;
#pragma unroll
                    for (int k = 0; k < K; k++) valid |= (lane * K + k < p.L ? 1u : 0u) << k;
                    if constexpr (KERN == 1) {
                        if (p.L == LP) vmin = update_linear<REAL, K, true>(gamma, o.alpha, p.lambda, valid, lane, Di, o.m, o.s, rk, o.x, cn, P);
                        else vmin = update_linear<REAL, K, false>(gamma, o.alpha, p.lambda, valid, lane, Di, o.m, o.s, rk, o.x, cn, P);
                    } else
                        vmin = update_quadratic<REAL, K>(gamma, o.alpha, p.lambda, valid, p.L, lane, Di, o.m, o.s, rk, o.x, cn, P);
                    if constexpr (MBOX) {
                        if (!to_next) {   // the receiver is in another strip: it polls these words
                            if (peer >= 0) {
                                // ... on another GPU: push the words (and the mirror of the message row, read
                                // as "old message" in the next pass) into its memory
                                unsigned long long *mb = p.peer_mbox[peer] + o.term * LP + lane * K;
#pragma unroll
                                for (int k = 0; k < K; k++)
                                    st_mbox_sys(mb + k, (unsigned long long)(unsigned)__float_as_int((float)o.m[k]) | ((unsigned long long)p.epoch << 32));
                                VecIO<REAL, K>::store(p.peer_msg[peer] + o.term * LP + lane * K, o.m);
                            } else {
                                unsigned long long *mb = p.mbox + o.term * LP + lane * K;
#pragma unroll
                                for (int k = 0; k < K; k++)
                                    st_mbox(mb + k, (unsigned long long)(unsigned)__float_as_int((float)o.m[k]) | ((unsigned long long)p.epoch << 32));
                            }
                        }
                    }
                    VecIO<REAL, K>::store(p.msg + o.term * LP + lane * K, o.m);
                    if (PASS == PASS_BWD) acc_lb += (double)vmin;
                    if (to_next) row_sts<REAL, K>(carry_ptr(par ^ 1, R_CM + oj), o.m, lane);
                }
            };

            // optional phase timers (SB_TRWS_PROFILE): term warp 0 -> prof[0..3] = wait FULL, read rows +
            // rounding, update + stores, nodes
            const bool prof_on = TDIAG && (p.prof != nullptr) && w == 0;
            long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tclk = prof_on ? clock64() : 0;
            auto tick = [&](int which) {
                if (prof_on) {
                    const long long now = clock64();
                    tp[which] += now - tclk;
                    tclk = now;
                }
            };
            OwnTerm<REAL, K> own;
            load_own(so0, 0, own);
            // EMPTY barriers start "armed": nothing has to be read before the helpers' first writes
#pragma unroll
            for (int hq = 0; hq < NHW; hq++) named_arrive(BAR_EMPTY + hq);

            for (int node = 0; node < n_nodes; node++) {
                const int par = node & 1;
                const int set = node % NHW;
                const int u = u0 + seg_i * du;
                const REAL cur_gamma = gamma;
                const int cur_halves = halves;
                const bool cur_carry = use_carry != 0;
                const SegOwn cur_so1 = so1;
                const int cur_i = seg_i;
                // ---- where the next node lives (descriptor fields of a new segment are fetched here,
                // long before they are needed)
                const bool has_next = node + 1 < n_nodes;
                if (has_next) {
                    if (seg_i + 1 < seg_n) {
                        seg_i++;
                    } else {
                        sg++;
                        g = segs + sg;
                        seg_n = __ldg(&g->n);
                        seg_i = 0;
                        u0 = __ldg(&g->u0); du = __ldg(&g->du); halves = __ldg(&g->halves); use_carry = __ldg(&g->use_carry);
                        gamma = REAL(1) / REAL(__ldg(&g->gamma_den));
                        so0 = g->own[w][0];
                        so1 = g->own[w][1];
                    }
                }
                // ---- operands of the next node: issued BEFORE the barrier (in rows the warp would only wait
                // there), a whole node ahead of their use.  (Issued after the update instead -- straight into
                // the registers it released -- the loop-carried copies wait for the loads: measured +20 %.)
                OwnTerm<REAL, K> nxt;
                nxt.flags = 0;
                if (has_next && cur_halves != 2) load_own(so0, seg_i, nxt);
                // ---- rows of this node are ready, every term warp has finished the previous node
                tick(2);
                named_sync(BAR_FULL + set);
                REAL Di[K];
                int xs = 0;
                {
                    REAL base[K];
                    row_lds<REAL, K>(base, set_ptr(set, R_BASE), lane);
                    if (prof_on) { tclk += (long long)(base[0] != base[0]); tick(0); }
#pragma unroll
                    for (int k = 0; k < K; k++) Di[k] = base[k];
                    if (cur_carry && do_send) {
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            REAL v[K];
                            row_lds<REAL, K>(v, carry_ptr(par, R_CM + jj), lane);
#pragma unroll
                            for (int k = 0; k < K; k++) Di[k] += v[k];
                        }
                    }
                    if (do_round) {
                        // minimize.cpp:240-260: DiB = D + sum_{lower nb} V(x_nb, .), Dr = DiB + forward messages
                        REAL dib[K], rms[K];
                        row_lds<REAL, K>(dib, set_ptr(set, R_DIB0), lane);
                        row_lds<REAL, K>(rms, set_ptr(set, R_RMS), lane);
                        if (cur_carry) {
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                REAL v[K];
                                row_lds<REAL, K>(v, carry_ptr(par, R_CC + jj), lane);
#pragma unroll
                                for (int k = 0; k < K; k++) dib[k] += v[k];
                            }
                        }
                        // Vector::ComputeMin: first minimum in label order (typeStereoLinear.h:238-252)
                        REAL best = BIG;
                        int bi = 0x7fffffff;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const int lbl = lane * K + k;
                            const REAL dr = dib[k] + rms[k];
                            if (lbl < p.L && dr < best) { best = dr; bi = lbl; }
                        }
                        const REAL wbest = warp_min(best);
                        bi = warp_min_s32(best == wbest ? bi : 0x7fffffff);
                        xs = bi;
                        if (w == 0) {
                            REAL dv = dib[0];
#pragma unroll
                            for (int k = 1; k < K; k++)
                                if (k == bi % K) dv = dib[k];
                            dv = __shfl_sync(0xffffffffu, dv, bi / K);
                            if (lane == 0) {
                                p.sol[u] = bi;
                                acc_energy += (double)dv;
                            }
                        }
                    }
                }
                named_arrive(BAR_EMPTY + set);     // this warp has read the node's rows
                tick(1);
                if (do_send && PASS == PASS_BWD) {
                    // ComputeAndSubtractMin + lower bound (minimize.cpp:79-81)
                    REAL vmin = BIG;
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (lane * K + k < p.L) vmin = min(vmin, Di[k]);
                    vmin = warp_min(vmin);
#pragma unroll
                    for (int k = 0; k < K; k++) Di[k] -= vmin;
                    if (w == 0) acc_lb += (double)vmin;
                }
                // ---- loads of the next step, then the update of this one
                if (cur_halves == 2) {
                    // (rare: a node that sends on more than four terms) second-half operands, then the first
                    // half's update, then the next node's operands
                    load_own(cur_so1, cur_i, nxt);
                    send(own, Di, xs, par, cur_gamma);
                    own = nxt;
                    nxt.flags = 0;
                    if (has_next) load_own(so0, seg_i, nxt);
                }
                tick(4);
                tick(5);
                send(own, Di, xs, par, cur_gamma);
                if (prof_on) { tclk += (long long)(own.m[0] != own.m[0]); tick(6); }
                own = nxt;
                // this warp's stores for the node are issued: tell the auxiliary warp
                if constexpr (MBOX) {
                    // completion counter: only paces the prefetch warp -- a plain shared store
                    if (lane == 0) *reinterpret_cast<volatile int *>(s_wdone + w) = node + 1;
                } else {
                    __syncwarp();
                    if (lane == 0) {
                        // watermark mode: the counter also releases this warp's stores to the publisher warp
                        __threadfence_block();
                        atomicExch(s_wdone + w, node + 1);
                    }
                }
                if (prof_on) tp[3]++;
            }
            if (prof_on && lane == 0) {
                tick(2);
                const int grp = (fs == 0) ? 0 : 1;
                for (int q = 0; q < 4; q++) atomicAdd((unsigned long long *)p.prof + grp * 16 + q, (unsigned long long)tp[q]);
                // second bank: [32 + 8 grp + ..] = first-half send, operand prefetch issue, update + stores
                for (int q = 4; q < 7; q++) atomicAdd((unsigned long long *)p.prof + 32 + grp * 8 + (q - 4), (unsigned long long)tp[q]);
            }
            continue;
        }

        if (is_helper) {
            // ============================================================ helper warps
            const int hid = warp - NCW;
            if (hid >= NHW) continue;
            REAL *landing = rows + (size_t)ROWS_FIXED * LP + (size_t)hid * SCHED_ITEMS * LP;
            Segment *sd = &s_seg[hid];
            int sg = sg0, seg_start = 0;    // segment of the current node and the strip index of its first node
            int loaded = -1;
            int seg_n = __ldg(&segs[sg].n);
            const bool prof_on = TDIAG && (p.prof != nullptr) && hid == 0;
            long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tclk = prof_on ? clock64() : 0;
            auto tick = [&](int which) {
                if (prof_on) {
                    const long long now = clock64();
                    tp[which] += now - tclk;
                    tclk = now;
                }
            };
            for (int node = hid; node < n_nodes; node += NHW) {
                tick(6);
                while (node >= seg_start + seg_n) {
                    seg_start += seg_n;
                    sg++;
                    seg_n = __ldg(&segs[sg].n);
                }
                if (loaded != sg) {
                    // this helper's copy of the descriptor (864 bytes)
                    __syncwarp();
                    for (int ch = lane; ch < (int)sizeof(Segment) / 16; ch += 32)
                        cp_async16(reinterpret_cast<char *>(sd) + ch * 16, reinterpret_cast<const char *>(segs + sg) + ch * 16);
                    cp_async_wait_all();
                    __syncwarp();
                    loaded = sg;
                }
                tick(0);
                const int i = node - seg_start;
                const int nitems = sd->nitems;
                // lane j manages item j: address, guard, scalars, and copies its whole row
                int kind = S_NONE, strip = -1, need = 0, fv = 0;
                long long term = 0;
                REAL al = REAL(0), sel = REAL(0);
                if (lane < nitems) {
                    const SegItem it = sd->item[lane];
                    kind = it.kind;
                    const int k0 = kind & 255;
                    if ((k0 == S_DYN && !do_send) || (k0 == S_RND && !do_round)) kind = S_NONE;
                    term = it.term0 + (long long)i * it.tstride;
                    strip = it.strip;
                    need = it.need0 + i * it.dneed;
                }
                constexpr int CH = LP * (int)sizeof(REAL) / 16;
                REAL *my_row = landing + (size_t)lane * LP;
                {
                    const int k0 = kind & 255;
                    const REAL *src = nullptr;
                    if (k0 == S_D) src = p.D + term * LP;
                    else if (k0 == S_SEND) src = p.msg + term * LP;
                    else if (k0 == S_RND) src = ((kind & 256) ? p.posqp : p.posq) + term * LP;
                    if (src) {
#pragma unroll 4
                        for (int ch = 0; ch < CH; ch++)
                            cp_async16(reinterpret_cast<char *>(my_row) + ch * 16, reinterpret_cast<const char *>(src) + ch * 16);
                    }
                    if (k0 == S_RND) al = __ldg(p.alpha + term);
                    if constexpr (!MBOX)
                        if (k0 == S_DYN || k0 == S_RND) fv = ld_flag(p.progress + strip);   // checked after the static part
                }
                tick(1);
                cp_async_wait_all();
                __syncwarp();
                tick(3);
                // static part of BASE / DIB0 / RMS
                REAL base[K], dib[K], rms[K];
#pragma unroll
                for (int k = 0; k < K; k++) { base[k] = REAL(0); dib[k] = REAL(0); rms[k] = REAL(0); }
                const unsigned m_d = __ballot_sync(0xffffffffu, (kind & 255) == S_D);
                const unsigned m_send = __ballot_sync(0xffffffffu, (kind & 255) == S_SEND);
                const unsigned m_dyn = __ballot_sync(0xffffffffu, (kind & 255) == S_DYN);
                const unsigned m_rnd = __ballot_sync(0xffffffffu, (kind & 255) == S_RND);
                // (bit scans: only the rows that exist are visited -- these loops sit on the hand-over path)
                for (unsigned rem_s = m_d | m_send; rem_s; rem_s &= rem_s - 1) {
                    const int j = __ffs(rem_s) - 1;
                    REAL v[K];
                    row_lds<REAL, K>(v, landing + (size_t)j * LP, lane);
#pragma unroll
                    for (int k = 0; k < K; k++) base[k] += v[k];
                    if ((m_d >> j) & 1u) {
#pragma unroll
                        for (int k = 0; k < K; k++) dib[k] += v[k];
                    } else {
#pragma unroll
                        for (int k = 0; k < K; k++) rms[k] += v[k];
                    }
                }
                if (prof_on) { tclk += (long long)(base[0] != base[0]); tick(4); }
                if constexpr (MBOX) {
                    // dependencies through mailboxes: every lane polls the K words of its own labels
                    // (value | epoch); the managing lane of a rounding item polls the sender's selected
                    // position.  A matching epoch proves the word was written in this pass.
                    const unsigned ep = p.epoch;
                    const bool sys_scope = p.world > 1;
                    // all polls of a batch are in flight together: the selected positions (managing
                    // lanes) and up to G dependency rows (every lane its own K words)
                    constexpr int G = (K <= 2) ? 4 : (K <= 4) ? 2 : 1;
                    const bool need_sel = (kind & 255) == S_RND;
                    bool have_sel = !need_sel;
                    unsigned rem = m_dyn;
                    do {
                        long long tj[G];
                        int nb = 0;
#pragma unroll
                        for (int q = 0; q < G; q++) {
                            tj[q] = 0;
                            if (rem) {
                                const int j = __ffs(rem) - 1;
                                rem &= rem - 1;
                                tj[q] = __shfl_sync(0xffffffffu, term, j);
                                nb = q + 1;
                            }
                        }
                        unsigned long long wv[G][K];
                        for (;;) {
                            unsigned long long sw = 0;
                            if (!have_sel) sw = sys_scope ? ld_mbox_sys(p.selbox + term) : ld_mbox(p.selbox + term);
#pragma unroll
                            for (int q = 0; q < G; q++)
                                if (q < nb) {
#pragma unroll
                                    for (int k = 0; k < K; k++)
                                        wv[q][k] = sys_scope ? ld_mbox_sys(p.mbox + tj[q] * LP + lane * K + k)
                                                             : ld_mbox(p.mbox + tj[q] * LP + lane * K + k);
                                }
                            bool ok = true;
                            if (!have_sel) {
                                if ((unsigned)(sw >> 32) == ep) { sel = (REAL)__int_as_float((int)(unsigned)sw); have_sel = true; }
                                else ok = false;
                            }
#pragma unroll
                            for (int q = 0; q < G; q++)
                                if (q < nb) {
#pragma unroll
                                    for (int k = 0; k < K; k++) ok = ok && ((unsigned)(wv[q][k] >> 32) == ep);
                                }
                            if (__all_sync(0xffffffffu, ok)) break;
                            __nanosleep(20);
                        }
#pragma unroll
                        for (int q = 0; q < G; q++)
                            if (q < nb) {
#pragma unroll
                                for (int k = 0; k < K; k++) base[k] += (REAL)__int_as_float((int)(unsigned)wv[q][k]);
                            }
                    } while (rem);
                    tick(2);
                    for (unsigned rem_r = m_rnd; rem_r; rem_r &= rem_r - 1) {
                        const int j = __ffs(rem_r) - 1;
                        REAL v[K];
                        row_lds<REAL, K>(v, landing + (size_t)j * LP, lane);
                        const REAL aj = __shfl_sync(0xffffffffu, al, j), sj = __shfl_sync(0xffffffffu, sel, j);
#pragma unroll
                        for (int k = 0; k < K; k++) dib[k] += aj * smooth<REAL, KERN>(v[k] - sj, p.lambda);
                    }
                } else {
                // dependencies: each managing lane waits for its strip's watermark, then fetches
                {
                    const int k0 = kind & 255;
                    if (k0 == S_DYN || k0 == S_RND) {
                        while (fv < need) {
                            __nanosleep(20);
                            fv = ld_flag(p.progress + strip);
                        }
                        if (k0 == S_RND) {
                            sel = __ldcg(p.selpos + term);
                        } else {
                            const REAL *src = p.msg + term * LP;
#pragma unroll 4
                            for (int ch = 0; ch < CH; ch++)
                                cp_async16(reinterpret_cast<char *>(my_row) + ch * 16, reinterpret_cast<const char *>(src) + ch * 16);
                        }
                    }
                }
                __syncwarp();
                tick(2);
                cp_async_wait_all();
                __syncwarp();
                for (unsigned rem_d = m_dyn | m_rnd; rem_d; rem_d &= rem_d - 1) {
                    const int j = __ffs(rem_d) - 1;
                    REAL v[K];
                    row_lds<REAL, K>(v, landing + (size_t)j * LP, lane);
                    if ((m_dyn >> j) & 1u) {
#pragma unroll
                        for (int k = 0; k < K; k++) base[k] += v[k];
                    } else {
                        const REAL aj = __shfl_sync(0xffffffffu, al, j), sj = __shfl_sync(0xffffffffu, sel, j);
#pragma unroll
                        for (int k = 0; k < K; k++) dib[k] += aj * smooth<REAL, KERN>(v[k] - sj, p.lambda);
                    }
                }
                }
                // the term warps have read the rows of node - NHW
                if (prof_on) { tclk += (long long)(base[0] != base[0]); tick(3); }
                named_sync(BAR_EMPTY + hid);
                tick(5);
                row_sts<REAL, K>(set_ptr(hid, R_BASE), base, lane);
                if (do_round) {
                    row_sts<REAL, K>(set_ptr(hid, R_DIB0), dib, lane);
                    row_sts<REAL, K>(set_ptr(hid, R_RMS), rms, lane);
                }
                named_arrive(BAR_FULL + hid);
                if (prof_on) tp[7]++;
            }
            if (prof_on && lane == 0) {
                const int grp = (fs == 0) ? 0 : 1;
                for (int q = 0; q < 8; q++) atomicAdd((unsigned long long *)p.prof + grp * 16 + 4 + q, (unsigned long long)tp[q]);
            }
            // consume the term warps' last arrival on this parity so the barrier is balanced
            named_sync(BAR_EMPTY + hid);
            continue;
        }

        if (warp == NCW + NHW) {
            // ============================================================ publisher warp
            if constexpr (MBOX) continue;   // nothing to publish: receivers validate the data itself
            int published = 0;
            while (published < n_nodes) {
                int c = (lane < NCW) ? ld_acquire_cta(s_wdone + lane) : 0x7fffffff;
#pragma unroll
                for (int o = 2; o > 0; o >>= 1) c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c > published) {
                    const long long t0 = (TDIAG && p.prof) ? clock64() : 0;
                    if (lane == 0) publish_flag(p.progress + gid, c);
                    __syncwarp();
                    if (TDIAG && p.prof && lane == 0) {
                        atomicAdd((unsigned long long *)p.prof + (fs == 0 ? 0 : 16) + 12, (unsigned long long)(clock64() - t0));
                        atomicAdd((unsigned long long *)p.prof + (fs == 0 ? 0 : 16) + 13, 1ull);
                        atomicAdd((unsigned long long *)p.prof + (fs == 0 ? 0 : 16) + 14, (unsigned long long)(c - published));
                    }
                    published = c;
                }
            }
            continue;
        }

        // ================================================================ prefetch warp
        {
            int pf_seg = sg0, pf_i = 0, pf_node = 0;     // next node to prefetch: segment, index in it, index in strip
            int pf_n = __ldg(&segs[pf_seg].n);
            for (;;) {
                int c = 0x7fffffff;
                if (lane < NCW) c = MBOX ? *reinterpret_cast<volatile int *>(s_wdone + lane) : ld_acquire_cta(s_wdone + lane);
#pragma unroll
                for (int o = 2; o > 0; o >>= 1) c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
                c = __shfl_sync(0xffffffffu, c, 0);
                const int pf_end = min(n_nodes, c + 3 + PF_DIST);
                if (pf_node >= n_nodes) break;
                if (pf_node >= pf_end) {
                    __nanosleep(200);
                    continue;
                }
                for (; pf_node < pf_end; pf_node++) {
                    if (pf_node > c + 2) {
                        // rows of node (pf_seg, pf_i): lanes 0..7 take the term warps' own terms,
                        // lanes 8..29 the helper items
                        constexpr int LR = (LP * (int)sizeof(REAL) + 127) / 128;
                        constexpr int LB = (LP + 127) / 128;
                        const Segment *g = segs + pf_seg;
                        if (lane < 2 * NCW) {
                            const SegOwn o = g->own[lane >> 1][lane & 1];
                            if (o.flags & OWN_HAS) {
                                const long long row = (o.term0 + (long long)pf_i * o.tstride) * LP;
                                const bool tail = (o.flags & OWN_TAIL) != 0;
                                for (int t = 0; t < LR; t++) {
                                    prefetch_l2(reinterpret_cast<const char *>(p.msg + row) + t * 128);
                                    prefetch_l2(reinterpret_cast<const char *>(p.posq + row) + t * 128);
                                    prefetch_l2(reinterpret_cast<const char *>(p.posqp + row) + t * 128);
                                }
                                for (int t = 0; t < LB; t++) {
                                    prefetch_l2(reinterpret_cast<const char *>((tail ? p.rank_qp : p.rank_q) + row) + t * 128);
                                    prefetch_l2(reinterpret_cast<const char *>((tail ? p.cnt_q : p.cnt_qp) + row) + t * 128);
                                }
                            }
                        } else if (lane - 2 * NCW < SCHED_ITEMS) {
                            const SegItem sl = g->item[lane - 2 * NCW];
                            const int kind = sl.kind & 255;
                            const long long row = (sl.term0 + (long long)pf_i * sl.tstride) * LP;
                            const REAL *base = nullptr;
                            if (kind == S_D) base = p.D + row;
                            else if (kind == S_RND && do_round) base = ((sl.kind & 256) ? p.posqp : p.posq) + row;
                            if (base)
                                for (int t = 0; t < LR; t++) prefetch_l2(reinterpret_cast<const char *>(base) + t * 128);
                            if constexpr (MBOX) {
                                // mailbox words the helper will poll: where the sender ran long ago (the ring in the
                                // backward pass) they have left the L2, and the poll would pay the HBM latency
                                if (kind == S_DYN && do_send) {
                                    constexpr int LM = (LP * 8 + 127) / 128;
                                    for (int t = 0; t < LM; t++) prefetch_l2(reinterpret_cast<const char *>(p.mbox + row) + t * 128);
                                } else if (kind == S_RND && do_round) {
                                    prefetch_l2(p.selbox + row / LP);
                                }
                            }
                        }
                    }
                    if (++pf_i >= pf_n) {
                        pf_i = 0;
                        pf_seg++;
                        pf_n = (pf_seg < sg1) ? __ldg(&segs[pf_seg].n) : 0x7fffffff;
                    }
                }
            }
        }
    }
    if (is_term && lane == 0) {
        if (acc_energy != 0.0) atomicAdd(p.acc + 0, acc_energy);
        if (acc_lb != 0.0) atomicAdd(p.acc + 1, acc_lb);
    }
}

// ---------------------------------------------------------------- setup kernels

// unary L x N doubles (MATLAB) -> D [N][LP] REAL, pad = 0
template <typename REAL>
__global__ void convert_unary_kernel(const double *__restrict__ src, REAL *__restrict__ dst, int L, int LP,
                                     long long N)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * LP) return;
    const long long u = i / LP;
    const int l = (int)(i % LP);
    dst[i] = (l < L) ? (REAL)src[u * L + l] : REAL(0);
}

template <typename REAL>
__global__ void convert_vec_kernel(const double *__restrict__ src, REAL *__restrict__ dst, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (REAL)src[i];
}

// One warp per term: converts q(:,p), qprim(:,p) to padded REAL rows and builds the
// rank / merge-count tables (the GPU counterpart of the argsort loop of
// trws_mex.cpp:84-119; ranks are what the sorted visit order needs, counts are
// the merge of the two sorted lists).  Non-finite positions set *bad.
template <typename REAL, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
term_tables_kernel(const double *__restrict__ q, const double *__restrict__ qp, int L, long long E,
                   REAL *__restrict__ posq, REAL *__restrict__ posqp, uint8_t *__restrict__ rank_q,
                   uint8_t *__restrict__ rank_qp, uint8_t *__restrict__ cnt_q, uint8_t *__restrict__ cnt_qp,
                   int *bad)
{
    __shared__ REAL sh[WARPS][2][32 * K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int LP = 32 * K;
    const REAL PMAX = Lim<REAL>::big() * REAL(1e-4);
    for (long long t = (long long)blockIdx.x * WARPS + warp; t < E; t += (long long)gridDim.x * WARPS) {
        REAL a[K], b[K];
        REAL amax = -PMAX, bmax = -PMAX;
        bool isbad = false, isbad2 = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int l = lane * K + k;
            if (l < L) {
                const double va = q[t * L + l], vb = qp[t * L + l];
                if (!(va == va)) isbad = true;
                if (!(vb == vb)) isbad2 = true;
                a[k] = (REAL)fmin(fmax(va, -(double)PMAX), (double)PMAX);
                b[k] = (REAL)fmin(fmax(vb, -(double)PMAX), (double)PMAX);
                amax = max(amax, a[k]);
                bmax = max(bmax, b[k]);
            }
        }
        if (isbad) atomicOr(bad, 1);    // NaN in q
        if (isbad2) atomicOr(bad, 2);   // NaN in qprim
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            amax = max(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            bmax = max(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int l = lane * K + k;
            if (l >= L) { a[k] = amax; b[k] = bmax; }
            sh[warp][0][l] = a[k];
            sh[warp][1][l] = b[k];
        }
        __syncwarp();
        int ra[K], rb[K], ca[K], cb[K];
#pragma unroll
        for (int k = 0; k < K; k++) ra[k] = rb[k] = ca[k] = cb[k] = 0;
        for (int m = 0; m < L; m++) {
            const REAL va = sh[warp][0][m], vb = sh[warp][1][m];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int l = lane * K + k;
                ra[k] += (va < a[k]) || (va == a[k] && m < l);
                rb[k] += (vb < b[k]) || (vb == b[k] && m < l);
                ca[k] += (vb <= a[k]); // #{qprim <= q[l]}
                cb[k] += (va <= b[k]); // #{q <= qprim[l]}
            }
        }
        uint8_t o0[K], o1[K], o2[K], o3[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int l = lane * K + k;
            o0[k] = (uint8_t)(l < L ? ra[k] : l);
            o1[k] = (uint8_t)(l < L ? rb[k] : l);
            o2[k] = (uint8_t)min(ca[k], 255);
            o3[k] = (uint8_t)min(cb[k], 255);
        }
        const long long off = t * LP + lane * K;
        VecIO<REAL, K>::store(posq + off, a);
        VecIO<REAL, K>::store(posqp + off, b);
        ByteIO<K>::store(rank_q + off, o0);
        ByteIO<K>::store(rank_qp + off, o1);
        ByteIO<K>::store(cnt_q + off, o2);
        ByteIO<K>::store(cnt_qp + off, o3);
        __syncwarp();
    }
}

} // namespace trws
} // namespace sb
