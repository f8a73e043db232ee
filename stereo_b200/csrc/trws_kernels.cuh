// trws_kernels.cuh -- sm_100a device code of the TRW-S message sweep (K5) and
// its setup kernels.  Replaces, on the GPU:
//   Minimize_TRW_S forward / backward sweeps      cpp/trw-s/minimize.cpp:31-95
//   ComputeSolutionAndEnergy (primal rounding)    cpp/trw-s/minimize.cpp:223-264
//   TypeStereoLinear::Edge::UpdateMessage         cpp/trw-s/typeStereoLinear.h:329-487
//   TypeStereoQuadratic::Edge::UpdateMessage      cpp/trw-s/typeStereoQuadratic.h:329-501
//   Edge::AddColumn / Smooth                      typeStereoLinear.h:324-327,491-518
//   the per-edge argsort of trws_mex.cpp:84-119   (rank / merge-count tables)
//
// Execution model (DESIGN.md "K5"):
//   * one warp owns one node at a time; the L labels of every per-node vector
//     are blocked over the lanes (label = lane*K + k, K = LP/32 in registers);
//   * work is dispatched in STRIPS (trws_order.cpp: the boundary ring, then one
//     strip per interior row) through an atomic ticket; a warp walks its strip
//     node by node, keeps the two messages it just sent to the next node of the
//     strip in registers (no round trip for the in-strip dependency) and waits
//     on epoch flags (release/acquire through L2) only for neighbours owned by
//     other strips -- normally satisfied long before.  A whole sweep is ONE
//     persistent launch with no grid barriers, and because it honours the
//     reference's orientation DAG its results equal the sequential sweep's;
//   * while a node is processed the operands of the next one are prefetched
//     into L2 (prefetch.global.L2);
//   * the min-plus update is O(L): labels are visited in the order of their
//     (irregular) positions through iteration-invariant uint8 rank tables, the
//     two directional distance transforms are warp-shuffle scans over
//     (offset, value) pairs -- no h - alpha*x cancellation -- and each
//     destination label looks its two bracketing sources up through a
//     precomputed merge count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {
namespace trws {

enum { PASS_FWD = 0, PASS_BWD = 1 };
enum { MODE_SEND = 1, MODE_ROUND = 2 };

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
    const uint8_t *info;    // [N] incidence byte: valid mask | lower mask << 4 (trws_order.cpp)
    const int32_t *nodes;   // [N] forward-sweep strips, concatenated (backward = reverse)
    const int32_t *strip_ptr; // [S+1]
    int S;
    int active_warps;       // warps per CTA that take strips (spreads few strips over all SMs)
    int32_t *done;          // [N] epoch flags
    int32_t *sol;           // [N] rounded labels (0-based)
    int *ticket;            // dispatch counter (zeroed before each launch)
    double *acc;            // [0] energy  [1] lower bound (zeroed before each launch)
    int epoch;
    int mode;               // PASS_FWD only: MODE_SEND | MODE_ROUND
};

// ---------------------------------------------------------------- helpers

__device__ __forceinline__ int ld_acquire(const int32_t *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int32_t *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T> __device__ __forceinline__ T warp_min(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
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

// Shared scratch of one warp: sorted-domain (value, position) pairs.
// Logical sorted index i in [0, LP) lives at phys(i); slot 0 and slot
// phys(LP) are the (+big, 0) sentinels for "no source on this side".
// Even K gets one pad slot per K so the blocked per-lane reads are
// bank-conflict free.
template <int K> __device__ __forceinline__ int phys(int i)
{
    if constexpr (K % 2 == 0) return 1 + i + i / K;
    else return 1 + i;
}
template <int K> __host__ __device__ constexpr int scratch_pairs() { return 2 + 32 * K + ((K % 2 == 0) ? 32 : 0); }

// ---------------------------------------------------------------- min-plus updates
//
// Both return vMin and leave the min-normalised new message in m[] (label
// domain).  Di is the (gamma-unscaled) node sum, m the old message on this term,
// s/rk the sender's positions and their ranks, x/cn the receiver's positions
// and merge counts.

// Truncated linear: msg[j] = min(vTrunc, min_i H_i + alpha |x_j - s_i|)
// (typeStereoLinear.h:375-480; equality with the reference's cone envelope is
// SURVEY.md 3.3 [probe] and tests/test_trws_message.py).
template <typename REAL, int K>
__device__ __forceinline__ REAL update_linear(REAL gamma, REAL alpha, REAL lambda, int L, int lane,
                                              const REAL (&Di)[K], REAL (&m)[K], const REAL (&s)[K],
                                              const uint8_t (&rk)[K], const REAL (&x)[K],
                                              const uint8_t (&cn)[K], Pair<REAL> *P)
{
    const REAL BIG = Lim<REAL>::big();
    REAL h[K];
    REAL hmin = BIG;
#pragma unroll
    for (int k = 0; k < K; k++) {
        h[k] = (lane * K + k < L) ? gamma * Di[k] - m[k] : BIG;
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
    // left-to-right transform  G_i = min(G_{i-1} + alpha (s_i - s_{i-1}), h_i)
    REAL gl[K], cl[K];
    {
        REAL sprev = __shfl_up_sync(0xffffffffu, ss[K - 1], 1);
        if (lane == 0) sprev = ss[0];
        REAL g = BIG, cum = REAL(0);
#pragma unroll
        for (int k = 0; k < K; k++) {
            const REAL d = alpha * (ss[k] - (k ? ss[k - 1] : sprev));
            g = min(g + d, hs[k]);
            cum += d;
            gl[k] = g;
            cl[k] = cum;
        }
        REAL Dl = cum, C = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const REAL Dp = __shfl_up_sync(0xffffffffu, Dl, o);
            const REAL Cp = __shfl_up_sync(0xffffffffu, C, o);
            if (lane >= o) {
                C = min(Cp + Dl, C);
                Dl = Dp + Dl;
            }
        }
        REAL Vin = __shfl_up_sync(0xffffffffu, C, 1);
        if (lane == 0) Vin = BIG;
#pragma unroll
        for (int k = 0; k < K; k++) gl[k] = min(Vin + cl[k], gl[k]);
    }
    // right-to-left transform, merged into F_i = DT(s_i)
    {
        REAL snext = __shfl_down_sync(0xffffffffu, ss[0], 1);
        if (lane == 31) snext = ss[K - 1];
        REAL g = BIG, cum = REAL(0);
        REAL gr[K];
#pragma unroll
        for (int k = K - 1; k >= 0; k--) {
            const REAL d = alpha * ((k < K - 1 ? ss[k + 1] : snext) - ss[k]);
            g = min(g + d, hs[k]);
            cum += d;
            gr[k] = g;
            cl[k] = cum;
        }
        REAL Dl = cum, C = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const REAL Dp = __shfl_down_sync(0xffffffffu, Dl, o);
            const REAL Cp = __shfl_down_sync(0xffffffffu, C, o);
            if (lane + o < 32) {
                C = min(Cp + Dl, C);
                Dl = Dp + Dl;
            }
        }
        REAL Vin = __shfl_down_sync(0xffffffffu, C, 1);
        if (lane == 31) Vin = BIG;
#pragma unroll
        for (int k = 0; k < K; k++) {
            Pair<REAL> t;
            t.a = min(gl[k], min(Vin + cl[k], gr[k]));
            t.b = ss[k];
            P[phys<K>(lane * K + k)] = t;
        }
    }
    __syncwarp();
    REAL vmin = BIG;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int c = cn[k];
        const Pair<REAL> lo = P[c ? phys<K>(c - 1) : 0];
        const Pair<REAL> hi = P[phys<K>(c)];
        const REAL v = min(vTrunc, min(lo.a + alpha * fabs(x[k] - lo.b), hi.a + alpha * fabs(x[k] - hi.b)));
        m[k] = v;
        if (lane * K + k < L) vmin = min(vmin, v);
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
__device__ __forceinline__ REAL update_quadratic(REAL gamma, REAL alpha, REAL lambda, int L, int lane,
                                                 const REAL (&Di)[K], REAL (&m)[K], const REAL (&s)[K],
                                                 const uint8_t (&rk)[K], const REAL (&x)[K],
                                                 const uint8_t (&cn)[K], Pair<REAL> *P)
{
    const REAL BIG = Lim<REAL>::big();
    REAL hmin = BIG;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const REAL h = (lane * K + k < L) ? gamma * Di[k] - m[k] : BIG;
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
        if (lane * K + k < L) {
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

// ---------------------------------------------------------------- incidence on the grid
// Direction d: 0 up (r-1), 1 down (r+1), 2 left (c-1), 3 right (c+1).
// j = 0/1 selects the two terms of the neighbour pair.  Term order follows
// dispmap_super.m:284-294.
struct Incidence {
    long long nb;     // neighbour node
    long long term[2];
    bool tail[2];     // am I the tail (conn(0,p)) of term[j]?
    bool valid;
};

__device__ __forceinline__ Incidence incidence(int d, int r, int c, long long u, int H, int W,
                                               long long nV, long long nH)
{
    Incidence I;
    if (d == 0) {
        I.valid = r > 0;
        I.nb = u - 1;
        I.term[0] = (long long)c * (H - 1) + (r - 1); // VD(r-1,c): tail = nb
        I.term[1] = I.term[0] + nV;                   // VU(r-1,c): tail = me
        I.tail[0] = false; I.tail[1] = true;
    } else if (d == 1) {
        I.valid = r < H - 1;
        I.nb = u + 1;
        I.term[0] = (long long)c * (H - 1) + r;       // VD(r,c): tail = me
        I.term[1] = I.term[0] + nV;                   // VU(r,c): tail = nb
        I.tail[0] = true; I.tail[1] = false;
    } else if (d == 2) {
        I.valid = c > 0;
        I.nb = u - H;
        I.term[0] = 2 * nV + (long long)(c - 1) * H + r; // HR(r,c-1): tail = nb
        I.term[1] = I.term[0] + nH;                      // HL(r,c-1): tail = me
        I.tail[0] = false; I.tail[1] = true;
    } else {
        I.valid = c < W - 1;
        I.nb = u + H;
        I.term[0] = 2 * nV + (long long)c * H + r;       // HR(r,c): tail = me
        I.term[1] = I.term[0] + nH;                      // HL(r,c): tail = nb
        I.tail[0] = true; I.tail[1] = false;
    }
    return I;
}

// ---------------------------------------------------------------- the sweep

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// direction (0 up, 1 down, 2 left, 3 right) from node u to node v, or -1
__device__ __forceinline__ int direction_to(long long u, long long v, int H)
{
    const long long d = v - u;
    return d == -1 ? 0 : d == 1 ? 1 : d == -H ? 2 : d == H ? 3 : -1;
}

// Pull the operands node `u` will need (its unary row and, for every term it sends on,
// the old message, both position rows, the rank row and the merge-count row) into L2.
template <typename REAL, int K, int PASS>
__device__ __forceinline__ void prefetch_node(const Problem<REAL> &p, long long u, int lane)
{
    constexpr int LP = 32 * K;
    constexpr int LR = (LP * (int)sizeof(REAL) + 127) / 128; // 128B lines per REAL row
    constexpr int LB = (LP + 127) / 128;                     // lines per byte row
    const unsigned info = __ldg(p.info + u);
    const unsigned valid = info & 15u, lower = info >> 4;
    const unsigned send_mask = (PASS == PASS_FWD) ? (valid & ~lower) : lower;
    const int r = (int)(u % p.H), c = (int)(u / p.H);
    if (lane < LR) prefetch_l2(reinterpret_cast<const char *>(p.D + u * LP) + lane * 128);
#pragma unroll 1
    for (int d = 0; d < 4; d++) {
        if (!((send_mask >> d) & 1u)) continue;
        const Incidence I = incidence(d, r, c, u, p.H, p.W, p.nV, p.nH);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const long long row = I.term[j] * LP;
            for (int i = lane; i < 3 * LR + 2 * LB; i += 32) {
                const char *base;
                int line;
                if (i < 3 * LR) {
                    const int a = i / LR;
                    line = i % LR;
                    base = reinterpret_cast<const char *>((a == 0 ? (const REAL *)p.msg : a == 1 ? p.posq : p.posqp) + row);
                } else {
                    const int a = (i - 3 * LR) / LB;
                    line = (i - 3 * LR) % LB;
                    base = reinterpret_cast<const char *>(
                        (a == 0 ? (I.tail[j] ? p.rank_qp : p.rank_q) : (I.tail[j] ? p.cnt_q : p.cnt_qp)) + row);
                }
                prefetch_l2(base + line * 128);
            }
        }
    }
}

template <typename REAL, int K, int KERN, int PASS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sweep_kernel(const Problem<REAL> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    Pair<REAL> *P = reinterpret_cast<Pair<REAL> *>(smem_raw) + (size_t)warp * scratch_pairs<K>();
    const REAL BIG = Lim<REAL>::big();
    constexpr int LP = 32 * K;
    if (lane == 0) {
        Pair<REAL> t;
        t.a = BIG;
        t.b = REAL(0);
        P[0] = t;
        P[phys<K>(LP)] = t;
    }
    __syncwarp();

    const bool do_send = (PASS == PASS_BWD) || (p.mode & MODE_SEND);
    const bool do_round = (PASS == PASS_FWD) && (p.mode & MODE_ROUND);
    double acc_energy = 0.0, acc_lb = 0.0;
    if (warp >= p.active_warps) return;

    for (;;) {
        int ts = 0;
        if (lane == 0) ts = atomicAdd(p.ticket, 1);
        ts = __shfl_sync(0xffffffffu, ts, 0);
        if (ts >= p.S) break;
        // the backward sweep runs the forward schedule in reverse
        const int fs = (PASS == PASS_BWD) ? p.S - 1 - ts : ts;
        const int sb = __ldg(p.strip_ptr + fs), se = __ldg(p.strip_ptr + fs + 1);
        const int step = (PASS == PASS_BWD) ? -1 : 1;
        int idx = (PASS == PASS_BWD) ? se - 1 : sb;
        long long u = (se > sb) ? __ldg(p.nodes + idx) : -1;
        long long u_prev = -1;
        bool carry_valid = false;
        REAL carry0[K], carry1[K]; // messages this warp just sent to u on the pair's terms 0 / 1
#pragma unroll
        for (int k = 0; k < K; k++) carry0[k] = carry1[k] = REAL(0);

        for (int n = se - sb; n > 0; n--, idx += step) {
            const long long u_next = (n > 1) ? (long long)__ldg(p.nodes + idx + step) : -1;
            if (u_next >= 0) prefetch_node<REAL, K, PASS>(p, u_next, lane);
            const int r = (int)(u % p.H), c = (int)(u / p.H);
            const unsigned info = __ldg(p.info + u);
            const unsigned valid_mask = info & 15u, lower_mask = info >> 4;
            // gamma = 1/max(nF, nB), two terms per neighbour (treeProbabilities.cpp:28-45)
            const int nB = 2 * __popc(lower_mask), nF = 2 * __popc(valid_mask & ~lower_mask);
            const REAL gamma = REAL(1) / REAL(max(1, max(nF, nB)));
            const unsigned dep_mask = (PASS == PASS_FWD) ? lower_mask : (valid_mask & ~lower_mask);
            const unsigned send_mask = (PASS == PASS_FWD) ? (valid_mask & ~lower_mask) : lower_mask;
            const int d_prev = carry_valid ? direction_to(u, u_prev, p.H) : -1;
            const int d_next = (u_next >= 0) ? direction_to(u, u_next, p.H) : -1;

            // wait for the neighbours this node depends on that other strips own
            if (lane < 4 && ((dep_mask >> lane) & 1u) && lane != d_prev) {
                const Incidence I = incidence(lane, r, c, u, p.H, p.W, p.nV, p.nH);
                while (ld_acquire(p.done + I.nb) < p.epoch) __nanosleep(20);
            }
            __syncwarp();

            REAL Di[K];
            VecIO<REAL, K>::load_ro(Di, p.D + u * LP + lane * K);

            if (do_round) {
                // minimize.cpp:240-260: DiB = D + sum_{lower nb} V(x_nb, .), Dr = DiB + forward messages
                REAL DiB[K], Dr[K];
#pragma unroll
                for (int k = 0; k < K; k++) DiB[k] = Di[k];
#pragma unroll 1
                for (int e = 0; e < 8; e++) {
                    const int d = e >> 1, j = e & 1;
                    if (!((lower_mask >> d) & 1u)) continue;
                    const Incidence I = incidence(d, r, c, u, p.H, p.W, p.nV, p.nH);
                    const long long tm = I.term[j];
                    const REAL *mine = (I.tail[j] ? p.posqp : p.posq) + tm * LP;
                    const REAL *theirs = (I.tail[j] ? p.posq : p.posqp) + tm * LP;
                    const int xs = __ldcg(p.sol + I.nb);
                    const REAL pos_nb = theirs[xs];
                    const REAL al = p.alpha[tm];
                    REAL mp[K];
                    VecIO<REAL, K>::load_ro(mp, mine + lane * K);
#pragma unroll
                    for (int k = 0; k < K; k++) DiB[k] += al * smooth<REAL, KERN>(mp[k] - pos_nb, p.lambda);
                }
#pragma unroll
                for (int k = 0; k < K; k++) Dr[k] = DiB[k];
#pragma unroll 1
                for (int e = 0; e < 8; e++) {
                    const int d = e >> 1, j = e & 1;
                    if (!(((valid_mask & ~lower_mask) >> d) & 1u)) continue;
                    const Incidence I = incidence(d, r, c, u, p.H, p.W, p.nV, p.nH);
                    REAL mm[K];
                    VecIO<REAL, K>::load_cg(mm, p.msg + I.term[j] * LP + lane * K);
#pragma unroll
                    for (int k = 0; k < K; k++) Dr[k] += mm[k];
                }
                // Vector::ComputeMin: first minimum in label order (typeStereoLinear.h:238-252)
                REAL best = BIG;
                int bi = 0x7fffffff;
                REAL bDiB = REAL(0);
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int lbl = lane * K + k;
                    if (lbl < p.L && Dr[k] < best) {
                        best = Dr[k];
                        bi = lbl;
                        bDiB = DiB[k];
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const REAL ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    const REAL od = __shfl_xor_sync(0xffffffffu, bDiB, o);
                    if (ob < best || (ob == best && oi < bi)) {
                        best = ob;
                        bi = oi;
                        bDiB = od;
                    }
                }
                if (lane == 0) {
                    p.sol[u] = bi;
                    acc_energy += (double)bDiB;
                }
            }

            if (do_send) {
                // Di = D + all incident messages (minimize.cpp:38-46 / 69-77); the pair just
                // sent by this warp from the previous node of the strip comes from registers
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int d = e >> 1, j = e & 1;
                    if (!((valid_mask >> d) & 1u)) continue;
                    if (d == d_prev) {
#pragma unroll
                        for (int k = 0; k < K; k++) Di[k] += (j ? carry1[k] : carry0[k]);
                    } else {
                        const Incidence I = incidence(d, r, c, u, p.H, p.W, p.nV, p.nH);
                        REAL mm[K];
                        VecIO<REAL, K>::load_cg(mm, p.msg + I.term[j] * LP + lane * K);
#pragma unroll
                        for (int k = 0; k < K; k++) Di[k] += mm[k];
                    }
                }
                if (PASS == PASS_BWD) {
                    // ComputeAndSubtractMin + lower bound (minimize.cpp:79-81)
                    REAL vmin = BIG;
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (lane * K + k < p.L) vmin = min(vmin, Di[k]);
                    vmin = warp_min(vmin);
#pragma unroll
                    for (int k = 0; k < K; k++) Di[k] -= vmin;
                    acc_lb += (double)vmin;
                }
                carry_valid = false;
#pragma unroll 1
                for (int e = 0; e < 8; e++) {
                    const int d = e >> 1, j = e & 1;
                    if (!((send_mask >> d) & 1u)) continue;
                    const Incidence I = incidence(d, r, c, u, p.H, p.W, p.nV, p.nH);
                    const long long tm = I.term[j];
                    const long long off = tm * LP + lane * K;
                    // sender's positions: qprim if I am the tail of the term, else q
                    // (typeStereoLinear.h:343-357 with Swap(), MRFEnergy.cpp:200-203)
                    REAL m[K], s[K], x[K];
                    uint8_t rk[K], cn[K];
                    VecIO<REAL, K>::load_cg(m, p.msg + off);
                    VecIO<REAL, K>::load_ro(s, (I.tail[j] ? p.posqp : p.posq) + off);
                    VecIO<REAL, K>::load_ro(x, (I.tail[j] ? p.posq : p.posqp) + off);
                    ByteIO<K>::load(rk, (I.tail[j] ? p.rank_qp : p.rank_q) + off);
                    ByteIO<K>::load(cn, (I.tail[j] ? p.cnt_q : p.cnt_qp) + off);
                    const REAL al = p.alpha[tm];
                    REAL vmin;
                    if constexpr (KERN == 1)
                        vmin = update_linear<REAL, K>(gamma, al, p.lambda, p.L, lane, Di, m, s, rk, x, cn, P);
                    else
                        vmin = update_quadratic<REAL, K>(gamma, al, p.lambda, p.L, lane, Di, m, s, rk, x, cn, P);
                    VecIO<REAL, K>::store(p.msg + off, m);
                    if (PASS == PASS_BWD) acc_lb += (double)vmin;
                    if (d == d_next) {
                        carry_valid = true;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            if (j) carry1[k] = m[k];
                            else carry0[k] = m[k];
                        }
                    }
                }
            } else {
                carry_valid = false;
            }

            // publish: this warp's stores happen-before the flag (syncwarp + release)
            __syncwarp();
            if (lane == 0) st_release(p.done + u, p.epoch);
            u_prev = u;
            u = u_next;
        }
    }
    if (lane == 0) {
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
        bool isbad = false;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int l = lane * K + k;
            if (l < L) {
                const double va = q[t * L + l], vb = qp[t * L + l];
                if (!(va == va) || !(vb == vb)) isbad = true;
                a[k] = (REAL)fmin(fmax(va, -(double)PMAX), (double)PMAX);
                b[k] = (REAL)fmin(fmax(vb, -(double)PMAX), (double)PMAX);
                amax = max(amax, a[k]);
                bmax = max(bmax, b[k]);
            }
        }
        if (isbad) atomicExch(bad, 1);
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
