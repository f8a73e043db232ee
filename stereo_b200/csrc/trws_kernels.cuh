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
#include "trws_sched.h"

// Build-time switches for timing experiments (make EXTRA=-DSB_TRWS_INSTRUMENT=1): per-phase
// cycle counters (SB_TRWS_PROFILE) and work-skipping bisection flags (SB_TRWS_DEBUG).
#ifndef SB_TRWS_INSTRUMENT
#define SB_TRWS_INSTRUMENT 0
#endif

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
    const SegWarp *segs;    // segment descriptors of THIS pass (trws_sched.h): [nseg][NCW]
    const int32_t *seg_ptr; // [S+1] segments of each forward strip
    const int32_t *strip_ptr; // [S+1] node count prefix of the forward strips
    int S;
    int32_t *progress;      // [S] nodes of each strip completed in this pass (zeroed per launch)
    int32_t *sol;           // [N] rounded labels (0-based)
    REAL *selpos;           // [E] position of the sender's rounded label on each forward term
    long long *prof;        // optional [2][8] cycle counters (SB_TRWS_PROFILE), else null
    int *ticket;            // dispatch counter (zeroed before each launch)
    double *acc;            // [0] energy  [1] lower bound (zeroed before each launch)
    int mode;               // PASS_FWD only: MODE_SEND | MODE_ROUND
    int debug;              // SB_TRWS_DEBUG bisection switches (timing experiments only; 0 in production)
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
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}
// step barrier of the NCW compute warps (the auxiliary warp is not paced by it)
__device__ __forceinline__ void step_barrier()
{
    asm volatile("bar.sync 1, %0;" ::"n"(128) : "memory");
}
__device__ __forceinline__ void st_release_cta(int *smem, int v)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int *smem)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
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
// CTA = NCW compute warps + 1 auxiliary warp; a CTA walks one strip at a time,
// driven by the segment descriptors of trws_sched.h (no grid arithmetic here).
// A node is processed in one STEP (two when it sends on more than NCW terms):
//
//   phase A (before the step's barrier), per compute warp: sum what this warp
//     fetched for the node -- the old message of its own send term (registers,
//     loaded one step ahead), rows that arrived in its landing buffers through
//     cp.async (the unary row, dependency messages published by other strips,
//     position rows for the rounding) -- into partial rows in shared memory;
//   barrier (compute warps only);
//   phase B: every compute warp issues the loads of the NEXT step, forms Di
//     (minimize.cpp:38-46 / 69-77) from the partial rows plus the two messages
//     the CTA sent to this node in the previous step (carry rows in shared
//     memory), rounds the node when the pass carries the primal sweep
//     (minimize.cpp:240-260), runs the min-plus update of its own term and
//     stores it.
//   The auxiliary warp watches the compute warps' completion counters,
//   publishes the strip's progress watermark (gpu-scope fence + store, far too
//   slow for the dependent chain) and pulls the operands of the nodes ahead
//   into L2.
//
// Shared rows are double buffered by node parity, so one barrier per step is
// enough.

constexpr int NCW = SCHED_NCW;
constexpr int CTA_THREADS = (NCW + 1) * 32;
constexpr int SLOTS = SCHED_SLOTS; // landing rows per compute warp
constexpr int ROWS_PER_PAR = 17;   // RM[4] RX[4] RB[4] DR CM[2] CC[2]
constexpr int PF_DIST = 6;

enum { R_RM = 0, R_RX = 4, R_RB = 8, R_DR = 12, R_CM = 13, R_CC = 15 };

template <typename REAL, int K> __host__ __device__ constexpr size_t sweep_smem_bytes()
{
    return (size_t)(2 * ROWS_PER_PAR + NCW * SLOTS) * 32 * K * sizeof(REAL) +
           (size_t)NCW * scratch_pairs<K>() * sizeof(Pair<REAL>);
}

template <typename REAL, int K> struct OwnTerm {
    REAL m[K], s[K], x[K];
    uint8_t rk[K], cn[K];
    REAL alpha;
    long long term;
    int flags;       // OWN_*
};

template <typename REAL, int K, int KERN, int PASS>
__global__ void __launch_bounds__(CTA_THREADS) sweep_kernel(const Problem<REAL> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    __shared__ int s_wdone[NCW]; // nodes of the current strip each compute warp has completed
    __shared__ __align__(16) SegWarp s_desc[NCW][2];
    constexpr int LP = 32 * K;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const bool is_aux = (warp == NCW);
    REAL *rows = reinterpret_cast<REAL *>(smem_raw);
    REAL *landing = rows + (size_t)2 * ROWS_PER_PAR * LP + (size_t)(is_aux ? 0 : warp) * SLOTS * LP;
    Pair<REAL> *P = reinterpret_cast<Pair<REAL> *>(rows + (size_t)(2 * ROWS_PER_PAR + NCW * SLOTS) * LP) +
                    (size_t)(is_aux ? 0 : warp) * scratch_pairs<K>();
    const REAL BIG = Lim<REAL>::big();
    if (!is_aux && lane == 0) {
        Pair<REAL> t;
        t.a = BIG;
        t.b = REAL(0);
        P[0] = t;
        P[phys<K>(LP)] = t;
    }
    const bool do_send = (PASS == PASS_BWD) || (p.mode & MODE_SEND);
    const bool do_round = (PASS == PASS_FWD) && (p.mode & MODE_ROUND);
    const SegWarp *const segs = p.segs;
    double acc_energy = 0.0, acc_lb = 0.0;

    auto row_ptr = [&](int par, int r) -> REAL * { return rows + (size_t)(par * ROWS_PER_PAR + r) * LP; };
    auto slot_active = [&](int kind) -> bool {
        const int k = kind & 255;
        return k == S_D || k == S_STAT || (k == S_DYN && do_send) || (k == S_RND && do_round);
    };

    for (;;) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1);
        __syncthreads(); // the auxiliary warp has published the previous strip completely
        const int ts = s_ticket;
        if (threadIdx.x < NCW) s_wdone[threadIdx.x] = 0;
        __syncthreads();
        if (ts >= p.S) break;
        // the backward sweep runs the forward schedule in reverse
        const int fs = (PASS == PASS_BWD) ? p.S - 1 - ts : ts;
        const int sg0 = __ldg(p.seg_ptr + fs), sg1 = __ldg(p.seg_ptr + fs + 1);
        const int n_nodes = __ldg(p.strip_ptr + fs + 1) - __ldg(p.strip_ptr + fs);
        if (sg1 <= sg0) continue;

        if (is_aux) {
            // ------------------------------------------------------------ auxiliary warp
            int published = 0;
            int pf_seg = sg0, pf_i = 0, pf_node = 0;     // next node to prefetch: segment, index in it, index in strip
            int pf_n = __ldg(&segs[(size_t)pf_seg * NCW].n);
            while (published < n_nodes) {
                int c = (lane < NCW) ? ld_acquire_cta(s_wdone + lane) : 0x7fffffff;
#pragma unroll
                for (int o = 2; o > 0; o >>= 1) c = min(c, __shfl_xor_sync(0xffffffffu, c, o));
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c > published) {
                    if (lane == 0) publish_flag(p.progress + fs, c);
                    published = c;
                }
                const int pf_end = min(n_nodes, c + 1 + PF_DIST);
                if (pf_node >= pf_end) {
                    if (published < n_nodes) __nanosleep(40);
                    continue;
                }
                for (; pf_node < pf_end; pf_node++) {
                    if (pf_node > c) {
                        // rows of node (pf_seg, pf_i): lanes 0..3 take the four warps' own terms,
                        // lanes 4..19 the sixteen slots
                        constexpr int LR = (LP * (int)sizeof(REAL) + 127) / 128;
                        constexpr int LB = (LP + 127) / 128;
                        const SegWarp *g = segs + (size_t)pf_seg * NCW;
                        if (lane < NCW) {
                            for (int h = 0; h < 2; h++) {
                                const SegOwn o = g[lane].own[h];
                                if (!(o.flags & OWN_HAS)) continue;
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
                        } else if (lane < NCW + NCW * SLOTS) {
                            const SegSlot sl = g[(lane - NCW) >> 2].slot[(lane - NCW) & 3];
                            const int kind = sl.kind & 255;
                            const long long row = (sl.term0 + (long long)pf_i * sl.tstride) * LP;
                            const REAL *base = nullptr;
                            if (kind == S_D) base = p.D + row;
                            else if (kind == S_STAT) base = p.msg + row;
                            else if (kind == S_RND && do_round) base = ((sl.kind & 256) ? p.posqp : p.posq) + row;
                            if (base)
                                for (int t = 0; t < LR; t++) prefetch_l2(reinterpret_cast<const char *>(base) + t * 128);
                        }
                    }
                    if (++pf_i >= pf_n) {
                        pf_i = 0;
                        pf_seg++;
                        pf_n = (pf_seg < sg1) ? __ldg(&segs[(size_t)pf_seg * NCW].n) : 0x7fffffff;
                    }
                }
            }
            continue;
        }

        // ---------------------------------------------------------------- compute warps
        const int w = warp;
        // optional phase timers (warp 0 only): 0 flag spin, 1 rest of phase A, 2 barrier, 3 prepare,
        // 4 B1 (Di / rounding), 5 resolve, 6 update + stores, 7 steps
        long long tprof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const bool prof_on = SB_TRWS_INSTRUMENT && (p.prof != nullptr) && w == 0;
        long long tclk = prof_on ? clock64() : 0;
        auto tick = [&](int which) {
            if (prof_on) {
                const long long now = clock64();
                tprof[which] += now - tclk;
                tclk = now;
            }
        };

        // descriptor of segment sg -> s_desc[w][buf] (12 lanes x 16 bytes, own warp only)
        auto fetch_desc = [&](int sg, int buf) {
            if (lane < (int)sizeof(SegWarp) / 16)
                cp_async16(reinterpret_cast<char *>(&s_desc[w][buf]) + lane * 16,
                           reinterpret_cast<const char *>(segs + (size_t)sg * NCW + w) + lane * 16);
        };

        OwnTerm<REAL, K> own;       // operands of this warp's term in the current step
        own.flags = 0;
        // per-slot bookkeeping of the node whose rows are in flight / in the landing buffers
        int slot_kind[SLOTS], slot_flag[SLOTS], slot_need[SLOTS], slot_fv[SLOTS];
        REAL slot_alpha[SLOTS], slot_sel[SLOTS];
        long long slot_term[SLOTS];
#pragma unroll
        for (int q = 0; q < SLOTS; q++) { slot_kind[q] = S_NONE; slot_flag[q] = -1; slot_need[q] = 0; slot_fv[q] = 0; slot_alpha[q] = REAL(0); slot_sel[q] = REAL(0); slot_term[q] = 0; }
        REAL Di[K];
#pragma unroll
        for (int k = 0; k < K; k++) Di[k] = REAL(0);
        int xs = 0;   // rounded label of the current node

        // prepare(): first part of issuing the loads of a step (node i of the segment
        // described by g, half `half`): own-term operands into registers, static rows into the
        // landing buffers, and -- without branching on them yet -- the progress counters that
        // guard the dependency rows.
        auto prepare = [&](const SegWarp *g, int i, int half, OwnTerm<REAL, K> &o) {
            const SegOwn so = g->own[half];
            o.flags = so.flags;
            if (so.flags & OWN_HAS) {
                o.term = so.term0 + (long long)i * so.tstride;
                const long long off = o.term * LP + lane * K;
                const bool tail = (so.flags & OWN_TAIL) != 0;
                if (SB_TRWS_INSTRUMENT && (p.debug & 4)) {
#pragma unroll
                    for (int k = 0; k < K; k++) { o.m[k] = REAL(0); o.s[k] = REAL(k); o.x[k] = REAL(k); o.rk[k] = (uint8_t)(lane * K + k); o.cn[k] = (uint8_t)(lane * K + k); }
                    o.alpha = REAL(1);
                } else {
                    // sender's positions: qprim if I am the tail of the term, else q
                    // (typeStereoLinear.h:343-357 with Swap(), MRFEnergy.cpp:200-203)
                    VecIO<REAL, K>::load_cg(o.m, p.msg + off);
                    VecIO<REAL, K>::load_ro(o.s, (tail ? p.posqp : p.posq) + off);
                    VecIO<REAL, K>::load_ro(o.x, (tail ? p.posq : p.posqp) + off);
                    ByteIO<K>::load(o.rk, (tail ? p.rank_qp : p.rank_q) + off);
                    ByteIO<K>::load(o.cn, (tail ? p.cnt_q : p.cnt_qp) + off);
                    o.alpha = __ldg(p.alpha + o.term);
                }
            }
            if (half == 0) {
#pragma unroll
                for (int q = 0; q < SLOTS; q++) {
                    const SegSlot sl = g->slot[q];
                    const int kind = slot_active(sl.kind) ? sl.kind : (int)S_NONE;
                    slot_kind[q] = kind;
                    slot_flag[q] = -1;
                    if (kind == S_NONE) continue;
                    const long long term = sl.term0 + (long long)i * sl.tstride;
                    slot_term[q] = term;
                    REAL *dst = landing + (size_t)q * LP;
                    const int k = kind & 255;
                    if (k == S_D) {
                        row_async<REAL, K>(dst, p.D + term * LP, lane);
                    } else if (k == S_STAT) {
                        row_async<REAL, K>(dst, p.msg + term * LP, lane);
                    } else {
                        slot_flag[q] = sl.strip;
                        slot_need[q] = sl.need0 + i * sl.dneed;
                        slot_fv[q] = ld_flag(p.progress + sl.strip);
                        if (k == S_RND) { // my positions on the term now, the neighbour's selected position later
                            row_async<REAL, K>(dst, ((kind & 256) ? p.posqp : p.posq) + term * LP, lane);
                            slot_alpha[q] = __ldg(p.alpha + term);
                        }
                    }
                }
            }
        };
        // resolve(): second part -- dependency rows whose counter was already high enough are
        // fetched now; the others are left to phase A of their step.
        auto resolve = [&]() {
#pragma unroll
            for (int q = 0; q < SLOTS; q++) {
                if (slot_flag[q] >= 0 && slot_fv[q] >= slot_need[q]) {
                    if ((slot_kind[q] & 255) == S_DYN) row_async<REAL, K>(landing + (size_t)q * LP, p.msg + slot_term[q] * LP, lane);
                    else slot_sel[q] = __ldcg(p.selpos + slot_term[q]);
                    slot_flag[q] = -1;
                }
            }
        };

        // first descriptor (blocking), the one after it in the background
        int sg = sg0, buf = 0;
        fetch_desc(sg, 0);
        cp_async_wait_all();
        __syncwarp();
        if (sg + 1 < sg1) fetch_desc(sg + 1, 1);
        const SegWarp *g = &s_desc[w][0];
        int seg_n = g->n, seg_i = 0;
        prepare(g, 0, 0, own);
        resolve();

        for (int node = 0; node < n_nodes; node++) {
            const int u = g->u0 + seg_i * g->du;
            const int par = node & 1;
            const REAL gamma = REAL(1) / REAL(g->gamma_den);
            const int halves = g->halves;
            const bool use_carry = g->use_carry != 0;
            // where the next node lives
            const bool seg_last = (seg_i + 1 >= seg_n);
            const bool has_next = (node + 1 < n_nodes);
            const SegWarp *gn = seg_last ? &s_desc[w][buf ^ 1] : g;
            const int in = seg_last ? 0 : seg_i + 1;

            for (int half = 0; half < halves; half++) {
                tick(6);
                if (half == 0) {
                    // ---------------- phase A: resolve pending dependencies, sum into partial rows
#pragma unroll
                    for (int q = 0; q < SLOTS; q++) {
                        if (slot_flag[q] >= 0) {
                            while (ld_flag(p.progress + slot_flag[q]) < slot_need[q]) __nanosleep(32);
                            if ((slot_kind[q] & 255) == S_DYN) row_async<REAL, K>(landing + (size_t)q * LP, p.msg + slot_term[q] * LP, lane);
                            else slot_sel[q] = __ldcg(p.selpos + slot_term[q]);
                            slot_flag[q] = -1;
                        }
                    }
                    tick(0);
                    cp_async_wait_all();
                    __syncwarp();
                    REAL rm[K], rx[K], rb[K];
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        rm[k] = (own.flags & OWN_HAS) ? own.m[k] : REAL(0);
                        rx[k] = REAL(0);
                        rb[k] = REAL(0);
                    }
#pragma unroll
                    for (int q = 0; q < SLOTS; q++) {
                        const int k0 = slot_kind[q] & 255;
                        if (k0 == S_NONE) continue;
                        REAL v[K];
                        row_lds<REAL, K>(v, landing + (size_t)q * LP, lane);
                        if (k0 == S_D) {
                            row_sts<REAL, K>(row_ptr(par, R_DR), v, lane);
                        } else if (k0 == S_STAT) {
#pragma unroll
                            for (int k = 0; k < K; k++) rm[k] += v[k];
                        } else if (k0 == S_DYN) {
#pragma unroll
                            for (int k = 0; k < K; k++) rx[k] += v[k];
                        } else {
#pragma unroll
                            for (int k = 0; k < K; k++) rb[k] += slot_alpha[q] * smooth<REAL, KERN>(v[k] - slot_sel[q], p.lambda);
                        }
                    }
                    row_sts<REAL, K>(row_ptr(par, R_RM + w), rm, lane);
                    if (do_send) row_sts<REAL, K>(row_ptr(par, R_RX + w), rx, lane);
                    if (do_round) row_sts<REAL, K>(row_ptr(par, R_RB + w), rb, lane);
                }
                tick(1);
                step_barrier();
                tick(2);
                // ---------------- issue the loads of the next step (part 1)
                OwnTerm<REAL, K> nxt;
                nxt.flags = 0;
                const bool last_half = (half + 1 == halves);
                if (!last_half) prepare(g, seg_i, half + 1, nxt);
                else if (has_next) prepare(gn, in, 0, nxt);
                tick(3);
                if (half == 0) {
                    // ---------------- phase B1: Di (and the rounding) from the partial rows
                    REAL dsum[K], msum[K];
                    row_lds<REAL, K>(dsum, row_ptr(par, R_DR), lane);
#pragma unroll
                    for (int k = 0; k < K; k++) msum[k] = REAL(0);
#pragma unroll
                    for (int ww = 0; ww < NCW; ww++) {
                        REAL v[K];
                        row_lds<REAL, K>(v, row_ptr(par, R_RM + ww), lane);
#pragma unroll
                        for (int k = 0; k < K; k++) msum[k] += v[k];
                    }
                    if (do_round) {
                        // minimize.cpp:240-260: DiB = D + sum_{lower nb} V(x_nb, .), Dr = DiB + forward messages
                        REAL dib[K];
#pragma unroll
                        for (int k = 0; k < K; k++) dib[k] = dsum[k];
#pragma unroll
                        for (int ww = 0; ww < NCW; ww++) {
                            REAL v[K];
                            row_lds<REAL, K>(v, row_ptr(par, R_RB + ww), lane);
#pragma unroll
                            for (int k = 0; k < K; k++) dib[k] += v[k];
                        }
                        if (use_carry) {
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                REAL v[K];
                                row_lds<REAL, K>(v, row_ptr(par, R_CC + jj), lane);
#pragma unroll
                                for (int k = 0; k < K; k++) dib[k] += v[k];
                            }
                        }
                        // Vector::ComputeMin: first minimum in label order (typeStereoLinear.h:238-252)
                        REAL best = BIG;
                        int bi = 0x7fffffff;
                        REAL bdib = REAL(0);
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const int lbl = lane * K + k;
                            const REAL dr = dib[k] + msum[k];
                            if (lbl < p.L && dr < best) { best = dr; bi = lbl; bdib = dib[k]; }
                        }
                        // first minimum over the warp: value, then the smallest label attaining it
                        const REAL wbest = warp_min(best);
                        bi = warp_min_s32(best == wbest ? bi : 0x7fffffff);
                        {
                            REAL dv = dib[0];
#pragma unroll
                            for (int k = 1; k < K; k++)
                                if (k == bi % K) dv = dib[k];
                            bdib = __shfl_sync(0xffffffffu, dv, bi / K);
                        }
                        xs = bi;
                        if (w == 0 && lane == 0) {
                            p.sol[u] = bi;
                            acc_energy += (double)bdib;
                        }
                    }
                    if (do_send) {
                        // Di = D + all incident messages (minimize.cpp:38-46 / 69-77)
#pragma unroll
                        for (int k = 0; k < K; k++) Di[k] = dsum[k] + msum[k];
#pragma unroll
                        for (int ww = 0; ww < NCW; ww++) {
                            REAL v[K];
                            row_lds<REAL, K>(v, row_ptr(par, R_RX + ww), lane);
#pragma unroll
                            for (int k = 0; k < K; k++) Di[k] += v[k];
                        }
                        if (use_carry) {
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                REAL v[K];
                                row_lds<REAL, K>(v, row_ptr(par, R_CM + jj), lane);
#pragma unroll
                                for (int k = 0; k < K; k++) Di[k] += v[k];
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
                            if (w == 0) acc_lb += (double)vmin;
                        }
                    }
                }

                // ---------------- issue the loads of the next step (part 2: flag-dependent rows)
                tick(4);
                if (last_half && has_next) resolve();
                tick(5);

                // ---------------- phase B5: min-plus update of this warp's term
                if (own.flags & OWN_HAS) {
                    const bool to_next = (own.flags & OWN_TO_NEXT) != 0;
                    const int oj = (own.flags & OWN_J) ? 1 : 0;
                    if (do_round) {
                        // position of the rounded label on this term, for the receiver's rounding
                        REAL sv = own.s[0];
#pragma unroll
                        for (int k = 1; k < K; k++)
                            if (k == xs % K) sv = own.s[k];
                        sv = __shfl_sync(0xffffffffu, sv, xs / K);
                        if (lane == 0) __stcg(p.selpos + own.term, sv);
                        if (to_next) {
                            REAL cc[K];
#pragma unroll
                            for (int k = 0; k < K; k++) cc[k] = own.alpha * smooth<REAL, KERN>(own.x[k] - sv, p.lambda);
                            row_sts<REAL, K>(row_ptr(par ^ 1, R_CC + oj), cc, lane);
                        }
                    }
                    if (do_send) {
                        REAL vmin = REAL(0);
                        if (SB_TRWS_INSTRUMENT && (p.debug & 2)) {
#pragma unroll
                            for (int k = 0; k < K; k++) own.m[k] = Di[k] * gamma - own.m[k];
                        } else if constexpr (KERN == 1)
                            vmin = update_linear<REAL, K>(gamma, own.alpha, p.lambda, p.L, lane, Di, own.m, own.s, own.rk, own.x, own.cn, P);
                        else
                            vmin = update_quadratic<REAL, K>(gamma, own.alpha, p.lambda, p.L, lane, Di, own.m, own.s, own.rk, own.x, own.cn, P);
                        if (!(SB_TRWS_INSTRUMENT && (p.debug & 1))) VecIO<REAL, K>::store(p.msg + own.term * LP + lane * K, own.m);
                        if (PASS == PASS_BWD) acc_lb += (double)vmin;
                        if (to_next) row_sts<REAL, K>(row_ptr(par ^ 1, R_CM + oj), own.m, lane);
                    }
                }
                own = nxt;
            }
            // this warp's stores for the node are issued: tell the auxiliary warp
            __syncwarp();
            if (lane == 0) st_release_cta(s_wdone + w, node + 1);
            if (prof_on) tprof[7] += 1;
            // advance to the next node / segment
            if (seg_last) {
                if (has_next) {
                    sg++;
                    buf ^= 1;
                    g = &s_desc[w][buf];
                    seg_n = g->n;
                    seg_i = 0;
                    if (sg + 1 < sg1) fetch_desc(sg + 1, buf ^ 1);
                }
            } else {
                seg_i++;
            }
        }
        if (prof_on && lane == 0) {
            tick(6);
            const int grp = (fs == 0) ? 0 : 1; // strip 0 is the boundary ring on regular grids
#pragma unroll
            for (int q = 0; q < 8; q++) atomicAdd((unsigned long long *)p.prof + grp * 8 + q, (unsigned long long)tprof[q]);
        }
    }
    if (!is_aux && lane == 0) {
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
