// gtrws_kernels.cuh -- sm_100a device code of the GRID-NATIVE TRW-S sweep behind
// sb_trws_grid_* (include/stereo_b200.h).
//
// Computes what Minimize_TRW_S computes (cpp/trw-s/minimize.cpp:31-95, primal rounding
// minimize.cpp:223-264, UpdateMessage typeStereoLinear.h:329-487 / typeStereoQuadratic.h:329-501)
// on the 4-connected dispmap_super grid, in the reference's node order, but from a layout in
// which the 4096 x 4096 x 256 problem exists:
//
//   * label positions are NOT stored per term (the reference keeps q and qprim per edge,
//     typeStereoLinear.h:274-311: 32 L bytes per term).  dispmap_super.m:180-183 builds them as
//       q(:, p)     = disparity of the HEAD node's planes at the head's own point
//       qprim(:, p) = disparity of the TAIL node's planes at the head's point
//     so per node and label three numbers suffice: own disparity and the disparity step per column /
//     per row (gx, gy):  q = own[head],  qprim = own[tail] +- g[tail];
//   * per node    [D | gx | own | gy] fp32 rows + the sort rank of own            (17 B / label)
//     per neighbour pair (two terms): two message rows + 6 byte rows (sort rank of each
//     term's tail positions, two merge-count rows per term)                       (2 x 14 B / label)
//     => 45 bytes per label and node, against 100 in the MATLAB-layout path;
//   * node ids are row-major and band-local: a rank of a multi-GPU run stores its column band (plus one halo
//     column per inner side) as a grid of its own width, and the sweep only ever sees that local grid;
//   * a message word carries its own validity: min-normalised messages are >= 0, so the SIGN BIT is
//     free, and every message slot is written exactly once per sending pass -- the sweep stores
//     messages with the sign bit = parity of the pass counter, and a receiver in another strip (another
//     SM, another GPU) simply polls the words of its labels until the sign matches.  No mailbox copy,
//     no flag, no fence (32-bit accesses are single-copy atomic);
//   * everything static a node needs arrives by TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx) issued by one lane two nodes ahead into a 3-stage ring; the term warps take their
//     operands from shared memory when the update runs, so they hold no prefetched operands in
//     registers and more strip walkers fit an SM.
//
// CTA = 4 term warps + 2 helper warps (alternating steps), one strip (the boundary ring, then one per interior row) at a
// time from an atomic ticket, one persistent launch per pass.
#pragma once
#include <type_traits>
#include "trws_kernels.cuh"
#include "gtrws_plan.h"

namespace sb {
namespace gtrws {

using trws::Pair;
using trws::Lim;
using trws::PASS_FWD;
using trws::PASS_BWD;
using trws::MODE_SEND;
using trws::MODE_ROUND;

// Diagnostics (SB_TRWS_PROFILE phase counters, SB_TRWS_RECORD flight recorder) cost ~8 % of the term warps'
// instructions even when switched off at run time (predicated-off instructions still issue), so they are compiled
// in only by  make EXTRA=-DSB_GTRWS_DIAG=1  (profiles/r2_phase_counters_cfg5.txt came from such a build).
#ifndef SB_GTRWS_DIAG
#define SB_GTRWS_DIAG SB_TRWS_DIAG
#endif
constexpr bool DIAG = SB_GTRWS_DIAG != 0;

constexpr int NTW = 4;                       // term warps
constexpr int NHW = 2;                       // helper warps: helper h prepares the steps of parity h
constexpr int CTA_THREADS = (NTW + NHW) * 32;
enum { NF_D = 0, NF_GX = 1, NF_OWN = 2, NF_GY = 3 };

template <typename REAL>
struct GProblem {
    int H, W, L, LP;
    int rows;                    // rows stored on this rank (band + halo)
    long long Nloc;              // rows * W
    const REAL *nodeF;           // [Nloc][4][LP]   D, gx, own, gy
    const uint8_t *nodeB;        // [Nloc][LP]      rank of own
    REAL *msg;                   // [2 Nloc pairs][2][LP]  one message per term, sign bit = pass parity
    const uint8_t *pairB;        // [pairs][2 sides][3][LP]
    const REAL *alpha;           // [pairs][2]
    unsigned long long *selbox;  // [pairs][2]  (the sender's rounded label | epoch << 32)
    REAL lambda;
    const GSeg *segs;
    const int32_t *seg_ptr, *strip_len, *is_ring;   // strips in processing order of the pass
    int S;
    unsigned long long *save;          // [save slots][LP][sizeof(REAL) / 4]: saved node totals (32 value bits | epoch << 32)
    int *ring_smid;                    // forward pass: 1 + SM id of the CTA that walks the first ring strip
    int isolate_ring;                  // forward pass: CTAs that share that SM leave (the ring is the serial bottleneck)
    int world;
    REAL *peer_msg[2];                 // rank - 1 / rank + 1 (peer mapped)
    unsigned long long *peer_selbox[2];
    int peer_dW[2];                    // the peer's local node id of my local node u = r W + c is u + r peer_dW + peer_dc
    long long peer_dc[2];              // (the bands differ in width and in their first stored column)
    int32_t *sol;                      // [Nloc]
    unsigned epoch;                    // launch counter: tag of the selbox words
    unsigned tag;                      // 0 / 1: sign the messages of this pass carry
    int *ticket;
    double *acc;                       // [0] energy [1] lower bound
    int mode;
    long long *prof;                   // optional cycle counters
    int prof_warp;                     // term warp they are taken on
    int head_strip;                    // profile: the first interior strip of the pass is reported on its own
    unsigned poll_ns;                  // back-off between two polls of a dependency that has not arrived
    int *rec;                          // SB_TRWS_RECORD: host-mapped flight recorder [cta][5 warps][4], else null
};

// ---------------------------------------------------------------- lane <-> label map
// Rows are label-contiguous in memory.  A lane owns its K labels in BLOCKS of 16 / 8 / 4 bytes: block b covers the
// labels [off_b, off_b + 32 w_b) and gives lane l the w_b labels off_b + w_b l ... -- so every vector access of a
// warp touches 32 consecutive chunks (no shared-memory bank conflicts, fully coalesced global sectors), where
// "K consecutive labels per lane" made 32-byte lane strides (2-way conflicts on every row read at K = 8).
template <typename REAL, int K> struct LaneMap {
    static constexpr int CW = 16 / (int)sizeof(REAL);     // labels per 16-byte chunk
    // width of the block that holds slot k, its first slot and its first label
    static __host__ __device__ constexpr int blk_w(int k)
    {
        int k0 = 0, rem = K;
        while (rem >= CW) { if (k < k0 + CW) return CW; k0 += CW; rem -= CW; }
        for (int w = CW / 2; w >= 1; w /= 2)
            if (rem >= w) { if (k < k0 + w) return w; k0 += w; rem -= w; }
        return 1;
    }
    static __host__ __device__ constexpr int blk_k0(int k)
    {
        int k0 = 0, rem = K;
        while (rem >= CW) { if (k < k0 + CW) return k0; k0 += CW; rem -= CW; }
        for (int w = CW / 2; w >= 1; w /= 2)
            if (rem >= w) { if (k < k0 + w) return k0; k0 += w; rem -= w; }
        return k0;
    }
    static __device__ __forceinline__ int label(int lane, int k) { return 32 * blk_k0(k) + blk_w(k) * lane + (k - blk_k0(k)); }
    // (lane, slot) that own label x
    static __device__ __forceinline__ void owner(int x, int &lane, int &slot)
    {
        lane = 0; slot = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (k == blk_k0(k)) {   // first slot of a block
                const int w = blk_w(k), off = 32 * k;
                if (x >= off && x < off + 32 * w) { lane = (x - off) / w; slot = k + (x - off) % w; }
            }
        }
    }
    static __device__ __forceinline__ unsigned valid_mask(int lane, int L)
    {
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < K; k++) v |= (label(lane, k) < L ? 1u : 0u) << k;
        return v;
    }
};

// compile-time loop over the slots 0 .. K-1 (the block structure is a compile-time property of the slot)
template <int K0, int KEND, typename F> __device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (K0 < KEND) {
        f(std::integral_constant<int, K0>{});
        static_for<K0 + 1, KEND>(f);
    }
}

// this lane's K values of a shared-memory row / to one
template <typename REAL, int K> __device__ __forceinline__ void lrow_lds(REAL (&r)[K], const REAL *row, int lane)
{
    typedef LaneMap<REAL, K> LM;
    static_for<0, K>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        if constexpr (k == LM::blk_k0(k)) {
            constexpr int w = LM::blk_w(k);
            constexpr int bytes = w * (int)sizeof(REAL);
            const REAL *q = row + 32 * k + w * lane;
            if constexpr (bytes == 16) {
                const float4 v = *reinterpret_cast<const float4 *>(q);
                if constexpr (sizeof(REAL) == 4) { r[k] = v.x; r[k + 1] = v.y; r[k + 2] = v.z; r[k + 3] = v.w; }
                else {
                    r[k] = __hiloint2double(__float_as_int(v.y), __float_as_int(v.x));
                    r[k + 1] = __hiloint2double(__float_as_int(v.w), __float_as_int(v.z));
                }
            } else if constexpr (bytes == 8) {
                const float2 v = *reinterpret_cast<const float2 *>(q);
                if constexpr (sizeof(REAL) == 4) { r[k] = v.x; r[k + 1] = v.y; }
                else r[k] = __hiloint2double(__float_as_int(v.y), __float_as_int(v.x));
            } else {
                r[k] = *q;
            }
        }
    });
}
template <typename REAL, int K> __device__ __forceinline__ void lrow_sts(REAL *row, const REAL (&r)[K], int lane)
{
    typedef LaneMap<REAL, K> LM;
    static_for<0, K>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        if constexpr (k == LM::blk_k0(k)) {
            constexpr int w = LM::blk_w(k);
            constexpr int bytes = w * (int)sizeof(REAL);
            REAL *q = row + 32 * k + w * lane;
            if constexpr (bytes == 16) {
                float4 v;
                if constexpr (sizeof(REAL) == 4) { v.x = r[k]; v.y = r[k + 1]; v.z = r[k + 2]; v.w = r[k + 3]; }
                else {
                    v.x = __int_as_float(__double2loint(r[k])); v.y = __int_as_float(__double2hiint(r[k]));
                    v.z = __int_as_float(__double2loint(r[k + 1])); v.w = __int_as_float(__double2hiint(r[k + 1]));
                }
                *reinterpret_cast<float4 *>(q) = v;
            } else if constexpr (bytes == 8) {
                float2 v;
                if constexpr (sizeof(REAL) == 4) { v.x = r[k]; v.y = r[k + 1]; }
                else { v.x = __int_as_float(__double2loint(r[k])); v.y = __int_as_float(__double2hiint(r[k])); }
                *reinterpret_cast<float2 *>(q) = v;
            } else {
                *q = r[k];
            }
        }
    });
}
// this lane's K bytes of a shared-memory byte row
template <typename REAL, int K> __device__ __forceinline__ void lrow_ldb(uint8_t (&r)[K], const uint8_t *row, int lane)
{
    typedef LaneMap<REAL, K> LM;
    static_for<0, K>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        if constexpr (k == LM::blk_k0(k)) {
            constexpr int w = LM::blk_w(k);
            const uint8_t *q = row + 32 * k + w * lane;
            if constexpr (w == 4) {
                const unsigned v = *reinterpret_cast<const unsigned *>(q);
                r[k] = (uint8_t)v; r[k + 1] = (uint8_t)(v >> 8); r[k + 2] = (uint8_t)(v >> 16); r[k + 3] = (uint8_t)(v >> 24);
            } else if constexpr (w == 2) {
                const unsigned short v = *reinterpret_cast<const unsigned short *>(q);
                r[k] = (uint8_t)v; r[k + 1] = (uint8_t)(v >> 8);
            } else {
                r[k] = *q;
            }
        }
    });
}

// ---------------------------------------------------------------- tagged message words
template <typename REAL> struct Tag;
template <> struct Tag<float> {
    typedef unsigned word;
    static __device__ __forceinline__ float put(float v, unsigned tag) { return __uint_as_float((__float_as_uint(v) & 0x7fffffffu) | (tag << 31)); }
    static __device__ __forceinline__ bool ok(float v, unsigned tag) { return (__float_as_uint(v) >> 31) == tag; }
    static __device__ __forceinline__ float val(float v) { return fabsf(v); }
};
template <> struct Tag<double> {
    static __device__ __forceinline__ double put(double v, unsigned tag)
    {
        return __hiloint2double((__double2hiint(v) & 0x7fffffff) | (int)(tag << 31), __double2loint(v));
    }
    static __device__ __forceinline__ bool ok(double v, unsigned tag) { return ((unsigned)__double2hiint(v) >> 31) == tag; }
    static __device__ __forceinline__ double val(double v) { return fabs(v); }
};

// This lane's K message words of a row in global memory (LaneMap blocks), L2-coherent relaxed accesses: the words
// are written by other SMs / GPUs during this launch, and each word validates itself, so per-word atomicity suffices.
template <int BYTES, bool SYS> __device__ __forceinline__ void ld_relaxed(unsigned (&w)[BYTES / 4], const void *p)
{
    if constexpr (BYTES == 16) {
        if constexpr (SYS) asm volatile("ld.relaxed.sys.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p) : "memory");
        else asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p) : "memory");
    } else if constexpr (BYTES == 8) {
        if constexpr (SYS) asm volatile("ld.relaxed.sys.global.v2.b32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p) : "memory");
        else asm volatile("ld.relaxed.gpu.global.v2.b32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p) : "memory");
    } else {
        if constexpr (SYS) asm volatile("ld.relaxed.sys.global.b32 %0, [%1];" : "=r"(w[0]) : "l"(p) : "memory");
        else asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(w[0]) : "l"(p) : "memory");
    }
}
template <int BYTES, bool SYS> __device__ __forceinline__ void st_relaxed(void *p, const unsigned (&w)[BYTES / 4])
{
    if constexpr (BYTES == 16) {
        if constexpr (SYS) asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        else asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    } else if constexpr (BYTES == 8) {
        if constexpr (SYS) asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(w[0]), "r"(w[1]) : "memory");
        else asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(w[0]), "r"(w[1]) : "memory");
    } else {
        if constexpr (SYS) asm volatile("st.relaxed.sys.global.b32 [%0], %1;" ::"l"(p), "r"(w[0]) : "memory");
        else asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(w[0]) : "memory");
    }
}
template <typename REAL> __device__ __forceinline__ REAL words_to_real(const unsigned *w)
{
    if constexpr (sizeof(REAL) == 4) return __uint_as_float(w[0]);
    else return __hiloint2double((int)w[1], (int)w[0]);
}
template <typename REAL> __device__ __forceinline__ void real_to_words(unsigned *w, REAL v)
{
    if constexpr (sizeof(REAL) == 4) w[0] = __float_as_uint(v);
    else { w[0] = (unsigned)__double2loint(v); w[1] = (unsigned)__double2hiint(v); }
}
// `row` = first label of the row
template <typename REAL, int K, bool SYS> __device__ __forceinline__ void ld_words(REAL (&r)[K], const REAL *row, int lane)
{
    typedef LaneMap<REAL, K> LM;
    constexpr int WR = (int)sizeof(REAL) / 4;
    static_for<0, K>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        if constexpr (k == LM::blk_k0(k)) {
            constexpr int w = LM::blk_w(k);
            constexpr int NW = w * WR;          // 32-bit words of this block per lane: 4, 2 or 1
            const REAL *q = row + 32 * k + w * lane;
            unsigned u[NW];
            ld_relaxed<NW * 4, SYS>(u, q);
#pragma unroll
            for (int i = 0; i < w; i++) r[k + i] = words_to_real<REAL>(u + i * WR);
        }
    });
}
template <typename REAL, int K, bool SYS> __device__ __forceinline__ void st_words(REAL *row, const REAL (&r)[K], int lane)
{
    typedef LaneMap<REAL, K> LM;
    constexpr int WR = (int)sizeof(REAL) / 4;
    static_for<0, K>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        if constexpr (k == LM::blk_k0(k)) {
            constexpr int w = LM::blk_w(k);
            constexpr int NW = w * WR;
            REAL *q = row + 32 * k + w * lane;
            unsigned u[NW];
#pragma unroll
            for (int i = 0; i < w; i++) real_to_words<REAL>(u + i * WR, r[k + i]);
            st_relaxed<NW * 4, SYS>(q, u);
        }
    });
}

// Saved node totals (GF_SAVE -> GF_DEFERRED): 32 value bits + launch epoch per 64-bit word, so a word validates
// itself (doubles travel as two words).
template <typename REAL, int K> __device__ __forceinline__ void save_total(unsigned long long *slot, const REAL (&v)[K], int lane, unsigned epoch)
{
    constexpr int WPR = (int)sizeof(REAL) / 4;
    unsigned long long *p = slot + (size_t)lane * K * WPR;
#pragma unroll
    for (int k = 0; k < K; k++) {
        if constexpr (WPR == 1) {
            trws::st_mbox(p + k, (unsigned long long)__float_as_uint((float)v[k]) | ((unsigned long long)epoch << 32));
        } else {
            trws::st_mbox(p + 2 * k, (unsigned long long)(unsigned)__double2loint((double)v[k]) | ((unsigned long long)epoch << 32));
            trws::st_mbox(p + 2 * k + 1, (unsigned long long)(unsigned)__double2hiint((double)v[k]) | ((unsigned long long)epoch << 32));
        }
    }
}
template <typename REAL, int K> __device__ __forceinline__ void poll_total(const unsigned long long *slot, REAL (&v)[K], int lane, unsigned epoch)
{
    constexpr int WPR = (int)sizeof(REAL) / 4;
    const unsigned long long *p = slot + (size_t)lane * K * WPR;
    unsigned long long w[K * WPR];
    for (;;) {
        bool ok = true;
#pragma unroll
        for (int q = 0; q < K * WPR; q++) {
            w[q] = trws::ld_mbox(p + q);
            ok = ok && ((unsigned)(w[q] >> 32) == epoch);
        }
        if (__all_sync(0xffffffffu, ok)) break;
        __nanosleep(200);
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        if constexpr (WPR == 1) v[k] = (REAL)__uint_as_float((unsigned)w[k]);
        else v[k] = (REAL)__hiloint2double((int)(unsigned)w[2 * k + 1], (int)(unsigned)w[2 * k]);
    }
}

// ---------------------------------------------------------------- mbarrier / TMA bulk copy
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, SASS UBLKCP), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 1 / d for the small integers gamma's denominators are (treeProbabilities.cpp:35-45: at most 8 terms per side):
// correctly rounded compile-time constants instead of a division (MUFU.RCP + fix-up call) on the dependent chain
template <typename REAL> __device__ __forceinline__ REAL inv_small(int d)
{
    // (a constant-bank table: one indexed LDC)
    static constexpr REAL tab[16] = {REAL(0),          REAL(1),          REAL(1) / REAL(2),  REAL(1) / REAL(3),
                                     REAL(1) / REAL(4),  REAL(1) / REAL(5),  REAL(1) / REAL(6),  REAL(1) / REAL(7),
                                     REAL(1) / REAL(8),  REAL(1) / REAL(9),  REAL(1) / REAL(10), REAL(1) / REAL(11),
                                     REAL(1) / REAL(12), REAL(1) / REAL(13), REAL(1) / REAL(14), REAL(1) / REAL(15)};
    return tab[d & 15];
}

// ---------------------------------------------------------------- geometry of a node step
struct StepGeo {
    int u;
    unsigned roles;
    int gamma_den, next_dir, flags, save;
    unsigned peer;   // 8 bits per direction
};
struct SegWalker {
    const GSeg *segs;
    int sg, i, n, u0, du, save0;
    unsigned roles, peer;
    int gamma_den, next_dir, flags;
    __device__ __forceinline__ void load()
    {
        const int4 a = __ldg(reinterpret_cast<const int4 *>(segs + sg));
        const int4 b = __ldg(reinterpret_cast<const int4 *>(segs + sg) + 1);
        u0 = a.x; du = a.y; n = a.z; roles = (unsigned)a.w;
        gamma_den = (int)(short)(b.x & 0xffff);
        next_dir = (int)(signed char)((b.x >> 16) & 0xff);
        flags = (b.x >> 24) & 0xff;
        peer = (unsigned)b.y;
        save0 = b.z;
        i = 0;
    }
    __device__ __forceinline__ void init(const GSeg *s, int sg0) { segs = s; sg = sg0; load(); }
    __device__ __forceinline__ void get(StepGeo &g) const
    {
        g.u = u0 + i * du; g.roles = roles; g.gamma_den = gamma_den; g.next_dir = next_dir; g.flags = flags; g.peer = peer;
        g.save = save0 + i;
    }
    // to the next step (the caller guarantees there is one)
    __device__ __forceinline__ void advance()
    {
        if (++i >= n) { sg++; load(); }
    }
};
__device__ __forceinline__ int role_of(unsigned roles, int d) { return (roles >> (4 * d)) & 15; }
// first direction whose role nibble equals `role`, else -1.  Branch-free: zero-nibble search in roles ^ (role x 0x1111);
// the borrow of x - 0x1111 can only raise false flags ABOVE a true zero nibble, and the lowest flag is taken.
__device__ __forceinline__ int dir_with_role(unsigned roles, int role)
{
    const unsigned x = (roles & 0xffffu) ^ ((unsigned)role * 0x1111u);
    const unsigned t = (x - 0x1111u) & ~x & 0x8888u;
    return ((int)__ffs((int)t) - 1) >> 2;          // t == 0: (0 - 1) >> 2 = -1
}
// (DIR_UP = 0, DIR_DOWN = 1, DIR_LEFT = 2, DIR_RIGHT = 3: bit 1 = horizontal, bit 0 = towards larger ids)
__device__ __forceinline__ long long nb_of(long long u, int d, int W)
{
    const int step = (d & 2) ? 1 : W;
    return (d & 1) ? u + step : u - step;
}
// pair record of the neighbour pair in direction d: pairs are owned by their upper / left node
__device__ __forceinline__ long long pair_of(long long u, int d, int W)
{
    const int step = (d & 2) ? 1 : W;
    return 2 * ((d & 1) ? u : u - step) + ((d >> 1) & 1);
}
// 0: I am the pair's first (upper / left) node, 1: its second.  Term j of a pair has tail = node j.
__device__ __forceinline__ int side_of(int d) { return (d & 1) ^ 1; }
__device__ __forceinline__ bool vertical(int d) { return (d & 2) == 0; }
static_assert(DIR_UP == 0 && DIR_DOWN == 1 && DIR_LEFT == 2 && DIR_RIGHT == 3, "direction encoding");

// index of my term `term` (= 4 x owner node + 2 x [horizontal] + j) in the arrays of neighbour rank `peer`
template <typename REAL> __device__ __forceinline__ long long peer_term(const GProblem<REAL> &p, long long term, int peer)
{
    const long long uo = term >> 2;
    return term + 4 * ((uo / p.W) * (long long)p.peer_dW[peer] + p.peer_dc[peer]);
}

// ---------------------------------------------------------------- shared memory plan
template <typename REAL, int K> struct StageLayout {
    static constexpr int LP = 32 * K;
    static constexpr int ROW = LP * (int)sizeof(REAL);
    static constexpr int OFF_NF = 0;                       // 4 rows
    static constexpr int OFF_MS = OFF_NF + 4 * ROW;        // 2 slots x 2 rows
    static constexpr int OFF_XN = OFF_MS + 4 * ROW;        // 2 slots x 2 rows
    static constexpr int OFF_NB = OFF_XN + 4 * ROW;        // 1 byte row
    static constexpr int OFF_PB = OFF_NB + LP;             // 2 slots x 3 byte rows
    static constexpr int BYTES = OFF_PB + 6 * LP;
    static_assert(BYTES % 16 == 0, "stage alignment");
};
enum { R_BASE = 0, R_DIB0 = 1, R_RMS = 2 };
template <typename REAL, int K, int NS> __host__ __device__ constexpr size_t gsweep_smem_bytes()
{
    // stages + 2 sets of {BASE, DIB0, RMS} + 2 x {CM0, CM1, CC0, CC1} + DI_SAVE + scratch pairs + mbarriers
    return (size_t)NS * StageLayout<REAL, K>::BYTES + (size_t)(6 + 8 + 1) * 32 * K * sizeof(REAL) +
           (size_t)NTW * trws::scratch_pairs<K>() * sizeof(Pair<REAL>) + (2 * NS + 2) * 8;
}

// the four term warps among themselves (carry rows of the previous step are complete)
template <int ID> __device__ __forceinline__ void term_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(NTW * 32) : "memory"); }

// resident CTAs per SM the register budget of the throughput build is set for (setmaxnreg re-division between the term
// warpgroup and the helpers was tried: ptxas 12.9 keeps allocating under the launch cap and spills, so the budget is
// uniform).  The register file is per scheduler (16 K registers): 3 CTAs x 6 warps put 5 warps on one scheduler => 96
// registers per thread, 2 CTAs => 168, 4 CTAs => 80.
template <typename REAL, int K> __host__ __device__ constexpr int gsweep_min_blocks()
{
    if (sizeof(REAL) == 8) return K <= 2 ? 2 : 1;
    // measured on a B200 (1980x2880x192 / 1024x2048x256): one CTA more per SM with ~1-2 KB of spills is 20 % slower
    // (K <= 2: 5 CTAs per SM cost 70-120 spill instructions and were 7-14 % slower than 4 at 375x450x64 / 1080x1920x64;
    // 3 instead of 4 loses 2-4 % at K = 3, 4; 2 instead of 3 loses 4 % at K = 6; 1 instead of 2 loses 16 % at K = 8)
    return K <= 4 ? 4 : K <= 6 ? 3 : 2;
}

// MB = resident CTAs per SM the registers are budgeted for: gsweep_min_blocks (throughput build: as many strip walkers per
// SM as pay) or at most 2 (latency build: <= 168 registers, no spills, operands taken before the step's waits -- shorter
// node steps for passes that run at the DAG's critical path, i.e. the column-banded multi-GPU sweeps; gtrws_solve.cu picks).
// Measured at K = 6 (1980 x 2880 x 192 / its 360-column band alone): MB 3: 38.1 / 12.1 ms per pass, MB 2 + early operands:
// 41.5 / 11.1 ms, MB 1 + early operands: 63.1 / 12.3 ms.
template <typename REAL, int K, int KERN, int PASS, int NS, int MB>
__global__ void __launch_bounds__(CTA_THREADS, MB) gsweep_kernel(const GProblem<REAL> p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_ticket;
    typedef StageLayout<REAL, K> SL;
    constexpr int LP = 32 * K;
    constexpr int PD = NS - 1;   // bulk copies are issued PD steps ahead
    // operands ahead of the chain (take_operands) where the registers hold them without spills: the latency builds and the
    // throughput builds of K <= 2 and K = 8; with 100-340 B of spills (K = 4, 6 under the 3- / 4-CTA caps) it LOSES 5-20 %
    constexpr bool EARLY = sizeof(REAL) == 4 && (MB <= 2 || K <= 2);
    const int lane = threadIdx.x & 31;
    // (warp w issues from scheduler w & 3: the term warps 2, 3 of send slot 1 -- the sends to the next node of the strip,
    // i.e. the dependent chain, gtrws_plan.cpp -- have their schedulers to themselves; the helpers share with slot 0)
    const int warp = threadIdx.x >> 5;
    const bool is_term = warp < NTW;
    const int W = p.W;
    const REAL BIG = Lim<REAL>::big();
    const bool do_send = (PASS == PASS_BWD) || (p.mode & MODE_SEND);
    const bool do_round = (PASS == PASS_FWD) && (p.mode & MODE_ROUND);
    const unsigned tag = p.tag;

    unsigned char *stage_base = smem_raw;
    REAL *rows = reinterpret_cast<REAL *>(smem_raw + (size_t)NS * SL::BYTES);
    auto stage_ptr = [&](int st) -> unsigned char * { return stage_base + (size_t)st * SL::BYTES; };
    auto set_ptr = [&](int set, int r) -> REAL * { return rows + (size_t)(set * 3 + r) * LP; };
    auto carry_ptr = [&](int par, int r) -> REAL * { return rows + (size_t)(6 + par * 4 + r) * LP; };
    REAL *di_save = rows + (size_t)14 * LP;
    Pair<REAL> *P = reinterpret_cast<Pair<REAL> *>(rows + (size_t)15 * LP) + (size_t)(is_term ? warp : 0) * trws::scratch_pairs<K>();
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<unsigned char *>(rows + (size_t)15 * LP) + (size_t)NTW * trws::scratch_pairs<K>() * sizeof(Pair<REAL>));
    unsigned long long *bar_full = bars, *bar_free = bars + NS, *bar_base = bars + 2 * NS;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_free + s, NTW);
        }
        mbar_init(bar_base + 0, 1);
        mbar_init(bar_base + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (is_term && lane == 0) {
        Pair<REAL> t;
        t.a = BIG;
        t.b = REAL(0);
        P[0] = t;
        P[trws::phys<K>(LP)] = t;
    }
    __syncthreads();
    double acc_energy = 0.0, acc_lb = 0.0;
    // flight recorder (debugging hangs): where every warp is -- strip, step, phase -- in host-mapped memory
    auto record = [&](int strip, int step, int phase, int extra) {
        if constexpr (!DIAG) return;
        if (p.rec && lane == 0) {
            volatile int *r = p.rec + ((size_t)blockIdx.x * (NTW + NHW) + warp) * 4;
            r[0] = strip; r[1] = step; r[2] = phase; r[3] = extra;
        }
    };
    unsigned gstep0 = 0;    // node steps this CTA has walked before the current strip (stage ring position); 32-bit:
                            // the stage / phase arithmetic below is a division by NS per use

    // Forward pass: the first strip is the boundary ring, a serial chain everything else waits for.  CTA 0 takes it
    // without a ticket and publishes its SM; CTAs that landed on the same SM leave, so the chain has the SM alone.
    bool first_round = true;
    if (PASS == PASS_FWD && p.isolate_ring) {
        __shared__ int s_leave;
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_leave = 0;
            if (blockIdx.x == 0) {
                *reinterpret_cast<volatile int *>(p.ring_smid) = (int)smid + 1;
            } else {
                int v;
                while ((v = *reinterpret_cast<volatile int *>(p.ring_smid)) == 0) __nanosleep(100);
                s_leave = (v == (int)smid + 1);
            }
        }
        __syncthreads();
        if (s_leave) return;
    }
    for (;;) {
        if (threadIdx.x == 0) {
            if (PASS == PASS_FWD && p.isolate_ring) s_ticket = (first_round && blockIdx.x == 0) ? 0 : atomicAdd(p.ticket, 1) + 1;
            else s_ticket = atomicAdd(p.ticket, 1);
        }
        first_round = false;
        __syncthreads();
        const int ts = s_ticket;
        __syncthreads();
        if (ts >= p.S) break;
        const int fs = ts;   // strips are listed in processing order
        const bool ring_strip = __ldg(p.is_ring + fs) != 0;
        const int sg0 = __ldg(p.seg_ptr + fs);
        const int n_steps = __ldg(p.strip_len + fs);
        if (n_steps <= 0) continue;

        if (is_term) {
            // ============================================================ term warps
            const int w = warp;
            const int slot = w >> 1, j = w & 1;
            SegWalker wk;
            wk.init(p.segs, sg0);
            int xs = 0;
            const unsigned valid = LaneMap<REAL, K>::valid_mask(lane, p.L);
            // optional phase timers (SB_TRWS_PROFILE): term warp 0 -> wait FULL, wait stage, node total + rounding,
            // operands, update, stores
            const bool prof_on = DIAG && (p.prof != nullptr) && w == p.prof_warp;
            long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tclk = prof_on ? clock64() : 0;
            auto tick = [&](int which) {
                if (prof_on) {
                    const long long now = clock64();
                    tp[which] += now - tclk;
                    tclk = now;
                }
            };
            // stage and phase of the CTA-lifetime step count gs = gstep0 + node, kept incrementally (gs % NS, gs / NS)
            int st = (int)(gstep0 % NS);
            unsigned ph = (unsigned)((gstep0 / NS) & 1);
            for (int node = 0; node < n_steps; node++, ph ^= (unsigned)(st + 1 == NS), st = (st + 1 == NS) ? 0 : st + 1) {
                const unsigned gs = gstep0 + (unsigned)node;
                const int par = (int)(gs & 1u);     // sets / carry rows / hand-over barriers alternate with the CTA's step count
                StepGeo g;
                wk.get(g);
                if (node + 1 < n_steps) wk.advance();
                const REAL gamma = inv_small<REAL>(g.gamma_den);
                const unsigned char *sp = stage_ptr(st);
                const REAL *NF = reinterpret_cast<const REAL *>(sp + SL::OFF_NF);
                // ---- my send term: pair in send slot `slot`, term j of it
                const int d = dir_with_role(g.roles, ROLE_SEND0 + slot);
                const bool vert = vertical(d);
                const int sd = side_of(d);
                const bool tail = (j == sd);
                const bool to_next = (d == g.next_dir);
                const long long pair = pair_of(g.u, d < 0 ? 0 : d, W);
                const long long term = 2 * pair + j;
                const int peer = d < 0 ? -1 : (int)((g.peer >> (8 * d)) & 255u) - 1;   // -1 local, 0 rank - 1, 1 rank + 1
                REAL alpha = REAL(0);
                if (d >= 0) alpha = __ldg(p.alpha + term);
                // operands of the update, from the stage ring
                REAL m[K], s[K], x[K];
                uint8_t rk[K], cn[K];
                auto take_operands = [&]() {
                    const REAL *MS = reinterpret_cast<const REAL *>(sp + SL::OFF_MS) + (size_t)(slot * 2 + j) * LP;
                    lrow_lds<REAL, K>(m, MS, lane);
#pragma unroll
                    for (int k = 0; k < K; k++) m[k] = Tag<REAL>::val(m[k]);
                    REAL own_me[K], g_me[K], own_nb[K], g_nb[K];
                    lrow_lds<REAL, K>(own_me, NF + NF_OWN * LP, lane);
                    const REAL *nb_own, *nb_g;
                    if (to_next) {
                        // the receiver is the next node of the strip: its rows are (or will shortly be) in the next stage
                        const bool wrap = st + 1 == NS;
                        const int stn = wrap ? 0 : st + 1;
                        record(fs, node, 4, stn);
                        mbar_wait(bar_full + stn, ph ^ (unsigned)wrap);
                        record(fs, node, 5, stn);
                        const REAL *NFn = reinterpret_cast<const REAL *>(stage_ptr(stn) + SL::OFF_NF);
                        nb_own = NFn + NF_OWN * LP;
                        nb_g = NFn + (vert ? NF_GY : NF_GX) * LP;
                    } else {
                        const REAL *XN = reinterpret_cast<const REAL *>(sp + SL::OFF_XN) + (size_t)(slot * 2) * LP;
                        nb_own = XN + (vert ? 0 : 1) * LP;
                        nb_g = XN + (vert ? 1 : 0) * LP;
                    }
                    lrow_lds<REAL, K>(own_nb, nb_own, lane);
                    const uint8_t *PB = sp + SL::OFF_PB + (size_t)slot * 3 * LP;
                    if (tail) {
                        // my positions: my planes at the receiver's point (qprim); receiver's: its own (q)
                        lrow_lds<REAL, K>(g_me, NF + (vert ? NF_GY : NF_GX) * LP, lane);
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            s[k] = sd == 0 ? own_me[k] + g_me[k] : own_me[k] - g_me[k];
                            x[k] = own_nb[k];
                        }
                        lrow_ldb<REAL, K>(rk, PB + 2 * LP, lane);
                        lrow_ldb<REAL, K>(cn, PB + 0 * LP, lane);
                    } else {
                        // I am the head: my own disparities (q); receiver's planes at my point (qprim)
                        lrow_lds<REAL, K>(g_nb, nb_g, lane);
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            s[k] = own_me[k];
                            x[k] = sd == 0 ? own_nb[k] - g_nb[k] : own_nb[k] + g_nb[k];
                        }
                        lrow_ldb<REAL, K>(rk, sp + SL::OFF_NB, lane);
                        lrow_ldb<REAL, K>(cn, PB + 1 * LP, lane);
                    }
                };
                // Latency build: the operands do not depend on the chain (carry rows, hand-over), so they are taken
                // BEFORE the step's waits -- the stage was filled two steps ago -- and only the node total, the update
                // and the stores remain behind the previous node.  (The throughput build has no registers to hold them.)
                if constexpr (EARLY) {
                    if (d >= 0) {
                        mbar_wait(bar_full + st, ph);
                        take_operands();
                    }
                }
                // ---- rows of this node are ready (helper), every term warp has finished the previous step
                tick(6);
                record(fs, node, 1, (int)g.roles);
                if (par) term_sync<2>(); else term_sync<1>();
                tick(0);
                record(fs, node, 2, st);
                // the helper's rows of this step; it waited for the stage's bulk copies before it built them, so they
                // are visible here as well (complete_tx -> helper's wait -> its arrive -> this wait)
                mbar_wait(bar_base + par, (unsigned)((gs >> 1) & 1u));
                record(fs, node, 3, st);
                tick(1);
                REAL Di[K];
                if (g.flags & GF_DEFERRED) {
                    // deferred sends of a node that sends on more than two pairs: its total, as saved by the main strip
                    lrow_lds<REAL, K>(Di, set_ptr(par, R_BASE), lane);
                } else {
                    lrow_lds<REAL, K>(Di, set_ptr(par, R_BASE), lane);
                    const int cd = dir_with_role(g.roles, ROLE_CARRY);
                    if (cd >= 0 && do_send) {
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            REAL v[K];
                            lrow_lds<REAL, K>(v, carry_ptr(par, jj), lane);
#pragma unroll
                            for (int k = 0; k < K; k++) Di[k] += v[k];
                        }
                    }
                    if (do_round) {
                        // minimize.cpp:240-260: DiB = D + sum_{lower nb} V(x_nb, .), Dr = DiB + forward messages
                        REAL dib[K], rms[K];
                        lrow_lds<REAL, K>(dib, set_ptr(par, R_DIB0), lane);
                        lrow_lds<REAL, K>(rms, set_ptr(par, R_RMS), lane);
                        if (cd >= 0) {
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                REAL v[K];
                                lrow_lds<REAL, K>(v, carry_ptr(par, 2 + jj), lane);
#pragma unroll
                                for (int k = 0; k < K; k++) dib[k] += v[k];
                            }
                        }
                        // Vector::ComputeMin: first minimum in label order (typeStereoLinear.h:238-252)
                        REAL best = BIG;
                        int bi = 0x7fffffff;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const int lbl = LaneMap<REAL, K>::label(lane, k);
                            const REAL dr = dib[k] + rms[k];
                            if (lbl < p.L && dr < best) { best = dr; bi = lbl; }
                        }
                        const REAL wbest = trws::warp_min(best);
                        bi = trws::warp_min_s32(best == wbest ? bi : 0x7fffffff);
                        xs = bi;
                        if (w == 0) {
                            int ol, os;
                            LaneMap<REAL, K>::owner(bi, ol, os);
                            REAL dv = dib[0];
#pragma unroll
                            for (int k = 1; k < K; k++)
                                if (k == os) dv = dib[k];
                            dv = __shfl_sync(0xffffffffu, dv, ol);
                            if (lane == 0) {
                                p.sol[g.u] = bi;
                                acc_energy += (double)dv;
                            }
                        }
                    }
                    if (do_send && PASS == PASS_BWD) {
                        // ComputeAndSubtractMin + lower bound (minimize.cpp:79-81)
                        REAL vmin = BIG;
#pragma unroll
                        for (int k = 0; k < K; k++)
                            if ((valid >> k) & 1u) vmin = min(vmin, Di[k]);
                        vmin = trws::warp_min(vmin);
#pragma unroll
                        for (int k = 0; k < K; k++) Di[k] -= vmin;
                        if (w == 0) acc_lb += (double)vmin;
                    }
                    if ((g.flags & GF_SAVE) && w == 0) save_total<REAL, K>(p.save + (size_t)g.save * LP * (sizeof(REAL) / 4), Di, lane, p.epoch);
                }
                if (prof_on) { tclk += (long long)(Di[0] != Di[0]); tick(2); }
                if (d >= 0) {
                    if constexpr (!EARLY) take_operands();
                    // the stage of this step is no longer needed by this warp
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_free + st);
                    if (prof_on) { tclk += (long long)(x[0] != x[0]) + (long long)(m[0] != m[0]); tick(3); }
                    if (do_round) {
                        // position of the rounded label on this term, for the receiver's rounding
                        int ol, os;
                        LaneMap<REAL, K>::owner(xs, ol, os);
                        REAL sv = s[0];
#pragma unroll
                        for (int k = 1; k < K; k++)
                            if (k == os) sv = s[k];
                        sv = __shfl_sync(0xffffffffu, sv, ol);
                        if (lane == 0 && !to_next) {
                            // fp32: the word carries the position itself; fp64 (64 value bits do not fit beside the
                            // tag): the label, and the receiver recomputes the position from my plane rows
                            unsigned lo;
                            if constexpr (sizeof(REAL) == 4) lo = __float_as_uint((float)sv); else lo = (unsigned)xs;
                            const unsigned long long wv = (unsigned long long)lo | ((unsigned long long)p.epoch << 32);
                            if (peer >= 0) trws::st_mbox_sys(p.peer_selbox[peer] + peer_term(p, term, peer), wv);
                            else trws::st_mbox(p.selbox + term, wv);
                        }
                        if (to_next) {
                            REAL cc[K];
#pragma unroll
                            for (int k = 0; k < K; k++) cc[k] = alpha * trws::smooth<REAL, KERN>(x[k] - sv, p.lambda);
                            lrow_sts<REAL, K>(carry_ptr(par ^ 1, 2 + j), cc, lane);
                        }
                    }
                    if (do_send) {
                        REAL vmin;
                        if constexpr (KERN == 1) {
                            // (L == 32 K, the usual case: the copy of the update without the label-validity tests)
                            if (p.L == LP) vmin = trws::update_linear<REAL, K, true>(gamma, alpha, p.lambda, valid, lane, Di, m, s, rk, x, cn, P);
                            else vmin = trws::update_linear<REAL, K, false>(gamma, alpha, p.lambda, valid, lane, Di, m, s, rk, x, cn, P);
                        } else
                            vmin = trws::update_quadratic<REAL, K>(gamma, alpha, p.lambda, valid, p.L, lane, Di, m, s, rk, x, cn, P);
                        if (PASS == PASS_BWD) acc_lb += (double)vmin;
                        if (prof_on) { tclk += (long long)(m[0] != m[0]); tick(4); }
                        if (to_next) lrow_sts<REAL, K>(carry_ptr(par ^ 1, j), m, lane);
#pragma unroll
                        for (int k = 0; k < K; k++) m[k] = Tag<REAL>::put(m[k], tag);
                        REAL *dst = p.msg + term * LP;
                        if (p.world > 1) {
                            st_words<REAL, K, true>(dst, m, lane);
                            if (peer >= 0) st_words<REAL, K, true>(p.peer_msg[peer] + peer_term(p, term, peer) * LP, m, lane);
                        } else {
                            st_words<REAL, K, false>(dst, m, lane);
                        }
                        tick(5);
                    }
                } else {
                    if (lane == 0) mbar_arrive(bar_free + st);
                }
                if (prof_on) tp[7]++;
            }
            if (prof_on && lane == 0) {
                tick(6);
                const int grp = ring_strip ? 0 : (fs == p.head_strip ? 2 : 1);
                for (int q = 0; q < 8; q++) atomicAdd((unsigned long long *)p.prof + grp * 24 + q, (unsigned long long)tp[q]);
            }
        } else {
            // ============================================================ helper warp
            // helper `hid` prepares the steps whose (CTA-lifetime) step count has parity hid: two helpers, so that
            // a helper has two step times for its serial work (polls, sums, hand-over, copy issue)
            const int hid = warp - NTW;
            const int first = (int)((gstep0 ^ (unsigned)hid) & 1u);   // first own step of this strip
            SegWalker wk, pf;
            wk.init(p.segs, sg0);
            pf.init(p.segs, sg0);
            int wk_pos = 0, pf_pos = 0;
            auto seek = [&](SegWalker &w_, int &pos, int target) {
                while (pos < target && pos + 1 < n_steps) { w_.advance(); pos++; }
            };
            int pf_node = first;   // next own step whose bulk copies are to be issued
            // helper phase timers: wait for a free stage, wait stage, static sum, message polls, rounding polls,
            // write + arrive, gate, issue work
            const bool prof_on = DIAG && p.prof != nullptr && hid == 0;
            long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tsteps = 0, tretry = 0;
            long long tclk = prof_on ? clock64() : 0;
            auto tick = [&](int which) {
                if (prof_on) {
                    const long long now = clock64();
                    tp[which] += now - tclk;
                    tclk = now;
                }
            };

            // Bulk-copy descriptors.  Within a segment every source address is affine in the step index, so each
            // lane works its copy out ONCE per segment (lane 0: node rows, 1: rank row, 2..7: send slot (lane-2)/3
            // {messages, byte rows, neighbour rows}, 8..11: rows to pull towards the L2 for the polls) and a step costs
            // one multiply-add, one expect_tx and one copy instruction.
            int d_seg = -1;                 // segment the descriptors below belong to
            const char *d_src = nullptr;    // source of the segment's first step
            long long d_stride = 0;         // ... advancing by this per step
            unsigned d_dst = 0, d_bytes = 0, d_total = 0;
            auto describe = [&]() {
                StepGeo g;
                pf.get(g);
                const long long u0 = pf.u0, du = pf.du;
                d_src = nullptr; d_stride = 0; d_dst = 0; d_bytes = 0;
                auto node_row = [&](long long uu) { return reinterpret_cast<const char *>(p.nodeF + uu * 4 * LP); };
                if (lane == 0) {
                    d_src = node_row(u0); d_stride = du * 4 * LP * (long long)sizeof(REAL); d_dst = SL::OFF_NF; d_bytes = 4 * SL::ROW;
                } else if (lane == 1) {
                    d_src = reinterpret_cast<const char *>(p.nodeB + u0 * LP); d_stride = du * LP; d_dst = SL::OFF_NB; d_bytes = LP;
                } else if (lane < 8) {
                    const int sl = (lane - 2) / 3, what = (lane - 2) % 3;
                    const int d = dir_with_role(g.roles, ROLE_SEND0 + sl);
                    if (d >= 0) {
                        const long long pair0 = pair_of(u0, d, W);     // pair ids advance by 2 du per step
                        if (what == 0) {
                            d_src = reinterpret_cast<const char *>(p.msg + pair0 * 2 * LP); d_stride = 2 * du * 2 * LP * (long long)sizeof(REAL);
                            d_dst = SL::OFF_MS + sl * 2 * SL::ROW; d_bytes = 2 * SL::ROW;
                        } else if (what == 1) {
                            d_src = reinterpret_cast<const char *>(p.pairB + (pair0 * 2 + side_of(d)) * 3 * LP); d_stride = 2 * du * 2 * 3 * LP;
                            d_dst = SL::OFF_PB + sl * 3 * LP; d_bytes = 3 * LP;
                        } else if (d != g.next_dir) {
                            d_src = node_row(nb_of(u0, d, W)) + (vertical(d) ? NF_OWN : NF_GX) * SL::ROW;
                            d_stride = du * 4 * LP * (long long)sizeof(REAL);
                            d_dst = SL::OFF_XN + sl * 2 * SL::ROW; d_bytes = 2 * SL::ROW;
                        }
                    }
                } else if (lane < 12 && do_send && !(g.flags & GF_DEFERRED)) {
                    const int d = lane - 8;
                    if (role_of(g.roles, d) == ROLE_POLL) {
                        d_src = reinterpret_cast<const char *>(p.msg + pair_of(u0, d, W) * 2 * LP);
                        d_stride = 2 * du * 2 * LP * (long long)sizeof(REAL);
                        d_bytes = 2 * SL::ROW;   // prefetch only (no stage slot)
                    }
                }
                unsigned total = lane < 8 ? d_bytes : 0u;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                d_total = __shfl_sync(0xffffffffu, total, 0);
                d_seg = pf.sg;
            };
            // (stage, phase) of the step the next issue() refills and of the step the main loop prepares: both advance by 2
            int ist = (int)((gstep0 + (unsigned)first) % NS);
            unsigned iph = (unsigned)(((gstep0 + (unsigned)first) / NS) & 1);
            int hst = ist;
            unsigned hph = iph;
            auto issue = [&](int node) {
                const int st = ist;
                // the stage was last used NS steps ago: the term warps must have taken their operands from it
                record(fs, node, 13, st);
                tick(5);
                // (first use of a stage: the barrier is in its initial phase, and waiting for the parity of the phase
                // before it returns at once)
                mbar_wait(bar_free + st, iph ^ 1u);
                ist += 2;
                if (ist >= NS) { ist -= NS; iph ^= 1u; }
                record(fs, node, 15, st);
                tick(0);
                seek(pf, pf_pos, node);
                if (pf.sg != d_seg) describe();
                const char *src = d_src + (long long)pf.i * d_stride;
                if (lane == 0) mbar_expect_tx(bar_full + st, d_total);
                __syncwarp();
                if (d_bytes) {
                    if (lane < 8) bulk_g2s(stage_ptr(st) + d_dst, src, d_bytes, bar_full + st);
                    else asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(d_bytes) : "memory");
                }
            };
            if (pf_node < n_steps) { issue(pf_node); pf_node += 2; }    // (PD = 2: one own step ahead)
            int pre_d = -1;
            unsigned long long pre_wv = 0;
            REAL pre_v0[K], pre_v1[K];
#pragma unroll
            for (int k = 0; k < K; k++) { pre_v0[k] = REAL(0); pre_v1[k] = REAL(0); }
            for (int node = first; node < n_steps; node += 2) {
                const unsigned gs = gstep0 + (unsigned)node;
                const int st = hst;
                const unsigned ph = hph;
                hst += 2;
                if (hst >= NS) { hst -= NS; hph ^= 1u; }
                const int par = (int)(gs & 1u);     // == hid
                StepGeo g;
                seek(wk, wk_pos, node);
                wk.get(g);
                record(fs, node, 11, st);
                mbar_wait(bar_full + st, ph);
                record(fs, node, 16, (int)g.roles);
                // Set `par` and its hand-over barrier were last used by step node - 2: the term warps have all passed
                // both once they have taken their operands of that step.
                if (node >= 2) {
                    // step gs - 2
                    record(fs, node, 14, 0);
                    mbar_wait(bar_free + (st >= 2 ? st - 2 : st + NS - 2), st >= 2 ? ph : ph ^ 1u);
                }
                tick(1);
                tick(6);
                if (g.flags & GF_DEFERRED) {
                    REAL base[K];
                    poll_total<REAL, K>(p.save + (size_t)g.save * LP * (sizeof(REAL) / 4), base, lane, p.epoch);
                    lrow_sts<REAL, K>(set_ptr(par, R_BASE), base, lane);
                } else {
                    const unsigned char *sp = stage_ptr(st);
                    const REAL *NF = reinterpret_cast<const REAL *>(sp + SL::OFF_NF);
                    REAL base[K], dib[K], rms[K];
                    lrow_lds<REAL, K>(base, NF + NF_D * LP, lane);
#pragma unroll
                    for (int k = 0; k < K; k++) { dib[k] = base[k]; rms[k] = REAL(0); }
                    // old messages of the send pairs (they are incoming messages of this node)
#pragma unroll
                    for (int sl = 0; sl < 2; sl++) {
                        if (dir_with_role(g.roles, ROLE_SEND0 + sl) < 0) continue;
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            REAL v[K];
                            lrow_lds<REAL, K>(v, reinterpret_cast<const REAL *>(sp + SL::OFF_MS) + (size_t)(sl * 2 + jj) * LP, lane);
#pragma unroll
                            for (int k = 0; k < K; k++) { const REAL a = Tag<REAL>::val(v[k]); base[k] += a; rms[k] += a; }
                        }
                    }
                    if (prof_on) { tclk += (long long)(base[0] != base[0]); tick(2); }
                    for (int d = 0; d < 4; d++) {
                        const int role = role_of(g.roles, d);
                        if (role != ROLE_ADD && role != ROLE_POLL) continue;
                        const long long pair = pair_of(g.u, d, W);
                        const REAL *mrow = p.msg + pair * 2 * LP;
                        if (role == ROLE_ADD) {
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                REAL v[K];
                                ld_words<REAL, K, false>(v, mrow + jj * LP, lane);
#pragma unroll
                                for (int k = 0; k < K; k++) { const REAL a = Tag<REAL>::val(v[k]); base[k] += a; rms[k] += a; }
                            }
                            continue;
                        }
                        // the neighbour's two messages of THIS pass (every word carries the pass tag once written) and,
                        // when the pass rounds, its rounded label: all polls of a round are in flight together
                        const bool need_sel = do_round && lane < 2;
                        bool have_sel = !need_sel;
                        unsigned long long wv = 0;
                        const unsigned long long *sb = p.selbox + pair * 2 + (lane & 1);
                        REAL v0[K], v1[K];
#pragma unroll
                        for (int k = 0; k < K; k++) { v0[k] = REAL(0); v1[k] = REAL(0); }
                        // the first round of polls of this pair was issued a step ago (behind the previous hand-over,
                        // in flight while the bulk copies were issued): use those words first
                        bool preloaded = pre_d == d;
                        if (preloaded) {
                            wv = pre_wv;
#pragma unroll
                            for (int k = 0; k < K; k++) { v0[k] = pre_v0[k]; v1[k] = pre_v1[k]; }
                        }
                        record(fs, node, 12, d);
                        for (;;) {
                            if (!have_sel && !preloaded) wv = p.world > 1 ? trws::ld_mbox_sys(sb) : trws::ld_mbox(sb);
                            bool ok = true;
                            if (do_send) {
                                if (!preloaded) {
                                    if (p.world > 1) { ld_words<REAL, K, true>(v0, mrow, lane); ld_words<REAL, K, true>(v1, mrow + LP, lane); }
                                    else { ld_words<REAL, K, false>(v0, mrow, lane); ld_words<REAL, K, false>(v1, mrow + LP, lane); }
                                }
#pragma unroll
                                for (int k = 0; k < K; k++) ok = ok && Tag<REAL>::ok(v0[k], tag) && Tag<REAL>::ok(v1[k], tag);
                            }
                            if (!have_sel) {
                                if ((unsigned)(wv >> 32) == p.epoch) have_sel = true; else ok = false;
                            }
                            preloaded = false;
                            if (__all_sync(0xffffffffu, ok)) break;
                            tretry++;
                            __nanosleep(p.poll_ns);
                        }
                        if (do_send) {
#pragma unroll
                            for (int k = 0; k < K; k++) base[k] += Tag<REAL>::val(v0[k]) + Tag<REAL>::val(v1[k]);
                        }
                        tick(3);
                        if (do_round) {
                            REAL sel = REAL(0), al = REAL(0);
                            const int sd = side_of(d);
                            if (lane < 2) {
                                if constexpr (sizeof(REAL) == 4) {
                                    sel = (REAL)__uint_as_float((unsigned)wv);
                                } else {
                                    // position of the neighbour's rounded label on term `lane`: its own disparity if it
                                    // is the head of the term, else its plane evaluated at my point
                                    const int xl = (int)(unsigned)wv;
                                    const REAL *nrec = p.nodeF + nb_of(g.u, d, W) * 4 * LP;
                                    sel = __ldg(nrec + NF_OWN * LP + xl);
                                    if (lane != sd) {   // I am the head of term `lane`: the neighbour is its tail
                                        const REAL gn = __ldg(nrec + (vertical(d) ? NF_GY : NF_GX) * LP + xl);
                                        sel = sd == 1 ? sel + gn : sel - gn;
                                    }
                                }
                                al = __ldg(p.alpha + pair * 2 + lane);
                            }
                            __syncwarp();
                            REAL own_me[K], g_me[K];
                            lrow_lds<REAL, K>(own_me, NF + NF_OWN * LP, lane);
                            lrow_lds<REAL, K>(g_me, NF + (vertical(d) ? NF_GY : NF_GX) * LP, lane);
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                const REAL aj = __shfl_sync(0xffffffffu, al, jj), sj = __shfl_sync(0xffffffffu, sel, jj);
                                const bool tail = (jj == sd);   // am I the tail of term jj
#pragma unroll
                                for (int k = 0; k < K; k++) {
                                    const REAL pos = tail ? (sd == 0 ? own_me[k] + g_me[k] : own_me[k] - g_me[k]) : own_me[k];
                                    dib[k] += aj * trws::smooth<REAL, KERN>(pos - sj, p.lambda);
                                }
                            }
                            if (prof_on) { tclk += (long long)(dib[0] != dib[0]); tick(4); }
                        }
                    }
                    lrow_sts<REAL, K>(set_ptr(par, R_BASE), base, lane);
                    if (do_round) {
                        lrow_sts<REAL, K>(set_ptr(par, R_DIB0), dib, lane);
                        lrow_sts<REAL, K>(set_ptr(par, R_RMS), rms, lane);
                    }
                }
                record(fs, node, 17, 0);
                // hand-over: an mbarrier arrive (a named-barrier arrive also waits for the warp's outstanding copies)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_base + par);
                tick(5);
                // bulk copies of the step PD ahead: behind the hand-over, so that the helper runs up to two steps
                // ahead of the term warps (the stage it refills was released when they STARTED step node - 1)
                // first round of polls of the NEXT step: in flight while the bulk copies below are issued
                pre_d = -1;
                if (node + 2 < n_steps) {
                    StepGeo gn;
                    seek(wk, wk_pos, node + 2);
                    wk.get(gn);
                    if (!(gn.flags & GF_DEFERRED)) {
                        const int dn = dir_with_role(gn.roles, ROLE_POLL);
                        if (dn >= 0) {
                            const long long pairn = pair_of(gn.u, dn, W);
                            const REAL *mrow = p.msg + pairn * 2 * LP;
                            if (do_send) {
                                if (p.world > 1) { ld_words<REAL, K, true>(pre_v0, mrow, lane); ld_words<REAL, K, true>(pre_v1, mrow + LP, lane); }
                                else { ld_words<REAL, K, false>(pre_v0, mrow, lane); ld_words<REAL, K, false>(pre_v1, mrow + LP, lane); }
                            }
                            if (do_round && lane < 2) {
                                const unsigned long long *sb = p.selbox + pairn * 2 + (lane & 1);
                                pre_wv = p.world > 1 ? trws::ld_mbox_sys(sb) : trws::ld_mbox(sb);
                            }
                            pre_d = dn;
                        }
                    }
                }
                if (pf_node < n_steps) { issue(pf_node); pf_node += 2; }
                tick(7);
                tsteps++;
            }
            if (prof_on && lane == 0) {
                const int grp = ring_strip ? 0 : (fs == p.head_strip ? 2 : 1);
                for (int q = 0; q < 8; q++) atomicAdd((unsigned long long *)p.prof + grp * 24 + 8 + q, (unsigned long long)tp[q]);
                atomicAdd((unsigned long long *)p.prof + grp * 24 + 16, (unsigned long long)tsteps);
                atomicAdd((unsigned long long *)p.prof + grp * 24 + 17, (unsigned long long)tretry);
            }
        }
        gstep0 += (unsigned)n_steps;
        record(fs, n_steps, 99, 0);
    }
    if (is_term && lane == 0) {
        if (acc_energy != 0.0) atomicAdd(p.acc + 0, acc_energy);
        if (acc_lb != 0.0) atomicAdd(p.acc + 1, acc_lb);
    }
}

} // namespace gtrws
} // namespace sb
