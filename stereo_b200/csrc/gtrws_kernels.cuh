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
//   * node ids are row-major and band-local, so a row band (multi-GPU) is a contiguous slab;
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
// CTA = 4 term warps + 1 helper warp, one strip (the boundary ring, then one per interior row) at a
// time from an atomic ticket, one persistent launch per pass.
#pragma once
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

constexpr int NTW = 4;                       // term warps
constexpr int CTA_THREADS = (NTW + 1) * 32;  // + helper warp
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
    const int32_t *seg_ptr, *strip_len;
    int S;
    int world;
    REAL *peer_msg[2];                 // rank - 1 / rank + 1 (peer mapped)
    unsigned long long *peer_selbox[2];
    long long peer_dn[2];              // their local node id of my local node u is u + peer_dn
    int32_t *sol;                      // [Nloc]
    unsigned epoch;                    // launch counter: tag of the selbox words
    unsigned tag;                      // 0 / 1: sign the messages of this pass carry
    int *ticket;
    double *acc;                       // [0] energy [1] lower bound
    int mode;
    long long *prof;                   // optional cycle counters
    int prof_warp;                     // term warp they are taken on
    int gate;                          // > 1: every `gate` steps a strip waits until its upstream strip is `gate` nodes ahead
    int *rec;                          // SB_TRWS_RECORD: host-mapped flight recorder [cta][5 warps][4], else null
};

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

// K consecutive message words of this lane, L2-coherent relaxed loads (the words are written by
// other SMs / GPUs during this launch).  Each word validates itself, so per-word atomicity suffices.
template <typename REAL, int K, bool SYS> __device__ __forceinline__ void ld_words(REAL (&r)[K], const REAL *p)
{
    if constexpr (sizeof(REAL) == 4) {
        if constexpr (K % 4 == 0) {
#pragma unroll
            for (int i = 0; i < K / 4; i++) {
                unsigned a, b, c, d;
                if constexpr (SYS) asm volatile("ld.relaxed.sys.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p + 4 * i) : "memory");
                else asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p + 4 * i) : "memory");
                r[4 * i] = __uint_as_float(a); r[4 * i + 1] = __uint_as_float(b); r[4 * i + 2] = __uint_as_float(c); r[4 * i + 3] = __uint_as_float(d);
            }
        } else if constexpr (K % 2 == 0) {
#pragma unroll
            for (int i = 0; i < K / 2; i++) {
                unsigned a, b;
                if constexpr (SYS) asm volatile("ld.relaxed.sys.global.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p + 2 * i) : "memory");
                else asm volatile("ld.relaxed.gpu.global.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p + 2 * i) : "memory");
                r[2 * i] = __uint_as_float(a); r[2 * i + 1] = __uint_as_float(b);
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) {
                unsigned a;
                if constexpr (SYS) asm volatile("ld.relaxed.sys.global.b32 %0, [%1];" : "=r"(a) : "l"(p + i) : "memory");
                else asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(a) : "l"(p + i) : "memory");
                r[i] = __uint_as_float(a);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; i++) {
            unsigned long long a;
            if constexpr (SYS) asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(a) : "l"(p + i) : "memory");
            else asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(a) : "l"(p + i) : "memory");
            r[i] = __longlong_as_double((long long)a);
        }
    }
}
// K consecutive message words of this lane to global memory (relaxed, L2; SYS: a peer GPU's memory)
template <typename REAL, int K, bool SYS> __device__ __forceinline__ void st_words(REAL *p, const REAL (&r)[K])
{
    if constexpr (sizeof(REAL) == 4) {
        if constexpr (K % 4 == 0) {
#pragma unroll
            for (int i = 0; i < K / 4; i++) {
                const unsigned a = __float_as_uint(r[4 * i]), b = __float_as_uint(r[4 * i + 1]), c = __float_as_uint(r[4 * i + 2]), d = __float_as_uint(r[4 * i + 3]);
                if constexpr (SYS) asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * i), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
                else asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p + 4 * i), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
            }
        } else if constexpr (K % 2 == 0) {
#pragma unroll
            for (int i = 0; i < K / 2; i++) {
                const unsigned a = __float_as_uint(r[2 * i]), b = __float_as_uint(r[2 * i + 1]);
                if constexpr (SYS) asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1,%2};" ::"l"(p + 2 * i), "r"(a), "r"(b) : "memory");
                else asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1,%2};" ::"l"(p + 2 * i), "r"(a), "r"(b) : "memory");
            }
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) {
                const unsigned a = __float_as_uint(r[i]);
                if constexpr (SYS) asm volatile("st.relaxed.sys.global.b32 [%0], %1;" ::"l"(p + i), "r"(a) : "memory");
                else asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p + i), "r"(a) : "memory");
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < K; i++) {
            const unsigned long long a = (unsigned long long)__double_as_longlong(r[i]);
            if constexpr (SYS) asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p + i), "l"(a) : "memory");
            else asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p + i), "l"(a) : "memory");
        }
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

// ---------------------------------------------------------------- geometry of a node step
struct StepGeo {
    int u;
    unsigned roles;
    int gamma_den, next_dir, flags;
    unsigned peer;   // 8 bits per direction
};
struct SegWalker {
    const GSeg *segs;
    int sg, i, n, u0, du;
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
        i = 0;
    }
    __device__ __forceinline__ void init(const GSeg *s, int sg0) { segs = s; sg = sg0; load(); }
    __device__ __forceinline__ void get(StepGeo &g) const
    {
        g.u = u0 + i * du; g.roles = roles; g.gamma_den = gamma_den; g.next_dir = next_dir; g.flags = flags; g.peer = peer;
    }
    // to the next step (the caller guarantees there is one)
    __device__ __forceinline__ void advance()
    {
        if (++i >= n) { sg++; load(); }
    }
};
__device__ __forceinline__ int role_of(unsigned roles, int d) { return (roles >> (4 * d)) & 15; }
__device__ __forceinline__ int dir_with_role(unsigned roles, int role)
{
#pragma unroll
    for (int d = 0; d < 4; d++)
        if (role_of(roles, d) == role) return d;
    return -1;
}
__device__ __forceinline__ long long nb_of(long long u, int d, int W) { return d == DIR_UP ? u - W : d == DIR_DOWN ? u + W : d == DIR_LEFT ? u - 1 : u + 1; }
// pair record of the neighbour pair in direction d: pairs are owned by their upper / left node
__device__ __forceinline__ long long pair_of(long long u, int d, int W)
{
    return d == DIR_DOWN ? 2 * u : d == DIR_RIGHT ? 2 * u + 1 : d == DIR_UP ? 2 * (u - W) : 2 * (u - 1) + 1;
}
// 0: I am the pair's first (upper / left) node, 1: its second.  Term j of a pair has tail = node j.
__device__ __forceinline__ int side_of(int d) { return (d == DIR_DOWN || d == DIR_RIGHT) ? 0 : 1; }
__device__ __forceinline__ bool vertical(int d) { return d == DIR_UP || d == DIR_DOWN; }

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
           (size_t)NTW * trws::scratch_pairs<K>() * sizeof(Pair<REAL>) + 2 * NS * 8;
}

template <int ID> __device__ __forceinline__ void full_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(CTA_THREADS) : "memory"); }
template <int ID> __device__ __forceinline__ void full_arrive() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(CTA_THREADS) : "memory"); }

// resident CTAs per SM the register budget is set for
template <typename REAL, int K> __host__ __device__ constexpr int gsweep_min_blocks()
{
    if (sizeof(REAL) == 8) return K <= 2 ? 3 : K <= 4 ? 2 : 1;
    return K <= 2 ? 6 : K <= 3 ? 5 : K <= 4 ? 4 : K <= 6 ? 4 : 3;
}

template <typename REAL, int K, int KERN, int PASS, int NS>
__global__ void __launch_bounds__(CTA_THREADS, (gsweep_min_blocks<REAL, K>())) gsweep_kernel(const GProblem<REAL> p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_ticket;
    typedef StageLayout<REAL, K> SL;
    constexpr int LP = 32 * K;
    constexpr int PD = NS - 1;   // bulk copies are issued PD steps ahead
    const int lane = threadIdx.x & 31;
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
    unsigned long long *bar_full = bars, *bar_free = bars + NS;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_free + s, NTW);
        }
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
        if (p.rec && lane == 0) {
            volatile int *r = p.rec + ((size_t)blockIdx.x * (NTW + 1) + warp) * 4;
            r[0] = strip; r[1] = step; r[2] = phase; r[3] = extra;
        }
    };
    long long gstep0 = 0;   // node steps this CTA has walked before the current strip (stage ring position)

    for (;;) {
        if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1);
        __syncthreads();
        const int ts = s_ticket;
        __syncthreads();
        if (ts >= p.S) break;
        const int fs = (PASS == PASS_BWD) ? p.S - 1 - ts : ts;
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
            // optional phase timers (SB_TRWS_PROFILE): term warp 0 -> wait FULL, wait stage, node total + rounding,
            // operands, update, stores
            const bool prof_on = (p.prof != nullptr) && w == p.prof_warp;
            long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tclk = prof_on ? clock64() : 0;
            auto tick = [&](int which) {
                if (prof_on) {
                    const long long now = clock64();
                    tp[which] += now - tclk;
                    tclk = now;
                }
            };
            for (int node = 0; node < n_steps; node++) {
                const long long gs = gstep0 + node;
                const int st = (int)(gs % NS);
                const unsigned ph = (unsigned)((gs / NS) & 1);
                const int par = node & 1;
                StepGeo g;
                wk.get(g);
                if (node + 1 < n_steps) wk.advance();
                const REAL gamma = REAL(1) / REAL(g.gamma_den);
                const unsigned char *sp = stage_ptr(st);
                const REAL *NF = reinterpret_cast<const REAL *>(sp + SL::OFF_NF);
                // ---- rows of this node are ready (helper), every term warp has finished the previous step
                tick(6);
                record(fs, node, 1, (int)g.roles);
                if (par) full_sync<2>(); else full_sync<1>();
                tick(0);
                record(fs, node, 2, st);
                mbar_wait(bar_full + st, ph);   // the bulk copies of this stage, as seen by THIS thread
                record(fs, node, 3, st);
                tick(1);
                REAL Di[K];
                if (g.flags & GF_SECOND) {
                    trws::row_lds<REAL, K>(Di, di_save, lane);
                } else {
                    trws::row_lds<REAL, K>(Di, set_ptr(par, R_BASE), lane);
                    const int cd = dir_with_role(g.roles, ROLE_CARRY);
                    if (cd >= 0 && do_send) {
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            REAL v[K];
                            trws::row_lds<REAL, K>(v, carry_ptr(par, jj), lane);
#pragma unroll
                            for (int k = 0; k < K; k++) Di[k] += v[k];
                        }
                    }
                    if (do_round) {
                        // minimize.cpp:240-260: DiB = D + sum_{lower nb} V(x_nb, .), Dr = DiB + forward messages
                        REAL dib[K], rms[K];
                        trws::row_lds<REAL, K>(dib, set_ptr(par, R_DIB0), lane);
                        trws::row_lds<REAL, K>(rms, set_ptr(par, R_RMS), lane);
                        if (cd >= 0) {
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                REAL v[K];
                                trws::row_lds<REAL, K>(v, carry_ptr(par, 2 + jj), lane);
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
                        const REAL wbest = trws::warp_min(best);
                        bi = trws::warp_min_s32(best == wbest ? bi : 0x7fffffff);
                        xs = bi;
                        if (w == 0) {
                            REAL dv = dib[0];
#pragma unroll
                            for (int k = 1; k < K; k++)
                                if (k == bi % K) dv = dib[k];
                            dv = __shfl_sync(0xffffffffu, dv, bi / K);
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
                            if (lane * K + k < p.L) vmin = min(vmin, Di[k]);
                        vmin = trws::warp_min(vmin);
#pragma unroll
                        for (int k = 0; k < K; k++) Di[k] -= vmin;
                        if (w == 0) acc_lb += (double)vmin;
                    }
                    if ((g.flags & GF_FIRST) && w == 0) trws::row_sts<REAL, K>(di_save, Di, lane);
                }
                if (prof_on) { tclk += (long long)(Di[0] != Di[0]); tick(2); }
                // ---- my send term: pair in send slot `slot`, term j of it
                const int d = dir_with_role(g.roles, ROLE_SEND0 + slot);
                if (d >= 0) {
                    const bool vert = vertical(d);
                    const int sd = side_of(d);
                    const bool tail = (j == sd);
                    const bool to_next = (d == g.next_dir);
                    const long long pair = pair_of(g.u, d, W);
                    const long long term = 2 * pair + j;
                    const int peer = (int)((g.peer >> (8 * d)) & 255u) - 1;   // -1 local, 0 rank - 1, 1 rank + 1
                    const REAL alpha = __ldg(p.alpha + term);
                    // operands from the stage ring
                    REAL m[K], s[K], x[K];
                    uint8_t rk[K], cn[K];
                    {
                        const REAL *MS = reinterpret_cast<const REAL *>(sp + SL::OFF_MS) + (size_t)(slot * 2 + j) * LP;
                        trws::row_lds<REAL, K>(m, MS, lane);
#pragma unroll
                        for (int k = 0; k < K; k++) m[k] = Tag<REAL>::val(m[k]);
                        REAL own_me[K], g_me[K], own_nb[K], g_nb[K];
                        trws::row_lds<REAL, K>(own_me, NF + NF_OWN * LP, lane);
                        const REAL *nb_own, *nb_g;
                        if (to_next) {
                            // the receiver is the next node of the strip: its rows are (or will shortly be) in the next stage
                            const long long gn = gs + 1;
                            const int stn = (int)(gn % NS);
                            record(fs, node, 4, stn);
                            mbar_wait(bar_full + stn, (unsigned)((gn / NS) & 1));
                            record(fs, node, 5, stn);
                            const REAL *NFn = reinterpret_cast<const REAL *>(stage_ptr(stn) + SL::OFF_NF);
                            nb_own = NFn + NF_OWN * LP;
                            nb_g = NFn + (vert ? NF_GY : NF_GX) * LP;
                        } else {
                            const REAL *XN = reinterpret_cast<const REAL *>(sp + SL::OFF_XN) + (size_t)(slot * 2) * LP;
                            nb_own = XN + (vert ? 0 : 1) * LP;
                            nb_g = XN + (vert ? 1 : 0) * LP;
                        }
                        trws::row_lds<REAL, K>(own_nb, nb_own, lane);
                        trws::PackedBytes<K> rkp, cnp;
                        const uint8_t *PB = sp + SL::OFF_PB + (size_t)slot * 3 * LP;
                        if (tail) {
                            // my positions: my planes at the receiver's point (qprim); receiver's: its own (q)
                            trws::row_lds<REAL, K>(g_me, NF + (vert ? NF_GY : NF_GX) * LP, lane);
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                s[k] = sd == 0 ? own_me[k] + g_me[k] : own_me[k] - g_me[k];
                                x[k] = own_nb[k];
                            }
                            rkp.load_shared(PB + 2 * LP + lane * K);
                            cnp.load_shared(PB + 0 * LP + lane * K);
                        } else {
                            // I am the head: my own disparities (q); receiver's planes at my point (qprim)
                            trws::row_lds<REAL, K>(g_nb, nb_g, lane);
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                s[k] = own_me[k];
                                x[k] = sd == 0 ? own_nb[k] - g_nb[k] : own_nb[k] + g_nb[k];
                            }
                            rkp.load_shared(sp + SL::OFF_NB + lane * K);
                            cnp.load_shared(PB + 1 * LP + lane * K);
                        }
                        rkp.unpack(rk);
                        cnp.unpack(cn);
                    }
                    // the stage of this step is no longer needed by this warp
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_free + st);
                    if (prof_on) { tclk += (long long)(x[0] != x[0]) + (long long)(m[0] != m[0]); tick(3); }
                    if (do_round) {
                        // position of the rounded label on this term, for the receiver's rounding
                        REAL sv = s[0];
#pragma unroll
                        for (int k = 1; k < K; k++)
                            if (k == xs % K) sv = s[k];
                        sv = __shfl_sync(0xffffffffu, sv, xs / K);
                        if (lane == 0 && !to_next) {
                            // fp32: the word carries the position itself; fp64 (64 value bits do not fit beside the
                            // tag): the label, and the receiver recomputes the position from my plane rows
                            unsigned lo;
                            if constexpr (sizeof(REAL) == 4) lo = __float_as_uint((float)sv); else lo = (unsigned)xs;
                            const unsigned long long wv = (unsigned long long)lo | ((unsigned long long)p.epoch << 32);
                            if (peer >= 0) trws::st_mbox_sys(p.peer_selbox[peer] + (term + 4 * p.peer_dn[peer]), wv);
                            else trws::st_mbox(p.selbox + term, wv);
                        }
                        if (to_next) {
                            REAL cc[K];
#pragma unroll
                            for (int k = 0; k < K; k++) cc[k] = alpha * trws::smooth<REAL, KERN>(x[k] - sv, p.lambda);
                            trws::row_sts<REAL, K>(carry_ptr(par ^ 1, 2 + j), cc, lane);
                        }
                    }
                    if (do_send) {
                        REAL vmin;
                        if constexpr (KERN == 1)
                            vmin = trws::update_linear<REAL, K>(gamma, alpha, p.lambda, p.L, lane, Di, m, s, rk, x, cn, P);
                        else
                            vmin = trws::update_quadratic<REAL, K>(gamma, alpha, p.lambda, p.L, lane, Di, m, s, rk, x, cn, P);
                        if (PASS == PASS_BWD) acc_lb += (double)vmin;
                        if (prof_on) { tclk += (long long)(m[0] != m[0]); tick(4); }
                        if (to_next) trws::row_sts<REAL, K>(carry_ptr(par ^ 1, j), m, lane);
#pragma unroll
                        for (int k = 0; k < K; k++) m[k] = Tag<REAL>::put(m[k], tag);
                        REAL *dst = p.msg + term * LP + lane * K;
                        if (p.world > 1) {
                            st_words<REAL, K, true>(dst, m);
                            if (peer >= 0) st_words<REAL, K, true>(p.peer_msg[peer] + (term + 4 * p.peer_dn[peer]) * LP + lane * K, m);
                        } else {
                            st_words<REAL, K, false>(dst, m);
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
                const int grp = (fs == 0) ? 0 : 1;
                for (int q = 0; q < 8; q++) atomicAdd((unsigned long long *)p.prof + grp * 16 + q, (unsigned long long)tp[q]);
            }
        } else {
            // ============================================================ helper warp
            SegWalker wk, pf;
            wk.init(p.segs, sg0);
            pf.init(p.segs, sg0);
            int pf_node = 0;   // next step whose bulk copies are to be issued
            auto issue = [&](int node) {
                const long long gs = gstep0 + node;
                const int st = (int)(gs % NS);
                // the stage was last used NS steps ago: the term warps must have taken their operands from it
                record(fs, node, 13, st);
                if (gs >= NS) mbar_wait(bar_free + st, (unsigned)(((gs / NS) - 1) & 1));
                record(fs, node, 15, st);
                StepGeo g;
                pf.get(g);
                if (node + 1 < n_steps) pf.advance();
                unsigned char *sp = stage_ptr(st);
                // lane 0: node rows, lane 1: rank row, lanes 2..7: send slot (lane-2)/3 {messages, byte rows, neighbour rows}
                const void *src = nullptr;
                void *dst = nullptr;
                unsigned bytes = 0;
                if (lane == 0) {
                    src = p.nodeF + (long long)g.u * 4 * LP; dst = sp + SL::OFF_NF; bytes = 4 * SL::ROW;
                } else if (lane == 1) {
                    src = p.nodeB + (long long)g.u * LP; dst = sp + SL::OFF_NB; bytes = LP;
                } else if (lane < 8) {
                    const int sl = (lane - 2) / 3, what = (lane - 2) % 3;
                    const int d = dir_with_role(g.roles, ROLE_SEND0 + sl);
                    if (d >= 0) {
                        const long long pair = pair_of(g.u, d, W);
                        if (what == 0) {
                            src = p.msg + pair * 2 * LP; dst = sp + SL::OFF_MS + (size_t)sl * 2 * SL::ROW; bytes = 2 * SL::ROW;
                        } else if (what == 1) {
                            src = p.pairB + (pair * 2 + side_of(d)) * 3 * LP; dst = sp + SL::OFF_PB + (size_t)sl * 3 * LP; bytes = 3 * LP;
                        } else if (d != g.next_dir) {
                            const long long nb = nb_of(g.u, d, W);
                            src = p.nodeF + nb * 4 * LP + (vertical(d) ? NF_OWN : NF_GX) * LP;
                            dst = sp + SL::OFF_XN + (size_t)sl * 2 * SL::ROW; bytes = 2 * SL::ROW;
                        }
                    }
                }
                unsigned total = bytes;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                if (lane == 0) mbar_expect_tx(bar_full + st, total);
                __syncwarp();
                if (bytes) bulk_g2s(dst, src, bytes, bar_full + st);
                // rows the helper will poll for that step: pull them towards the L2 (where the sender ran long
                // ago -- the ring in the backward pass -- they have left it)
                if (lane >= 8 && lane < 12 && do_send) {
                    const int d = lane - 8;
                    if (role_of(g.roles, d) == ROLE_POLL) {
                        const char *row = reinterpret_cast<const char *>(p.msg + pair_of(g.u, d, W) * 2 * LP);
                        for (int t = 0; t < 2 * SL::ROW; t += 128) trws::prefetch_l2(row + t);
                    }
                }
            };
            for (; pf_node < PD && pf_node < n_steps; pf_node++) issue(pf_node);
            // helper phase timers: issue (incl. wait for a free stage), wait stage, static sum, message polls,
            // rounding polls, write + arrive
            const bool prof_on = p.prof != nullptr;
            long long tp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            long long tclk = prof_on ? clock64() : 0;
            auto tick = [&](int which) {
                if (prof_on) {
                    const long long now = clock64();
                    tp[which] += now - tclk;
                    tclk = now;
                }
            };

            for (int node = 0; node < n_steps; node++) {
                const long long gs = gstep0 + node;
                const int st = (int)(gs % NS);
                const unsigned ph = (unsigned)((gs / NS) & 1);
                const int par = node & 1;
                StepGeo g;
                wk.get(g);
                const int wk_i = wk.i, wk_n = wk.n, wk_du = wk.du;
                if (node + 1 < n_steps) wk.advance();
                record(fs, node, 11, st);
                mbar_wait(bar_full + st, ph);
                record(fs, node, 16, (int)g.roles);
                // Set `par` and the named barrier of this parity were last used by step node - 2: the term warps
                // have all passed both once they have taken their operands of that step.  issue() at the end of the
                // previous step waited for exactly that; where nothing was issued there (end of a strip), wait here.
                if (node >= 2 && node - 1 + PD >= n_steps) {
                    const long long g2 = gs - 2;
                    record(fs, node, 14, 0);
                    mbar_wait(bar_free + (int)(g2 % NS), (unsigned)((g2 / NS) & 1));
                }
                tick(1);
                // Slack gate.  A strip's step waits for its upstream strip's step, which waits for ITS upstream's ...:
                // with every hand-over taken at the last moment, the jitter of all strips above accumulates (a
                // last-passage lattice).  Every `gate` steps this strip therefore waits until the upstream strip is
                // `gate` nodes ahead, so that the polls of the steps in between succeed at once.
                if (p.gate > 1 && do_send && !g.flags && (wk_i % p.gate) == 0 && wk_i + p.gate - 1 < wk_n) {
                    const long long ua = (long long)g.u + (long long)(p.gate - 1) * wk_du;
                    for (int d = 0; d < 4; d++) {
                        if (role_of(g.roles, d) != ROLE_POLL) continue;
                        const REAL *mrow = p.msg + pair_of(ua, d, W) * 2 * LP + lane * K;
                        for (;;) {
                            REAL a0[1], a1[1];
                            if (p.world > 1) { ld_words<REAL, 1, true>(a0, mrow); ld_words<REAL, 1, true>(a1, mrow + LP); }
                            else { ld_words<REAL, 1, false>(a0, mrow); ld_words<REAL, 1, false>(a1, mrow + LP); }
                            if (__all_sync(0xffffffffu, Tag<REAL>::ok(a0[0], tag) && Tag<REAL>::ok(a1[0], tag))) break;
                            __nanosleep(100);
                        }
                    }
                }
                tick(6);
                if (!(g.flags & GF_SECOND)) {
                    const unsigned char *sp = stage_ptr(st);
                    const REAL *NF = reinterpret_cast<const REAL *>(sp + SL::OFF_NF);
                    REAL base[K], dib[K], rms[K];
                    trws::row_lds<REAL, K>(base, NF + NF_D * LP, lane);
#pragma unroll
                    for (int k = 0; k < K; k++) { dib[k] = base[k]; rms[k] = REAL(0); }
                    // old messages of the send pairs (they are incoming messages of this node)
#pragma unroll
                    for (int sl = 0; sl < 2; sl++) {
                        if (dir_with_role(g.roles, ROLE_SEND0 + sl) < 0) continue;
#pragma unroll
                        for (int jj = 0; jj < 2; jj++) {
                            REAL v[K];
                            trws::row_lds<REAL, K>(v, reinterpret_cast<const REAL *>(sp + SL::OFF_MS) + (size_t)(sl * 2 + jj) * LP, lane);
#pragma unroll
                            for (int k = 0; k < K; k++) { const REAL a = Tag<REAL>::val(v[k]); base[k] += a; rms[k] += a; }
                        }
                    }
                    if (prof_on) { tclk += (long long)(base[0] != base[0]); tick(2); }
                    for (int d = 0; d < 4; d++) {
                        const int role = role_of(g.roles, d);
                        if (role != ROLE_ADD && role != ROLE_POLL) continue;
                        const long long pair = pair_of(g.u, d, W);
                        const REAL *mrow = p.msg + pair * 2 * LP + lane * K;
                        if (role == ROLE_ADD) {
#pragma unroll
                            for (int jj = 0; jj < 2; jj++) {
                                REAL v[K];
                                ld_words<REAL, K, false>(v, mrow + jj * LP);
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
                        record(fs, node, 12, d);
                        for (;;) {
                            if (!have_sel) wv = p.world > 1 ? trws::ld_mbox_sys(sb) : trws::ld_mbox(sb);
                            bool ok = true;
                            if (do_send) {
                                if (p.world > 1) { ld_words<REAL, K, true>(v0, mrow); ld_words<REAL, K, true>(v1, mrow + LP); }
                                else { ld_words<REAL, K, false>(v0, mrow); ld_words<REAL, K, false>(v1, mrow + LP); }
#pragma unroll
                                for (int k = 0; k < K; k++) ok = ok && Tag<REAL>::ok(v0[k], tag) && Tag<REAL>::ok(v1[k], tag);
                            }
                            if (!have_sel) {
                                if ((unsigned)(wv >> 32) == p.epoch) have_sel = true; else ok = false;
                            }
                            if (__all_sync(0xffffffffu, ok)) break;
                            __nanosleep(20);
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
                            trws::row_lds<REAL, K>(own_me, NF + NF_OWN * LP, lane);
                            trws::row_lds<REAL, K>(g_me, NF + (vertical(d) ? NF_GY : NF_GX) * LP, lane);
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
                    trws::row_sts<REAL, K>(set_ptr(par, R_BASE), base, lane);
                    if (do_round) {
                        trws::row_sts<REAL, K>(set_ptr(par, R_DIB0), dib, lane);
                        trws::row_sts<REAL, K>(set_ptr(par, R_RMS), rms, lane);
                    }
                }
                record(fs, node, 17, 0);
                if (par) full_arrive<2>(); else full_arrive<1>();
                tick(5);
                // bulk copies of the step PD ahead: behind the hand-over, so that the helper runs up to two steps
                // ahead of the term warps (the stage it refills was released when they STARTED step node - 1)
                if (pf_node < n_steps) { issue(pf_node); pf_node++; }
                tick(0);
                if (prof_on) tp[7]++;
            }
            if (prof_on && lane == 0) {
                const int grp = (fs == 0) ? 0 : 1;
                for (int q = 0; q < 8; q++) atomicAdd((unsigned long long *)p.prof + grp * 16 + 8 + q, (unsigned long long)tp[q]);
            }
        }
        gstep0 += n_steps;
        record(fs, n_steps, 99, 0);
    }
    if (is_term && lane == 0) {
        if (acc_energy != 0.0) atomicAdd(p.acc + 0, acc_energy);
        if (acc_lb != 0.0) atomicAdd(p.acc + 1, acc_lb);
    }
}

} // namespace gtrws
} // namespace sb
