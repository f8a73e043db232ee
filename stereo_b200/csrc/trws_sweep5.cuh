// trws_sweep5.cuh -- the fp32 TRW-S sweep kernel, fifth generation ("carry in registers").
//
// Same work as sweep_kernel of trws_kernels.cuh (Minimize_TRW_S forward / backward sweeps,
// cpp/trw-s/minimize.cpp:31-95; UpdateMessage, typeStereoLinear.h:329-487 /
// typeStereoQuadratic.h:329-501; ComputeSolutionAndEnergy, minimize.cpp:223-264), same strip /
// segment schedule (trws_sched.h), same mailbox protocol between strips -- but a different
// division of labour inside the CTA, chosen after the per-phase cycle counters showed that the
// old kernel's step time was set by its helper warps and barriers, not by the update:
//
//   chain warp   owns BOTH messages a node sends to the next node of its strip and keeps them in
//                registers: the dependent chain along a strip is add -> two interleaved min-plus
//                updates -> add, with no barrier, no shared-memory hand-over and no other warp
//                on it.  It also forms the node sum Di and publishes it for the side warps.
//   side warps   (2) own the messages to nodes of other strips: they pick Di up from shared
//                memory, update, and write the message + its self-validating mailbox words.
//   round warp   carries the primal rounding (minimize.cpp:240-260) as its own, much shorter,
//                chain: rounded label of the previous node -> pairwise column -> arg-min.
//   static warp  streams everything that does not depend on this pass (unary row, old messages
//                of the send terms, position rows for the rounding) through a cp.async ring
//                LAND nodes deep and reduces it to BASE = D + sum(old messages).
//   poll warp    polls the mailbox words of the messages arriving from other strips and the
//                selected positions of their rounded labels.
//   prefetch warp pulls the update operands of the nodes ahead into L2.
//
// Hand-over between the warps is by monotone node counters in shared memory (st.release /
// ld.acquire at CTA scope) over a ring of SLOTS node slots; nothing in the CTA ever executes a
// CTA-wide barrier inside a strip.
#pragma once
#include "trws_kernels.cuh"

namespace sb {
namespace trws {
namespace v5 {

// warp w issues on scheduler w % 4: the two chain warps and the two side warps each get a scheduler
// of their own for the updates; round / poll / static / prefetch warps are light or mostly waiting
enum { W_CHAIN = 0, W_CHAIN1 = 1, W_SIDE0 = 2, W_SIDE1 = 3, W_ROUND = 4, W_POLL = 5, W_STATIC = 6, W_PREF = 7, NWARPS = 8 };
constexpr int THREADS = NWARPS * 32;
constexpr int SLOTS = 4;          // node slots between producers and consumers (power of two)
constexpr int LAND = 4;           // depth of the static warp's cp.async ring (power of two)
constexpr int MAX_RND = 6;        // rounding rows per node (lower neighbours in other strips: <= 3 pairs)
constexpr int MAX_STATIC = 1 + 8 + MAX_RND;
enum { SR_BASE = 0, SR_D = 1, SR_DYN = 2, SR_DI = 3, SR_XR = 4, SLOT_ROWS = 4 + MAX_RND };
enum { F_H = 0, F_P = 1, F_C = 2, F_DONE = 3 /* + chain0, side0, side1, round, chain1 */, NDONE = 5, F_X = 8 /* + chain warp */, NFLAGS = 10 };
constexpr int XROWS = 4;          // exchange rows of the two chain warps: [node parity][warp]
constexpr int PF_AHEAD = 8;

template <int K> __host__ __device__ constexpr size_t smem_bytes()
{
    return (size_t)(SLOTS * SLOT_ROWS + XROWS + LAND * MAX_STATIC) * 32 * K * sizeof(float) +
           (size_t)4 * scratch_pairs<K>() * sizeof(Pair<float>);
}

// Node counters in shared memory.  Everything they order is shared memory of the same CTA, which
// the SM's load/store unit processes in issue order: a warp's STS of a row followed (in program
// order, after __syncwarp) by the STS of the counter cannot be seen in the other order by an LDS of
// another warp, and an LDS issued before the counter STS has read its data before the counter
// moves.  So plain volatile accesses suffice -- st.release.cta would put a MEMBAR.ALL.CTA in front
// of every counter store, and that barrier also waits for the warp's outstanding GLOBAL loads and
// stores (measured: ~3.4 k cycles per node on the chain warp).
__device__ __forceinline__ int ld_flag_cta(const int *p)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_flag_cta(int *p, int v)
{
    asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// All lanes poll (one broadcast LDS per round) until the counter reaches `need`.  `seen` caches the
// last value read so that a producer running ahead costs nothing; it is kept WARP-UNIFORM (the exit
// test is a vote, the cached value is lane 0's): lanes that left a polling loop with different
// cached values would later disagree on whether to poll at all, and a __syncwarp / shuffle reached
// by only part of the warp is undefined behaviour (seen on the GPU as garbage shuffle results).
__device__ __forceinline__ void wait_ge(const int *flag, int need, int &seen)
{
    if (seen < need) {
        int v;
        do { v = ld_flag_cta(flag); } while (!__all_sync(0xffffffffu, v >= need));
        seen = __shfl_sync(0xffffffffu, v, 0);
    }
}
// min over the consumers' completion counters (slot recycling), same uniformity rule
__device__ __forceinline__ void wait_free(const int *done4, int need, int &seen)
{
    if (seen < need) {
        int v;
        for (;;) {
            v = min(min(min(ld_flag_cta(done4 + 0), ld_flag_cta(done4 + 1)), min(ld_flag_cta(done4 + 2), ld_flag_cta(done4 + 3))), ld_flag_cta(done4 + 4));
            if (__all_sync(0xffffffffu, v >= need)) break;
            __nanosleep(64);   // slot recycling is never latency critical: leave the issue slots to the working warps
        }
        seen = __shfl_sync(0xffffffffu, v, 0);
    }
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ const void *shfl_ptr(const void *p, int src)
{
    unsigned long long v = (unsigned long long)p;
    unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return (const void *)(((unsigned long long)hi << 32) | lo);
}

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
    __device__ __forceinline__ void unpack(uint8_t (&r)[K]) const
    {
#pragma unroll
        for (int k = 0; k < K; k++) {
            if constexpr (VB == 8) r[k] = (uint8_t)(w[k / 4] >> (8 * (k % 4)));
            else r[k] = (uint8_t)(w[k / VB] >> (8 * (k % VB)));
        }
    }
};

template <int K> struct Own {
    float m[K], s[K], x[K];
    PackedBytes<K> rkp, cnp;
    float alpha;
    long long term;
    int flags;
};

// Two independent truncated-linear updates of the same node (same Di, gamma) with their
// instruction streams interleaved: the shuffle / shared-memory latencies of one hide behind the
// other.  Same arithmetic as update_linear; requires alpha != 0 on both terms.
template <int K>
__device__ __forceinline__ void update_linear2(float gamma, float lambda, int L, int lane, const float (&Di)[K],
                                               Own<K> (&o)[2], Pair<float> *P0, Pair<float> *P1, float (&vout)[2])
{
    const float BIG = Lim<float>::big();
    Pair<float> *P[2] = {P0, P1};
    uint8_t rk[2][K], cn[2][K];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        o[t].rkp.unpack(rk[t]);
        o[t].cnp.unpack(cn[t]);
    }
    float h[2][K], hmin[2], vTrunc[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        hmin[t] = BIG;
#pragma unroll
        for (int k = 0; k < K; k++) {
            h[t][k] = (lane * K + k < L) ? gamma * Di[k] - o[t].m[k] : BIG;
            hmin[t] = min(hmin[t], h[t][k]);
        }
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            Pair<float> q;
            q.a = h[t][k];
            q.b = o[t].s[k];
            P[t][phys<K>(rk[t][k])] = q;
        }
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
        hmin[t] = warp_min(hmin[t]);
        vTrunc[t] = hmin[t] + o[t].alpha * lambda;
    }
    __syncwarp();
    float hs[2][K], ss[2][K];
#pragma unroll
    for (int t = 0; t < 2; t++) {
#pragma unroll
        for (int k = 0; k < K; k++) {
            const Pair<float> q = P[t][phys<K>(lane * K + k)];
            hs[t][k] = q.a;
            ss[t][k] = q.b;
        }
    }
    float gl[2][K], cl[2][K], gr[2][K], cr[2][K];
    float DlL[2], CL[2], DlR[2], CR[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const float al = o[t].alpha;
        float sprev = __shfl_up_sync(0xffffffffu, ss[t][K - 1], 1);
        float snext = __shfl_down_sync(0xffffffffu, ss[t][0], 1);
        if (lane == 0) sprev = ss[t][0];
        if (lane == 31) snext = ss[t][K - 1];
        float g = BIG, cum = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const float d = al * (ss[t][k] - (k ? ss[t][k - 1] : sprev));
            g = min(g + d, hs[t][k]);
            cum += d;
            gl[t][k] = g;
            cl[t][k] = cum;
        }
        DlL[t] = cum;
        CL[t] = g;
        g = BIG;
        cum = 0.f;
#pragma unroll
        for (int k = K - 1; k >= 0; k--) {
            const float d = al * ((k < K - 1 ? ss[t][k + 1] : snext) - ss[t][k]);
            g = min(g + d, hs[t][k]);
            cum += d;
            gr[t][k] = g;
            cr[t][k] = cum;
        }
        DlR[t] = cum;
        CR[t] = g;
    }
#pragma unroll
    for (int ofs = 1; ofs < 32; ofs <<= 1) {
        float DpL[2], CpL[2], DpR[2], CpR[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            DpL[t] = __shfl_up_sync(0xffffffffu, DlL[t], ofs);
            CpL[t] = __shfl_up_sync(0xffffffffu, CL[t], ofs);
            DpR[t] = __shfl_down_sync(0xffffffffu, DlR[t], ofs);
            CpR[t] = __shfl_down_sync(0xffffffffu, CR[t], ofs);
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {
            if (lane >= ofs) {
                CL[t] = min(CpL[t] + DlL[t], CL[t]);
                DlL[t] = DpL[t] + DlL[t];
            }
            if (lane + ofs < 32) {
                CR[t] = min(CpR[t] + DlR[t], CR[t]);
                DlR[t] = DpR[t] + DlR[t];
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
        float VinL = __shfl_up_sync(0xffffffffu, CL[t], 1);
        float VinR = __shfl_down_sync(0xffffffffu, CR[t], 1);
        if (lane == 0) VinL = BIG;
        if (lane == 31) VinR = BIG;
#pragma unroll
        for (int k = 0; k < K; k++) {
            Pair<float> q;
            q.a = min(min(VinL + cl[t][k], gl[t][k]), min(VinR + cr[t][k], gr[t][k]));
            q.b = ss[t][k];
            P[t][phys<K>(lane * K + k)] = q;
        }
    }
    __syncwarp();
    float vmin[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
        const float al = o[t].alpha;
        vmin[t] = BIG;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int c = cn[t][k];
            const Pair<float> lo = P[t][c ? phys<K>(c - 1) : 0];
            const Pair<float> hi = P[t][phys<K>(c)];
            const float xk = o[t].x[k];
            const float v = min(vTrunc[t], min(lo.a + al * fabsf(xk - lo.b), hi.a + al * fabsf(xk - hi.b)));
            o[t].m[k] = v;
            if (lane * K + k < L) vmin[t] = min(vmin[t], v);
        }
    }
#pragma unroll
    for (int t = 0; t < 2; t++) {
        vmin[t] = warp_min(vmin[t]);
#pragma unroll
        for (int k = 0; k < K; k++) o[t].m[k] -= vmin[t];
        vout[t] = vmin[t];
    }
    __syncwarp();
}

template <int K, int KERN>
__device__ __forceinline__ float update_one(float gamma, float lambda, int L, int lane, const float (&Di)[K], Own<K> &o,
                                            Pair<float> *P)
{
    uint8_t rk[K], cn[K];
    o.rkp.unpack(rk);
    o.cnp.unpack(cn);
    if constexpr (KERN == 1)
        return update_linear<float, K>(gamma, o.alpha, lambda, L, lane, Di, o.m, o.s, rk, o.x, cn, P);
    else
        return update_quadratic<float, K>(gamma, o.alpha, lambda, L, lane, Di, o.m, o.s, rk, o.x, cn, P);
}

// walks the nodes of a strip through its segments
struct Cursor {
    int sg, n, i;
    __device__ __forceinline__ void start(const Segment *segs, int sg0)
    {
        sg = sg0;
        n = __ldg(&segs[sg].n);
        i = 0;
    }
    // returns true when the next node starts a new segment
    __device__ __forceinline__ bool advance(const Segment *segs)
    {
        if (++i < n) return false;
        sg++;
        n = __ldg(&segs[sg].n);
        i = 0;
        return true;
    }
};

__device__ __forceinline__ SegOwn ld_own(const Segment *g, int slot)
{
    // slot = w + 4 * half: the order in which build_pass_plan hands the send terms out
    const int4 v = __ldg(reinterpret_cast<const int4 *>(&g->own[slot & 3][slot >> 2]));
    SegOwn so;
    so.term0 = ((long long)(unsigned)v.x) | ((long long)v.y << 32);
    so.tstride = v.z;
    so.flags = v.w;
    return so;
}

template <int K, int KERN, int PASS>
__global__ void __launch_bounds__(THREADS, 1) sweep5_kernel(const Problem<float> p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    __shared__ int s_flag[NFLAGS];
    __shared__ float s_sel[SLOTS][MAX_RND], s_alpha[SLOTS][MAX_RND];
    __shared__ float s_land_alpha[LAND][MAX_RND];
    __shared__ int s_nrnd[SLOTS];
    constexpr int LP = 32 * K;
    constexpr int CH = LP * 4 / 16;       // 16-byte chunks per row
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float *rows = reinterpret_cast<float *>(smem_raw);
    float *xbuf = rows + (size_t)SLOTS * SLOT_ROWS * LP;
    float *land = xbuf + (size_t)XROWS * LP;
    Pair<float> *scratch = reinterpret_cast<Pair<float> *>(land + (size_t)LAND * MAX_STATIC * LP);
    const float BIG = Lim<float>::big();
    const bool do_send = (PASS == PASS_BWD) || (p.mode & MODE_SEND);
    const bool do_round = (PASS == PASS_FWD) && (p.mode & MODE_ROUND);
    const Segment *const segs = p.segs;
    const unsigned ep = p.epoch;
    double acc_energy = 0.0, acc_lb = 0.0;
    auto slot_row = [&](int node, int r) -> float * { return rows + (size_t)((node & (SLOTS - 1)) * SLOT_ROWS + r) * LP; };
    // update scratch: chain 0,1; side0 2; side1 3 -- sentinels written once
    if (warp <= W_SIDE1 && lane == 0) {
        Pair<float> t;
        t.a = BIG;
        t.b = 0.f;
        scratch[(size_t)warp * scratch_pairs<K>()] = t;
        scratch[(size_t)warp * scratch_pairs<K>() + phys<K>(LP)] = t;
    }
    const long long t_begin = p.prof ? clock64() : 0;
    long long prof_c[6] = {0, 0, 0, 0, 0, 0};

    auto load_own = [&](const SegOwn &so, int i, Own<K> &o) {
        o.flags = so.flags;
        o.term = so.term0 + (long long)i * so.tstride;
        const long long off = o.term * LP + lane * K;
        const bool tail = (so.flags & OWN_TAIL) != 0;
        VecIO<float, K>::load_cg(o.m, p.msg + off);
        VecIO<float, K>::load_ro(o.s, (tail ? p.posqp : p.posq) + off);
        VecIO<float, K>::load_ro(o.x, (tail ? p.posq : p.posqp) + off);
        o.rkp.load((tail ? p.rank_qp : p.rank_q) + off);
        o.cnp.load((tail ? p.cnt_q : p.cnt_qp) + off);
        o.alpha = __ldg(p.alpha + o.term);
    };
    auto mbox_word = [&](float v) -> unsigned long long {
        return (unsigned long long)(unsigned)__float_as_int(v) | ((unsigned long long)ep << 32);
    };

    for (;;) {
        __syncthreads();   // every warp is done with the previous strip
        if (threadIdx.x == 0) s_ticket = atomicAdd(p.ticket, 1);
        if (threadIdx.x < NFLAGS) {
            int v = 0;
            // consumers that do not run in this pass never hold a slot
            if (threadIdx.x >= F_DONE && threadIdx.x < F_DONE + NDONE && threadIdx.x != F_DONE + 3 && !do_send) v = 0x7fffffff;
            if (threadIdx.x == F_DONE + 3 && !do_round) v = 0x7fffffff;
            s_flag[threadIdx.x] = v;
        }
        __syncthreads();
        const int ts = s_ticket;
        if (p.rec && lane == 0) {
            volatile int4 *r = reinterpret_cast<volatile int4 *>(p.rec) + (blockIdx.x * 8 + warp);
            r->x = -2; r->y = ts; r->z = 2; r->w = (int)ep * 4 + PASS * 2 + (do_round ? 1 : 0);
        }
        if (ts >= p.S) break;
        const int fs = (PASS == PASS_BWD) ? p.S - 1 - ts : ts;
        const int sg0 = __ldg(p.seg_ptr + fs), sg1 = __ldg(p.seg_ptr + fs + 1);
        const int n_nodes = __ldg(p.strip_len + fs);
        if (sg1 <= sg0 || n_nodes <= 0) continue;
        // flight recorder (SB_TRWS_RECORD): where every warp of every CTA last was
        auto mark = [&](int stage, int node) {
            if (p.rec && lane == 0) {
                volatile int4 *r = reinterpret_cast<volatile int4 *>(p.rec) + (blockIdx.x * 8 + warp);
                r->x = fs; r->y = node; r->z = stage; r->w = (int)ep * 4 + PASS * 2 + (do_round ? 1 : 0);
            }
        };
        mark(1, 0);

        if (warp == W_CHAIN || warp == W_CHAIN1) {
            // ============================================================ chain warps
            // Each owns ONE of the two messages to the next node of the strip and keeps it in
            // registers; the partner's message arrives through an exchange row in shared memory.
            // Both form the node sum (redundantly); chain warp 0 publishes it for the side warps.
            if (!do_send) continue;
            const int cw = warp - W_CHAIN;
            Pair<float> *P0 = scratch + (size_t)cw * scratch_pairs<K>();
            int *my_x = &s_flag[F_X + cw], *other_x = &s_flag[F_X + (cw ^ 1)];
            int *my_done = &s_flag[cw ? F_DONE + 4 : F_DONE + 0];
            const long long t_strip = p.prof ? clock64() : 0;
            Cursor cu;
            cu.start(segs, sg0);
            SegOwn mine;
            int has_mine = 0, use_carry = 0, has_dyn = 0;
            float gamma = 1.f;
            auto load_seg = [&](int sg) {
                const Segment *g = segs + sg;
                use_carry = __ldg(&g->use_carry);
                gamma = 1.f / (float)__ldg(&g->gamma_den);
                const int nitems = __ldg(&g->nitems);
                const int kind = lane < nitems ? (__ldg(&g->item[lane].kind) & 255) : S_NONE;
                has_dyn = __ballot_sync(0xffffffffu, kind == S_DYN) != 0;
                int ncs = 0;
                has_mine = 0;
                mine.flags = 0;
#pragma unroll
                for (int sl = 0; sl < 8; sl++) {
                    const SegOwn so = ld_own(g, sl);
                    if ((so.flags & OWN_HAS) && (so.flags & OWN_TO_NEXT)) {
                        if (ncs == cw) { mine = so; has_mine = 1; }
                        ncs++;
                    }
                }
            };
            load_seg(cu.sg);
            Own<K> cur, nxt;
            cur.flags = nxt.flags = 0;
            if (has_mine) load_own(mine, 0, cur);
            int h_seen = 0, p_seen = 0, s0_seen = 0, s1_seen = 0, x_seen = 0;
            float base[K], carry[K];
#pragma unroll
            for (int k = 0; k < K; k++) carry[k] = 0.f;
            {
                const long long t0 = p.prof ? clock64() : 0;
                wait_ge(&s_flag[F_H], 1, h_seen);
                if (p.prof) prof_c[1] += clock64() - t0;
            }
            row_lds<float, K>(base, slot_row(0, SR_BASE), lane);
            for (int node = 0; node < n_nodes; node++) {
                mark(10, node);
                float Di[K];
#pragma unroll
                for (int k = 0; k < K; k++) Di[k] = base[k];
                // the partner's message of the previous node (also orders the reuse of the exchange rows)
                if (node > 0) {
                    const long long t0 = p.prof ? clock64() : 0;
                    wait_ge(other_x, node, x_seen);
                    if (p.prof) prof_c[3] += clock64() - t0;
                    if (use_carry) {
                        float v[K];
                        row_lds<float, K>(v, xbuf + (size_t)((((node - 1) & 1) << 1) + (cw ^ 1)) * LP, lane);
#pragma unroll
                        for (int k = 0; k < K; k++) Di[k] += carry[k] + v[k];
                    }
                }
                if (has_dyn) {
                    mark(11, node);
                    const long long t0 = p.prof ? clock64() : 0;
                    wait_ge(&s_flag[F_P], node + 1, p_seen);
                    if (p.prof) prof_c[2] += clock64() - t0;
                    float v[K];
                    row_lds<float, K>(v, slot_row(node, SR_DYN), lane);
#pragma unroll
                    for (int k = 0; k < K; k++) Di[k] += v[k];
                }
                if (PASS == PASS_BWD) {
                    // ComputeAndSubtractMin + lower bound (minimize.cpp:79-81)
                    float vmin = BIG;
#pragma unroll
                    for (int k = 0; k < K; k++)
                        if (lane * K + k < p.L) vmin = min(vmin, Di[k]);
                    vmin = warp_min(vmin);
#pragma unroll
                    for (int k = 0; k < K; k++) Di[k] -= vmin;
                    if (cw == 0) acc_lb += (double)vmin;
                }
                if (cw == 0) {
                    // the node sum for the side warps
                    mark(12, node);
                    wait_ge(&s_flag[F_DONE + 1], node - SLOTS + 1, s0_seen);
                    wait_ge(&s_flag[F_DONE + 2], node - SLOTS + 1, s1_seen);
                    row_sts<float, K>(slot_row(node, SR_DI), Di, lane);
                    __syncwarp();
                    if (lane == 0) {
                        st_flag_cta(&s_flag[F_C], node + 1);
                        st_flag_cta(my_done, node + 1);
                    }
                } else {
                    __syncwarp();
                    if (lane == 0) st_flag_cta(my_done, node + 1);
                }
                // operands of the next node
                const float cur_gamma = gamma;
                const int cur_has = has_mine;
                const bool has_next = node + 1 < n_nodes;
                if (has_next) {
                    if (cu.advance(segs)) load_seg(cu.sg);
                    nxt.flags = 0;
                    if (has_mine) load_own(mine, cu.i, nxt);
                }
                // my message to the next node of the strip
                mark(13, node);
                const long long t_upd = p.prof ? clock64() : 0;
#pragma unroll
                for (int k = 0; k < K; k++) carry[k] = 0.f;
                if (cur_has) {
                    const float vm = update_one<K, KERN>(cur_gamma, p.lambda, p.L, lane, Di, cur, P0);
                    if (PASS == PASS_BWD) acc_lb += (double)vm;
#pragma unroll
                    for (int k = 0; k < K; k++) carry[k] = cur.m[k];
                    row_sts<float, K>(xbuf + (size_t)(((node & 1) << 1) + cw) * LP, cur.m, lane);
                }
                __syncwarp();
                if (lane == 0) st_flag_cta(my_x, node + 1);
                if (cur_has) VecIO<float, K>::store(p.msg + cur.term * LP + lane * K, cur.m);
                if (p.prof) { prof_c[5] += clock64() - t_upd + (long long)(carry[0] != carry[0]); }
                mark(16, node);
                if (has_next) {
                    mark(14, node);
                    const long long t0 = p.prof ? clock64() : 0;
                    wait_ge(&s_flag[F_H], node + 2, h_seen);
                    if (p.prof) prof_c[1] += clock64() - t0;
                    row_lds<float, K>(base, slot_row(node + 1, SR_BASE), lane);
                    cur = nxt;
                }
            }
            if (p.prof && cw == 0) {
                prof_c[0] = n_nodes;
                prof_c[4] = clock64() - t_strip;
                if (lane == 0)
                    for (int q = 0; q < 6; q++) atomicAdd((unsigned long long *)p.prof + (fs == 0 ? 0 : 8) + q, (unsigned long long)prof_c[q]);
            }
            for (int q = 0; q < 6; q++) prof_c[q] = 0;
            mark(19, n_nodes);
            continue;
        }

        if (warp == W_SIDE0 || warp == W_SIDE1) {
            // ============================================================ side warps
            if (!do_send) continue;
            const int sid = warp - W_SIDE0;
            Pair<float> *P = scratch + (size_t)(W_SIDE0 + sid) * scratch_pairs<K>();
            int *my_done = &s_flag[F_DONE + 1 + sid];
            Cursor cu;
            cu.start(segs, sg0);
            SegOwn so[4];
            int nmy = 0;
            float gamma = 1.f;
            auto load_seg = [&](int sg) {
                const Segment *g = segs + sg;
                gamma = 1.f / (float)__ldg(&g->gamma_den);
                nmy = 0;
                int nside = 0;
                so[0].flags = so[1].flags = so[2].flags = so[3].flags = 0;
#pragma unroll
                for (int sl = 0; sl < 8; sl++) {
                    const SegOwn s = ld_own(g, sl);
                    if ((s.flags & OWN_HAS) && !(s.flags & OWN_TO_NEXT)) {
                        if ((nside & 1) == sid) {
                            if (nmy == 0) so[0] = s; else if (nmy == 1) so[1] = s; else if (nmy == 2) so[2] = s; else so[3] = s;
                            nmy++;
                        }
                        nside++;
                    }
                }
            };
            load_seg(cu.sg);
            Own<K> cur, nxt;
            cur.flags = 0;
            if (nmy > 0) load_own(so[0], 0, cur);
            int c_seen = 0;
            for (int node = 0; node < n_nodes; node++) {
                const int cur_n = nmy, cur_i = cu.i;
                const float cur_gamma = gamma;
                const SegOwn c1 = so[1], c2 = so[2], c3 = so[3];
                const bool has_next = node + 1 < n_nodes;
                if (has_next && cu.advance(segs)) load_seg(cu.sg);
                if (cur_n == 0) {
                    if (lane == 0) st_flag_cta(my_done, node + 1);
                    if (has_next && nmy > 0) load_own(so[0], cu.i, cur);
                    continue;
                }
                nxt.flags = 0;
                if (has_next && nmy > 0) load_own(so[0], cu.i, nxt);
                mark(21, node);
                wait_ge(&s_flag[F_C], node + 1, c_seen);
                float Di[K];
                row_lds<float, K>(Di, slot_row(node, SR_DI), lane);
                __syncwarp();
                if (lane == 0) st_flag_cta(my_done, node + 1);
                mark(22, node);
                for (int t = 0; t < cur_n; t++) {
                    if (t == 1) load_own(c1, cur_i, cur);
                    if (t == 2) load_own(c2, cur_i, cur);
                    if (t == 3) load_own(c3, cur_i, cur);
                    const float vm = update_one<K, KERN>(cur_gamma, p.lambda, p.L, lane, Di, cur, P);
                    mark(23, node);
                    if (PASS == PASS_BWD) acc_lb += (double)vm;
                    const int peer = (cur.flags & OWN_PEER_UP) ? 0 : (cur.flags & OWN_PEER_DOWN) ? 1 : -1;
                    if (peer >= 0) {
                        // receiver on a neighbouring GPU: push the words and the mirror of the message row
                        unsigned long long *mb = p.peer_mbox[peer] + cur.term * LP + lane * K;
#pragma unroll
                        for (int k = 0; k < K; k++) st_mbox_sys(mb + k, mbox_word(cur.m[k]));
                        VecIO<float, K>::store(p.peer_msg[peer] + cur.term * LP + lane * K, cur.m);
                    } else {
                        unsigned long long *mb = p.mbox + cur.term * LP + lane * K;
#pragma unroll
                        for (int k = 0; k < K; k++) st_mbox(mb + k, mbox_word(cur.m[k]));
                    }
                    mark(24, node);
                    VecIO<float, K>::store(p.msg + cur.term * LP + lane * K, cur.m);
                    mark(25, node);
                }
                cur = nxt;
            }
            mark(29, n_nodes);
            continue;
        }

        if (warp == W_ROUND) {
            // ============================================================ round warp
            if (!do_round) continue;
            Cursor cu;
            cu.start(segs, sg0);
            int u0 = 0, du = 0, use_carry = 0, ncs = 0, nside = 0;
            SegOwn cs[2];
            SegOwn my_side;     // lane j < nside: the j-th cross-strip send term
            auto load_seg = [&](int sg) {
                const Segment *g = segs + sg;
                u0 = __ldg(&g->u0);
                du = __ldg(&g->du);
                use_carry = __ldg(&g->use_carry);
                ncs = 0;
                nside = 0;
                cs[0].flags = cs[1].flags = 0;
                my_side.flags = 0;
#pragma unroll
                for (int sl = 0; sl < 8; sl++) {
                    const SegOwn s = ld_own(g, sl);
                    if (!(s.flags & OWN_HAS)) continue;
                    if (s.flags & OWN_TO_NEXT) {
                        if (ncs == 0) cs[0] = s; else cs[1] = s;
                        ncs++;
                    } else {
                        if (lane == nside) my_side = s;
                        nside++;
                    }
                }
            };
            load_seg(cu.sg);
            // positions on the to-next terms: sender's (this node's labels) and receiver's (next node's)
            float ns_[2][K], nx_[2][K], nal[2];
            float px[2][K], pal[2], psv[2];   // of the previous node
            int pn = 0;
            auto load_next_terms = [&](int i) {
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    if (t < ncs) {
                        const long long term = cs[t].term0 + (long long)i * cs[t].tstride;
                        const long long off = term * LP + lane * K;
                        const bool tail = (cs[t].flags & OWN_TAIL) != 0;
                        VecIO<float, K>::load_ro(ns_[t], (tail ? p.posqp : p.posq) + off);
                        VecIO<float, K>::load_ro(nx_[t], (tail ? p.posq : p.posqp) + off);
                        nal[t] = __ldg(p.alpha + term);
                    }
                }
            };
            load_next_terms(0);
            int h_seen = 0, p_seen = 0;
            for (int node = 0; node < n_nodes; node++) {
                const int u = u0 + cu.i * du;
                const int cur_i = cu.i, cur_ncs = ncs, cur_nside = nside, cur_carry = use_carry;
                const SegOwn cur_side = my_side;
                mark(31, node);
                wait_ge(&s_flag[F_H], node + 1, h_seen);
                const int nrnd = s_nrnd[node & (SLOTS - 1)];
                mark(32, node);
                if (nrnd > 0) wait_ge(&s_flag[F_P], node + 1, p_seen);
                mark(33, node);
                // minimize.cpp:240-260: DiB = D + sum_{lower nb} V(x_nb, .), Dr = DiB + forward messages
                float pv[K], bs[K], dd[K];
                row_lds<float, K>(bs, slot_row(node, SR_BASE), lane);
                row_lds<float, K>(dd, slot_row(node, SR_D), lane);
#pragma unroll
                for (int k = 0; k < K; k++) pv[k] = 0.f;
                for (int t = 0; t < nrnd; t++) {
                    float v[K];
                    row_lds<float, K>(v, slot_row(node, SR_XR + t), lane);
                    const float aj = s_alpha[node & (SLOTS - 1)][t], sj = s_sel[node & (SLOTS - 1)][t];
#pragma unroll
                    for (int k = 0; k < K; k++) pv[k] += aj * smooth<float, KERN>(v[k] - sj, p.lambda);
                }
                if (cur_carry) {
#pragma unroll
                    for (int t = 0; t < 2; t++)
                        if (t < pn) {
#pragma unroll
                            for (int k = 0; k < K; k++) pv[k] += pal[t] * smooth<float, KERN>(px[t][k] - psv[t], p.lambda);
                        }
                }
                __syncwarp();
                if (lane == 0) st_flag_cta(&s_flag[F_DONE + 3], node + 1);
                // Vector::ComputeMin: first minimum in label order (typeStereoLinear.h:238-252)
                float best = BIG;
                int bi = 0x7fffffff;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int lbl = lane * K + k;
                    const float dr = bs[k] + pv[k];
                    if (lbl < p.L && dr < best) { best = dr; bi = lbl; }
                }
                const float wbest = warp_min(best);
                bi = warp_min_s32(best == wbest ? bi : 0x7fffffff);
                {
                    float dv = dd[0] + pv[0];
#pragma unroll
                    for (int k = 1; k < K; k++)
                        if (k == bi % K) dv = dd[k] + pv[k];
                    dv = __shfl_sync(0xffffffffu, dv, bi / K);
                    if (lane == 0) {
                        p.sol[u] = bi;
                        acc_energy += (double)dv;
                    }
                }
                // position of the rounded label on the terms to the next node of the strip
                pn = cur_ncs;
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    if (t < cur_ncs) {
                        float sv = ns_[t][0];
#pragma unroll
                        for (int k = 1; k < K; k++)
                            if (k == bi % K) sv = ns_[t][k];
                        psv[t] = __shfl_sync(0xffffffffu, sv, bi / K);
                        pal[t] = nal[t];
#pragma unroll
                        for (int k = 0; k < K; k++) px[t][k] = nx_[t][k];
                    }
                }
                // ... and on the terms to other strips: their receivers poll these words
                if (lane < cur_nside) {
                    const long long term = cur_side.term0 + (long long)cur_i * cur_side.tstride;
                    const bool tail = (cur_side.flags & OWN_TAIL) != 0;
                    const float sv = __ldg((tail ? p.posqp : p.posq) + term * LP + bi);
                    const int peer = (cur_side.flags & OWN_PEER_UP) ? 0 : (cur_side.flags & OWN_PEER_DOWN) ? 1 : -1;
                    if (peer >= 0) st_mbox_sys(p.peer_selbox[peer] + term, mbox_word(sv));
                    else st_mbox(p.selbox + term, mbox_word(sv));
                }
                if (node + 1 < n_nodes) {
                    if (cu.advance(segs)) load_seg(cu.sg);
                    load_next_terms(cu.i);
                }
            }
            mark(39, n_nodes);
            continue;
        }

        if (warp == W_STATIC) {
            // ============================================================ static warp
            Cursor ci, cc;            // issue cursor (LAND - 1 nodes ahead) and consume cursor
            ci.start(segs, sg0);
            cc.start(segs, sg0);
            int i_kind = S_NONE, i_tstride = 0, c_kind = S_NONE;
            long long i_term0 = 0;
            auto load_items_issue = [&](int sg) {
                const Segment *g = segs + sg;
                const int nitems = __ldg(&g->nitems);
                i_kind = S_NONE;
                if (lane < nitems) {
                    i_kind = __ldg(&g->item[lane].kind);
                    i_term0 = __ldg(&g->item[lane].term0);
                    i_tstride = __ldg(&g->item[lane].tstride);
                }
            };
            auto load_items_consume = [&](int sg) {
                const Segment *g = segs + sg;
                const int nitems = __ldg(&g->nitems);
                c_kind = lane < nitems ? __ldg(&g->item[lane].kind) : S_NONE;
            };
            auto is_static = [&](int kind) {
                const int k0 = kind & 255;
                return k0 == S_D || k0 == S_SEND || (k0 == S_RND && do_round);
            };
            int issued = 0;   // nodes issued so far
            auto issue = [&]() {
                if (issued < n_nodes) {
                    const int k0 = i_kind & 255;
                    const long long term = i_term0 + (long long)ci.i * i_tstride;
                    const bool st = is_static(i_kind);
                    const float *src = nullptr;
                    if (st) src = (k0 == S_D ? p.D : k0 == S_SEND ? p.msg : ((i_kind & 256) ? p.posqp : p.posq)) + term * LP;
                    unsigned rem = __ballot_sync(0xffffffffu, st);
                    const unsigned m_rnd = __ballot_sync(0xffffffffu, k0 == S_RND && do_round);
                    float *dst = land + (size_t)(issued & (LAND - 1)) * MAX_STATIC * LP;
                    while (rem) {
                        const int jl = __ffs(rem) - 1;
                        rem &= rem - 1;
                        const char *sp = reinterpret_cast<const char *>(shfl_ptr(src, jl));
#pragma unroll
                        for (int ch = lane; ch < CH; ch += 32) cp_async16(reinterpret_cast<char *>(dst) + ch * 16, sp + ch * 16);
                        dst += LP;
                    }
                    if (k0 == S_RND && do_round) {
                        const int t = __popc(m_rnd & ((1u << lane) - 1u));
                        if (t < MAX_RND) cp_async4(&s_land_alpha[issued & (LAND - 1)][t], p.alpha + term);
                    }
                    issued++;
                    if (issued < n_nodes && ci.advance(segs)) load_items_issue(ci.sg);
                }
                cp_async_commit();
            };
            load_items_issue(ci.sg);
            load_items_consume(cc.sg);
            for (int q = 0; q < LAND - 1; q++) issue();
            int free_seen = 0;
            for (int node = 0; node < n_nodes; node++) {
                mark(40, node);
                issue();
                mark(44, node);
                if (p.debug & 4) cp_async_wait_group<0>();
                mark(41, node);
                cp_async_wait_group<LAND - 1>();
                __syncwarp();
                mark(42, node);
                // slot of node - SLOTS released by every consumer
                wait_free(&s_flag[F_DONE], node - SLOTS + 1, free_seen);
                mark(43, node);
                const int k0 = c_kind & 255;
                const unsigned m_d = __ballot_sync(0xffffffffu, k0 == S_D);
                const unsigned m_send = __ballot_sync(0xffffffffu, k0 == S_SEND);
                const unsigned m_rnd = __ballot_sync(0xffffffffu, k0 == S_RND && do_round);
                unsigned rem = m_d | m_send | m_rnd;
                const float *src = land + (size_t)(node & (LAND - 1)) * MAX_STATIC * LP;
                float base[K];
#pragma unroll
                for (int k = 0; k < K; k++) base[k] = 0.f;
                int t = 0;
                while (rem) {
                    const int jl = __ffs(rem) - 1;
                    rem &= rem - 1;
                    float v[K];
                    row_lds<float, K>(v, src, lane);
                    src += LP;
                    if ((m_rnd >> jl) & 1u) {
                        if (t < MAX_RND) row_sts<float, K>(slot_row(node, SR_XR + t), v, lane);
                        t++;
                    } else {
#pragma unroll
                        for (int k = 0; k < K; k++) base[k] += v[k];
                        if (((m_d >> jl) & 1u) && do_round) row_sts<float, K>(slot_row(node, SR_D), v, lane);
                    }
                }
                row_sts<float, K>(slot_row(node, SR_BASE), base, lane);
                if (do_round) {
                    if (lane < t && lane < MAX_RND) s_alpha[node & (SLOTS - 1)][lane] = s_land_alpha[node & (LAND - 1)][lane];
                    if (lane == 0) s_nrnd[node & (SLOTS - 1)] = min(t, MAX_RND);
                }
                __syncwarp();
                if (lane == 0) st_flag_cta(&s_flag[F_H], node + 1);
                if (node + 1 < n_nodes && cc.advance(segs)) load_items_consume(cc.sg);
            }
            cp_async_wait_group<0>();
            mark(49, n_nodes);
            continue;
        }

        if (warp == W_POLL) {
            // ============================================================ poll warp
            Cursor cu;
            cu.start(segs, sg0);
            int kind = S_NONE, tstride = 0;
            long long term0 = 0;
            auto load_items = [&](int sg) {
                const Segment *g = segs + sg;
                const int nitems = __ldg(&g->nitems);
                kind = S_NONE;
                if (lane < nitems) {
                    kind = __ldg(&g->item[lane].kind);
                    term0 = __ldg(&g->item[lane].term0);
                    tstride = __ldg(&g->item[lane].tstride);
                }
            };
            load_items(cu.sg);
            const bool sys_scope = p.world > 1;
            int free_seen = 0;
            for (int node = 0; node < n_nodes; node++) {
                const int k0 = kind & 255;
                const bool is_dyn = (k0 == S_DYN) && do_send, is_rnd = (k0 == S_RND) && do_round;
                const long long term = term0 + (long long)cu.i * tstride;
                const unsigned m_dyn = __ballot_sync(0xffffffffu, is_dyn);
                const unsigned m_rnd = __ballot_sync(0xffffffffu, is_rnd);
                mark(50, node);
                if (m_dyn | m_rnd) {
                    mark(51, node);
                    wait_free(&s_flag[F_DONE], node - SLOTS + 1, free_seen);
                    mark(52, node);
                    float dyn[K];
#pragma unroll
                    for (int k = 0; k < K; k++) dyn[k] = 0.f;
                    float sel = 0.f;
                    constexpr int G = (K <= 2) ? 4 : (K <= 4) ? 2 : 1;
                    bool have_sel = !is_rnd;
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
                                if ((unsigned)(sw >> 32) == ep) { sel = __int_as_float((int)(unsigned)sw); have_sel = true; }
                                else ok = false;
                            }
#pragma unroll
                            for (int q = 0; q < G; q++)
                                if (q < nb) {
#pragma unroll
                                    for (int k = 0; k < K; k++) ok = ok && ((unsigned)(wv[q][k] >> 32) == ep);
                                }
                            if (__all_sync(0xffffffffu, ok)) break;
                            if (p.debug & 8) __nanosleep(100);
                        }
#pragma unroll
                        for (int q = 0; q < G; q++)
                            if (q < nb) {
#pragma unroll
                                for (int k = 0; k < K; k++) dyn[k] += __int_as_float((int)(unsigned)wv[q][k]);
                            }
                    } while (rem);
                    mark(53, node);
                    if (m_dyn) row_sts<float, K>(slot_row(node, SR_DYN), dyn, lane);
                    if (is_rnd) {
                        const int t = __popc(m_rnd & ((1u << lane) - 1u));
                        if (t < MAX_RND) s_sel[node & (SLOTS - 1)][t] = sel;
                    }
                    __syncwarp();
                }
                if (lane == 0) st_flag_cta(&s_flag[F_P], node + 1);
                if (node + 1 < n_nodes && cu.advance(segs)) load_items(cu.sg);
            }
            mark(59, n_nodes);
            continue;
        }

        // ================================================================ prefetch warp
        {
            // rows of the update operands (message, both position rows, rank and count bytes) of the
            // send terms of the nodes ahead -> L2; paced by the consumer that runs in this pass
            const int *pace = do_send ? &s_flag[F_DONE + 0] : &s_flag[F_DONE + 3];
            Cursor cu;
            cu.start(segs, sg0);
            int pf_node = 0;
            constexpr int LR = (LP * 4 + 127) / 128;
            constexpr int LB = (LP + 127) / 128;
            while (pf_node < n_nodes && !(p.debug & 1)) {
                const int c = __shfl_sync(0xffffffffu, ld_flag_cta(pace), 0);   // warp-uniform
                mark(60, pf_node * 1000 + min(c, 999));
                const int pf_end = min(n_nodes, c + PF_AHEAD);
                if (pf_node >= pf_end) {
                    __nanosleep(100);
                    continue;
                }
                for (; pf_node < pf_end; pf_node++) {
                    if (pf_node > c + 1 && lane < 8) {
                        const SegOwn o = ld_own(segs + cu.sg, lane);
                        if (o.flags & OWN_HAS) {
                            const long long row = (o.term0 + (long long)cu.i * o.tstride) * LP;
                            const bool tail = (o.flags & OWN_TAIL) != 0;
                            if (do_send) {
                                for (int t = 0; t < LR; t++) {
                                    prefetch_l2(reinterpret_cast<const char *>(p.msg + row) + t * 128);
                                    prefetch_l2(reinterpret_cast<const char *>(p.posq + row) + t * 128);
                                    prefetch_l2(reinterpret_cast<const char *>(p.posqp + row) + t * 128);
                                }
                                for (int t = 0; t < LB; t++) {
                                    prefetch_l2(reinterpret_cast<const char *>((tail ? p.rank_qp : p.rank_q) + row) + t * 128);
                                    prefetch_l2(reinterpret_cast<const char *>((tail ? p.cnt_q : p.cnt_qp) + row) + t * 128);
                                }
                            } else if (o.flags & OWN_TO_NEXT) {
                                for (int t = 0; t < LR; t++) {
                                    prefetch_l2(reinterpret_cast<const char *>(p.posq + row) + t * 128);
                                    prefetch_l2(reinterpret_cast<const char *>(p.posqp + row) + t * 128);
                                }
                            }
                        }
                    }
                    if (pf_node > c + 1 && lane >= 8 && lane - 8 < SCHED_ITEMS) {
                        // mailbox rows the poll warp will read: a row written long ago (or not yet) is
                        // pulled from HBM into L2 ahead of the poll, the sender's store then hits L2
                        const Segment *g = segs + cu.sg;
                        if (lane - 8 < __ldg(&g->nitems)) {
                            const int kind = __ldg(&g->item[lane - 8].kind) & 255;
                            const long long term = __ldg(&g->item[lane - 8].term0) + (long long)cu.i * __ldg(&g->item[lane - 8].tstride);
                            if (kind == S_DYN && do_send) {
                                constexpr int LM = (LP * 8 + 127) / 128;
                                for (int t = 0; t < LM; t++) prefetch_l2(reinterpret_cast<const char *>(p.mbox + term * LP) + t * 128);
                            } else if (kind == S_RND && do_round) {
                                prefetch_l2(p.selbox + term);
                            }
                        }
                    }
                    if (pf_node + 1 < n_nodes) cu.advance(segs);
                }
            }
            mark(69, n_nodes);
        }
    }
    if (p.rec && lane == 0) {
        volatile int4 *r = reinterpret_cast<volatile int4 *>(p.rec) + (blockIdx.x * 8 + warp);
        r->z = 3;
    }
    if (lane == 0) {
        if (acc_energy != 0.0) atomicAdd(p.acc + 0, acc_energy);
        if (acc_lb != 0.0) atomicAdd(p.acc + 1, acc_lb);
    }
}

} // namespace v5
} // namespace trws
} // namespace sb
