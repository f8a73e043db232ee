// trws_sweep5.cuh -- EXPERIMENTAL fp32 TRW-S sweep kernel, fifth generation (SB_TRWS_SWEEP=5).
//
// Same work as sweep_kernel of trws_kernels.cuh (Minimize_TRW_S forward / backward sweeps,
// cpp/trw-s/minimize.cpp:31-95; UpdateMessage, typeStereoLinear.h:329-487 /
// typeStereoQuadratic.h:329-501; ComputeSolutionAndEnergy, minimize.cpp:223-264), same strip /
// segment schedule (trws_sched.h), same mailbox protocol between strips, same parity tests
// (tests/test_trws_v5_gpu.py) -- but a different division of labour inside the CTA:
//
//   chain warps  (2) each own ONE of the two messages a node sends to the next node of its strip
//                and keep it in registers; the partner's message arrives through an exchange row
//                in shared memory.  Both form the node sum Di; warp 0 publishes it.
//   side warps   (2) own the messages to nodes of other strips: they pick Di up from shared
//                memory, update, and write the message + its self-validating mailbox words.
//   round warp   carries the primal rounding (minimize.cpp:240-260) as its own, much shorter,
//                chain: rounded label of the previous node -> pairwise column -> arg-min.
//   static warp  streams everything that does not depend on this pass (unary row, old messages
//                of the send terms, position rows for the rounding) through a cp.async ring
//                LAND nodes deep and reduces it to BASE = D + sum(old messages).
//   operand warp copies the update operands of the node's send terms (position rows, rank and
//                merge-count bytes, alpha) LAND - 1 nodes ahead straight into the node's slot in
//                shared memory: chain and side warps never touch global memory for operands.
//   poll warp    polls the mailbox words of the messages arriving from other strips and the
//                selected positions of their rounded labels.
//
// Hand-over between the warps is by monotone node counters in shared memory (plain volatile
// accesses, see ld_flag_cta) over a ring of SLOTS node slots; nothing in the CTA executes a
// CTA-wide barrier inside a strip.  Measured against sweep_kernel: equal at
// 128x160x64 and 256x256x256, 25 % slower at 375x450x64 -- its chain step is ~2.6 k cycles
// (update 970, the rest counter waits, shared-memory reads and bookkeeping at an instruction-level
// parallelism of one).  Not the default.
#pragma once
#include "trws_kernels.cuh"

namespace sb {
namespace trws {
namespace v5 {

// warp w issues on scheduler w % 4: the two chain warps and the two side warps each get a scheduler
// of their own for the updates; round / poll / static / operand warps are light or mostly waiting
enum { W_CHAIN = 0, W_CHAIN1 = 1, W_SIDE0 = 2, W_SIDE1 = 3, W_ROUND = 4, W_POLL = 5, W_STATIC = 6, W_OPS = 7, NWARPS = 8 };
constexpr int THREADS = NWARPS * 32;
// node slots between producers and consumers / depth of the static warp's cp.async ring (powers of
// two; the ring runs LAND - 1 nodes ahead of what it publishes and copies the update operands
// straight into the slot of the node, so SLOTS must exceed LAND by a margin)
template <int K> struct Depth {
    static constexpr int SLOTS = (K <= 2) ? 8 : 4;
    static constexpr int LAND = (K <= 2) ? 4 : 2;
};
constexpr int MAX_SLOTS = 8, MAX_LAND = 4;
constexpr int MAX_OPS = 4;        // send terms of a node whose operands are staged in its slot (first half)
constexpr int MAX_RND = 6;        // rounding rows per node (lower neighbours in other strips: <= 3 pairs)
constexpr int MAX_STATIC = 1 + 8 + MAX_RND;
// rows of a node slot; per staged send term: old message, sender positions, receiver positions, then
// one row holding the rank bytes (first LP bytes) and the merge-count bytes (next LP bytes)
enum { SR_BASE = 0, SR_D = 1, SR_DYN = 2, SR_DI = 3, SR_XR = 4, SR_OPS = 4 + MAX_RND, OP_M = 0, OP_S = 1, OP_X = 2, OP_RC = 3,
       SLOT_ROWS = 4 + MAX_RND + 4 * MAX_OPS };
enum { F_H = 0, F_P = 1, F_C = 2, F_DONE = 3 /* + chain0, side0, side1, round, chain1 */, NDONE = 5, F_X = 8 /* + chain warp */, F_O = 10, NFLAGS = 11 };
constexpr int XROWS = 4;          // exchange rows of the two chain warps: [node parity][warp]

template <int K> __host__ __device__ constexpr size_t smem_bytes()
{
    return (size_t)(Depth<K>::SLOTS * SLOT_ROWS + XROWS + Depth<K>::LAND * MAX_STATIC) * 32 * K * sizeof(float) +
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

template <int K> struct Own {
    float m[K], s[K], x[K];
    PackedBytes<K> rkp, cnp;
    float alpha;
    long long term;
    int flags;
};

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

// DBG = true: the instantiation with the cycle counters (SB_TRWS_PROFILE) and the flight recorder
// (SB_TRWS_RECORD); the production instantiation carries neither.
template <int K, int KERN, int PASS, bool DBG>
__global__ void __launch_bounds__(THREADS, 1) sweep5_kernel(const Problem<float> p)
{
    long long *const prof_buf = DBG ? p.prof : nullptr;
    int *const rec_buf = DBG ? p.rec : nullptr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ticket;
    __shared__ int s_flag[NFLAGS];
    constexpr int SLOTS = Depth<K>::SLOTS, LAND = Depth<K>::LAND;
    __shared__ float s_sel[MAX_SLOTS][MAX_RND], s_alpha[MAX_SLOTS][MAX_RND], s_op_alpha[MAX_SLOTS][MAX_OPS];
    __shared__ float s_land_alpha[MAX_LAND][MAX_RND];
    __shared__ int s_nrnd[MAX_SLOTS];
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
    const long long t_begin = prof_buf ? clock64() : 0;
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
    // operands staged by the static warp in the node's slot (send terms of the first half); the few
    // nodes that send on more than MAX_OPS terms fetch the others from global memory
    auto fetch_own = [&](const SegOwn &so, int sl, int i, int node, Own<K> &o) {
        if (sl < MAX_OPS) {
            o.flags = so.flags;
            o.term = so.term0 + (long long)i * so.tstride;
            const float *ops = slot_row(node, SR_OPS + 4 * sl);
            row_lds<float, K>(o.m, ops + OP_M * LP, lane);
            row_lds<float, K>(o.s, ops + OP_S * LP, lane);
            row_lds<float, K>(o.x, ops + OP_X * LP, lane);
            const uint8_t *rc = reinterpret_cast<const uint8_t *>(ops + OP_RC * LP);
            o.rkp.load_shared(rc + lane * K);
            o.cnp.load_shared(rc + LP + lane * K);
            o.alpha = s_op_alpha[node & (SLOTS - 1)][sl];
        } else {
            load_own(so, i, o);
        }
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
        if (rec_buf && lane == 0) {
            volatile int4 *r = reinterpret_cast<volatile int4 *>(rec_buf) + (blockIdx.x * 8 + warp);
            r->x = -2; r->y = ts; r->z = 2; r->w = (int)ep * 4 + PASS * 2 + (do_round ? 1 : 0);
        }
        if (ts >= p.S) break;
        const int fs = (PASS == PASS_BWD) ? p.S - 1 - ts : ts;
        const int sg0 = __ldg(p.seg_ptr + fs), sg1 = __ldg(p.seg_ptr + fs + 1);
        const int n_nodes = __ldg(p.strip_len + fs);
        if (sg1 <= sg0 || n_nodes <= 0) continue;
        // flight recorder (SB_TRWS_RECORD): where every warp of every CTA last was
        auto mark = [&](int stage, int node) {
            if (rec_buf && lane == 0) {
                volatile int4 *r = reinterpret_cast<volatile int4 *>(rec_buf) + (blockIdx.x * 8 + warp);
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
            const long long t_strip = prof_buf ? clock64() : 0;
            Cursor cu;
            cu.start(segs, sg0);
            SegOwn mine;
            int has_mine = 0, my_sl = 0, use_carry = 0, has_dyn = 0;
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
                        if (ncs == cw) { mine = so; has_mine = 1; my_sl = sl; }
                        ncs++;
                    }
                }
            };
            load_seg(cu.sg);
            Own<K> cur;
            cur.flags = 0;
            int h_seen = 0, p_seen = 0, s0_seen = 0, s1_seen = 0, x_seen = 0, o_seen = 0;
            float carry[K];
#pragma unroll
            for (int k = 0; k < K; k++) carry[k] = 0.f;
            for (int node = 0; node < n_nodes; node++) {
                mark(10, node);
                // static part of the node (the static warp runs ahead): BASE row and my operands
                {
                    const long long t0 = prof_buf ? clock64() : 0;
                    wait_ge(&s_flag[F_H], node + 1, h_seen);
                    if (prof_buf) prof_c[1] += clock64() - t0;
                }
                float Di[K];
                row_lds<float, K>(Di, slot_row(node, SR_BASE), lane);
                if (has_mine) {
                    wait_ge(&s_flag[F_O], node + 1, o_seen);
                    fetch_own(mine, my_sl, cu.i, node, cur);
                }
                // the partner's message of the previous node (also orders the reuse of the exchange rows)
                if (node > 0) {
                    const long long t0 = prof_buf ? clock64() : 0;
                    wait_ge(other_x, node, x_seen);
                    if (prof_buf) prof_c[3] += clock64() - t0;
                    if (use_carry) {
                        float v[K];
                        row_lds<float, K>(v, xbuf + (size_t)((((node - 1) & 1) << 1) + (cw ^ 1)) * LP, lane);
#pragma unroll
                        for (int k = 0; k < K; k++) Di[k] += carry[k] + v[k];
                    }
                }
                if (has_dyn) {
                    mark(11, node);
                    const long long t0 = prof_buf ? clock64() : 0;
                    wait_ge(&s_flag[F_P], node + 1, p_seen);
                    if (prof_buf) prof_c[2] += clock64() - t0;
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
                // my message to the next node of the strip
                mark(13, node);
                const long long t_upd = prof_buf ? clock64() : 0;
#pragma unroll
                for (int k = 0; k < K; k++) carry[k] = 0.f;
                if (has_mine) {
                    const float vm = update_one<K, KERN>(gamma, p.lambda, p.L, lane, Di, cur, P0);
                    if (PASS == PASS_BWD) acc_lb += (double)vm;
#pragma unroll
                    for (int k = 0; k < K; k++) carry[k] = cur.m[k];
                    row_sts<float, K>(xbuf + (size_t)(((node & 1) << 1) + cw) * LP, cur.m, lane);
                }
                __syncwarp();
                if (lane == 0) st_flag_cta(my_x, node + 1);
                if (has_mine) VecIO<float, K>::store(p.msg + cur.term * LP + lane * K, cur.m);
                if (prof_buf) { prof_c[5] += clock64() - t_upd + (long long)(carry[0] != carry[0]); }
                mark(16, node);
                if (node + 1 < n_nodes && cu.advance(segs)) load_seg(cu.sg);
            }
            if (prof_buf && cw == 0) {
                prof_c[0] = n_nodes;
                prof_c[4] = clock64() - t_strip;
                if (lane == 0)
                    for (int q = 0; q < 6; q++) atomicAdd((unsigned long long *)prof_buf + (fs == 0 ? 0 : 8) + q, (unsigned long long)prof_c[q]);
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
            int sls[4] = {0, 0, 0, 0};
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
                            if (nmy == 0) { so[0] = s; sls[0] = sl; }
                            else if (nmy == 1) { so[1] = s; sls[1] = sl; }
                            else if (nmy == 2) { so[2] = s; sls[2] = sl; }
                            else { so[3] = s; sls[3] = sl; }
                            nmy++;
                        }
                        nside++;
                    }
                }
            };
            load_seg(cu.sg);
            Own<K> cur;
            cur.flags = 0;
            int c_seen = 0, h_seen = 0, o_seen = 0;
            for (int node = 0; node < n_nodes; node++) {
                if (nmy == 0) {
                    if (lane == 0) st_flag_cta(my_done, node + 1);
                } else {
                    // operands (staged ahead by the static warp), then the node sum from the chain warp
                    wait_ge(&s_flag[F_H], node + 1, h_seen);
                    wait_ge(&s_flag[F_O], node + 1, o_seen);
                    fetch_own(so[0], sls[0], cu.i, node, cur);
                    mark(21, node);
                    wait_ge(&s_flag[F_C], node + 1, c_seen);
                    float Di[K];
                    row_lds<float, K>(Di, slot_row(node, SR_DI), lane);
                    mark(22, node);
                    for (int t = 0; t < nmy; t++) {
                        if (t == 1) fetch_own(so[1], sls[1], cu.i, node, cur);
                        if (t == 2) fetch_own(so[2], sls[2], cu.i, node, cur);
                        if (t == 3) fetch_own(so[3], sls[3], cu.i, node, cur);
                        if (t == nmy - 1) {
                            // last read of the slot
                            __syncwarp();
                            if (lane == 0) st_flag_cta(my_done, node + 1);
                        }
                        const float vm = update_one<K, KERN>(gamma, p.lambda, p.L, lane, Di, cur, P);
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
                        VecIO<float, K>::store(p.msg + cur.term * LP + lane * K, cur.m);
                    }
                }
                if (node + 1 < n_nodes && cu.advance(segs)) load_seg(cu.sg);
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
            int cs_sl[2] = {0, 0};
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
                        if (ncs == 0) { cs[0] = s; cs_sl[0] = sl; } else { cs[1] = s; cs_sl[1] = sl; }
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
            // (staged in the node's slot by the static warp; a term of the second half comes from global memory)
            auto load_next_terms = [&](int i, int node) {
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    if (t < ncs) {
                        if (cs_sl[t] < MAX_OPS) {
                            const float *ops = slot_row(node, SR_OPS + 4 * cs_sl[t]);
                            row_lds<float, K>(ns_[t], ops + OP_S * LP, lane);
                            row_lds<float, K>(nx_[t], ops + OP_X * LP, lane);
                            nal[t] = s_op_alpha[node & (SLOTS - 1)][cs_sl[t]];
                        } else {
                            const long long term = cs[t].term0 + (long long)i * cs[t].tstride;
                            const long long off = term * LP + lane * K;
                            const bool tail = (cs[t].flags & OWN_TAIL) != 0;
                            VecIO<float, K>::load_ro(ns_[t], (tail ? p.posqp : p.posq) + off);
                            VecIO<float, K>::load_ro(nx_[t], (tail ? p.posq : p.posqp) + off);
                            nal[t] = __ldg(p.alpha + term);
                        }
                    }
                }
            };
            int h_seen = 0, p_seen = 0, o_seen = 0;
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
                if (cur_ncs > 0) wait_ge(&s_flag[F_O], node + 1, o_seen);
                load_next_terms(cur_i, node);
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
                if (node + 1 < n_nodes && cu.advance(segs)) load_seg(cu.sg);
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
            int free_seen = 0;
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
            for (int node = 0; node < n_nodes; node++) {
                mark(40, node);
                issue();
                mark(44, node);
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
                int t = 0, ts = 0;
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
                        if ((m_d >> jl) & 1u) {
                            if (do_round) row_sts<float, K>(slot_row(node, SR_D), v, lane);
                        } else {
                            // the ts-th send row is the old message of own slot ts (build_pass_plan hands both out together)
                            if (ts < MAX_OPS) row_sts<float, K>(slot_row(node, SR_OPS + 4 * ts + OP_M), v, lane);
                            ts++;
                        }
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

        // ================================================================ operand warp
        {
            // Update operands of the node's first MAX_OPS send terms -- sender / receiver position rows,
            // rank and merge-count bytes, alpha -- copied LAND - 1 nodes ahead straight into the node's
            // slot.  No shuffles: two lanes own one (term, row) pair and copy its 16-byte chunks.
            const int sl = lane >> 3, rtype = (lane >> 1) & 3, half = lane & 1;
            Cursor ci;
            ci.start(segs, sg0);
            SegOwn my_own = ld_own(segs + ci.sg, sl);
            int issued = 0, free_seen = 0;
            constexpr int CB = LP / 16;   // 16-byte chunks of a byte row
            auto issue = [&]() {
                if (issued < n_nodes) {
                    wait_free(&s_flag[F_DONE], issued - SLOTS + 1, free_seen);   // the slot must be free
                    if (my_own.flags & OWN_HAS) {
                        const long long term = my_own.term0 + (long long)ci.i * my_own.tstride;
                        const bool tail = (my_own.flags & OWN_TAIL) != 0;
                        char *ops = reinterpret_cast<char *>(slot_row(issued, SR_OPS + 4 * sl));
                        const char *src;
                        char *dst;
                        int nch;
                        if (rtype == 0) { src = reinterpret_cast<const char *>((tail ? p.posqp : p.posq) + term * LP); dst = ops + (size_t)OP_S * LP * 4; nch = CH; }
                        else if (rtype == 1) { src = reinterpret_cast<const char *>((tail ? p.posq : p.posqp) + term * LP); dst = ops + (size_t)OP_X * LP * 4; nch = CH; }
                        else if (rtype == 2) { src = reinterpret_cast<const char *>((tail ? p.rank_qp : p.rank_q) + term * LP); dst = ops + (size_t)OP_RC * LP * 4; nch = CB; }
                        else { src = reinterpret_cast<const char *>((tail ? p.cnt_q : p.cnt_qp) + term * LP); dst = ops + (size_t)OP_RC * LP * 4 + LP; nch = CB; }
                        for (int ch = half; ch < nch; ch += 2) cp_async16(dst + ch * 16, src + ch * 16);
                        if (rtype == 0 && half == 0) cp_async4(&s_op_alpha[issued & (SLOTS - 1)][sl], p.alpha + term);
                    }
                    issued++;
                    if (issued < n_nodes && ci.advance(segs)) my_own = ld_own(segs + ci.sg, sl);
                }
                cp_async_commit();
            };
            for (int q = 0; q < LAND - 1; q++) issue();
            for (int node = 0; node < n_nodes; node++) {
                mark(60, node);
                issue();
                cp_async_wait_group<LAND - 1>();
                __syncwarp();
                if (lane == 0) st_flag_cta(&s_flag[F_O], node + 1);
            }
            cp_async_wait_group<0>();
            mark(69, n_nodes);
        }
    }
    if (rec_buf && lane == 0) {
        volatile int4 *r = reinterpret_cast<volatile int4 *>(rec_buf) + (blockIdx.x * 8 + warp);
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
