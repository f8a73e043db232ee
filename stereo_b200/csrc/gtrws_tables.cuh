// gtrws_tables.cuh -- set-up kernels of the grid-native TRW-S path: sort ranks and merge counts
// of the label positions, built on the device from the per-node plane rows.  They are the
// counterpart of the per-edge argsort loop of cpp/trws_mex.cpp:84-119 (the reference sorts q(:,p)
// and qprim(:,p) of every term with std::sort inside the j loop); here
//   q(:, p)     = own[head]                         dispmap_super.m:181
//   qprim(:, p) = own[tail] +- g[tail]              dispmap_super.m:182  (tail planes at the head's point)
// are recomputed with the SAME fp32 expressions the sweep kernel uses, so the tables always agree
// with the positions the sweep sees.
#pragma once
#include "gtrws_kernels.cuh"

namespace sb {
namespace gtrws {

// labels L..LP-1 of every node: D = 0, slopes 0, own = a value beyond every position the node's real
// labels can take (own +- gx, own +- gy), so that the padding sorts LAST in every sorted array of this
// node -- sorted indices 0..L-1 are exactly the real labels, as the update kernels assume
template <typename REAL, int K>
__global__ void gpad_kernel(REAL *nodeF, long long Nloc, int L)
{
    constexpr int LP = 32 * K;
    const int lane = threadIdx.x & 31;
    const long long v = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (v >= Nloc || L >= LP) return;
    REAL *rec = nodeF + v * 4 * LP;
    REAL mx = -Lim<REAL>::big();
    for (int l = lane; l < L; l += 32)
        mx = max(mx, rec[NF_OWN * LP + l] + fabs(rec[NF_GX * LP + l]) + fabs(rec[NF_GY * LP + l]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mx = mx + REAL(1) + fabs(mx) * REAL(1e-3);
    for (int l = L + lane; l < LP; l += 32) {
        rec[NF_D * LP + l] = REAL(0);
        rec[NF_GX * LP + l] = REAL(0);
        rec[NF_OWN * LP + l] = mx;
        rec[NF_GY * LP + l] = REAL(0);
    }
}

// rank of every label of A among the LP entries of A (ties: lower label first), and
// cnt[l] = #{m : B[m] <= A[l]} clamped to 255
template <typename REAL, int K>
__device__ __forceinline__ void rank_rows(const REAL *A, int lane, uint8_t (&rk)[K])
{
    constexpr int LP = 32 * K;
    REAL a[K];
    int r[K];
#pragma unroll
    for (int k = 0; k < K; k++) { a[k] = A[lane * K + k]; r[k] = 0; }
    for (int m = 0; m < LP; m++) {
        const REAL v = A[m];
#pragma unroll
        for (int k = 0; k < K; k++) r[k] += (v < a[k]) || (v == a[k] && m < lane * K + k);
    }
#pragma unroll
    for (int k = 0; k < K; k++) rk[k] = (uint8_t)r[k];
}
template <typename REAL, int K>
__device__ __forceinline__ void count_rows(const REAL *A, const REAL *B, int lane, uint8_t (&cn)[K])
{
    constexpr int LP = 32 * K;
    REAL a[K];
    int c[K];
#pragma unroll
    for (int k = 0; k < K; k++) { a[k] = A[lane * K + k]; c[k] = 0; }
    for (int m = 0; m < LP; m++) {
        const REAL v = B[m];
#pragma unroll
        for (int k = 0; k < K; k++) c[k] += (v <= a[k]);
    }
#pragma unroll
    for (int k = 0; k < K; k++) cn[k] = (uint8_t)min(c[k], 255);
}

// ---------------------------------------------------------------- sort-based tables (one warp per row)
// The counting loops above are O(LP^2) per row; a grid of 1980 x 2880 nodes with 192 labels has 68 million rows.
// A stable bitonic sort of (position, label) in shared memory gives the ranks in O(LP log^2 LP), and the merge
// counts are upper bounds in the sorted rows (binary search): same integers, ~6x fewer instructions.
template <int LP> __host__ __device__ constexpr int pow2_at_least() { int n = 32; while (n < LP) n <<= 1; return n; }

// keys[0..LP) in, LPS - LP padding slots appended (+big); out: keys sorted ascending (ties: lower label first),
// idx[pos] = label at sorted position pos.  One warp; keys / idx hold LPS entries.
template <typename REAL, int LP, int LPS>
__device__ __forceinline__ void warp_sort(REAL *keys, int *idx, int lane)
{
    for (int i = lane; i < LPS; i += 32) {
        idx[i] = i;
        if (i >= LP) keys[i] = Lim<REAL>::big();
    }
    __syncwarp();
    for (int k = 2; k <= LPS; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < LPS / 2; t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // t with a zero bit inserted at position log2(j)
                const int p = i | j;
                const REAL a = keys[i], b = keys[p];
                const int ia = idx[i], ib = idx[p];
                const bool up = (i & k) == 0;
                const bool a_after_b = (b < a) || (b == a && ib < ia);
                if (a_after_b == up) {
                    keys[i] = b; keys[p] = a;
                    idx[i] = ib; idx[p] = ia;
                }
            }
            __syncwarp();
        }
}
// #{m < LP : S[m] <= v} for S sorted ascending
template <typename REAL, int LP>
__device__ __forceinline__ int upper_bound(const REAL *S, REAL v)
{
    int lo = 0, hi = LP;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (S[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one warp per node: rank of own
template <typename REAL, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gnode_tables_kernel(const REAL *__restrict__ nodeF, uint8_t *__restrict__ nodeB, long long Nloc)
{
    constexpr int LP = 32 * K, LPS = pow2_at_least<LP>();
    __shared__ REAL keys[WARPS][LPS];
    __shared__ int idx[WARPS][LPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long v = (long long)blockIdx.x * WARPS + warp; v < Nloc; v += (long long)gridDim.x * WARPS) {
        const REAL *rec = nodeF + v * 4 * LP;
        for (int l = lane; l < LP; l += 32) keys[warp][l] = rec[NF_OWN * LP + l];
        __syncwarp();
        warp_sort<REAL, LP, LPS>(keys[warp], idx[warp], lane);
        for (int pos = lane; pos < LP; pos += 32) nodeB[v * LP + idx[warp][pos]] = (uint8_t)pos;
        __syncwarp();
    }
}

// one warp per neighbour pair (first node a = pair / 2, direction down / right = pair % 2):
//   term 0: tail a, head b:  q0 = own_b,  qp0 = own_a + g_a
//   term 1: tail b, head a:  q1 = own_a,  qp1 = own_b - g_b
// side record s (what node s of the pair needs when it sends): [cnt_q[s] | cnt_qp[1-s] | rank_tail[s]]
//   cnt_q[j][l]  = #{m : qp_j[m] <= q_j[l]}     cnt_qp[j][l] = #{m : q_j[m] <= qp_j[l]}
// Runs after gnode_tables_kernel: the sorted q rows come from the node ranks.
template <typename REAL, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gpair_tables_kernel(const REAL *__restrict__ nodeF, const uint8_t *__restrict__ nodeB,
                                                                  uint8_t *__restrict__ pairB, int W, int rows)
{
    constexpr int LP = 32 * K, LPS = pow2_at_least<LP>();
    __shared__ REAL q[WARPS][2][LP];       // q_j by label
    __shared__ REAL qp[WARPS][2][LP];      // qp_j by label
    __shared__ REAL sq[WARPS][2][LP];      // q_j sorted
    __shared__ REAL sqp[WARPS][LPS];       // qp_j sorted (one term at a time)
    __shared__ int idx[WARPS][LPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long Nloc = (long long)rows * W;
    for (long long pr = (long long)blockIdx.x * WARPS + warp; pr < 2 * Nloc; pr += (long long)gridDim.x * WARPS) {
        const long long a = pr >> 1;
        const int dirn = (int)(pr & 1);   // 0 down, 1 right
        const int r = (int)(a / W), c = (int)(a % W);
        if (dirn == 0 ? (r + 1 >= rows) : (c + 1 >= W)) continue;
        const long long b = dirn == 0 ? a + W : a + 1;
        const REAL *ra = nodeF + a * 4 * LP, *rb = nodeF + b * 4 * LP;
        const uint8_t *ka = nodeB + a * LP, *kb = nodeB + b * LP;
        const int gsel = dirn == 0 ? NF_GY : NF_GX;
        for (int l = lane; l < LP; l += 32) {
            const REAL oa = ra[NF_OWN * LP + l], ob = rb[NF_OWN * LP + l];
            q[warp][0][l] = ob;
            qp[warp][0][l] = oa + ra[gsel * LP + l];
            q[warp][1][l] = oa;
            qp[warp][1][l] = ob - rb[gsel * LP + l];
            sq[warp][0][kb[l]] = ob;      // rank of own: a permutation of 0 .. LP-1
            sq[warp][1][ka[l]] = oa;
        }
        __syncwarp();
        uint8_t *rec = pairB + pr * 6 * LP;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            // record layout: term 0 -> rows 0 (cnt_q), 4 (cnt_qp), 2 (rank_tail); term 1 -> rows 3, 1, 5
            const int row_cq = j == 0 ? 0 : 3, row_cqp = j == 0 ? 4 : 1, row_rk = j == 0 ? 2 : 5;
            for (int l = lane; l < LP; l += 32) sqp[warp][l] = qp[warp][j][l];
            __syncwarp();
            warp_sort<REAL, LP, LPS>(sqp[warp], idx[warp], lane);
            for (int pos = lane; pos < LP; pos += 32) rec[row_rk * LP + idx[warp][pos]] = (uint8_t)pos;
            for (int l = lane; l < LP; l += 32) {
                rec[row_cq * LP + l] = (uint8_t)min(upper_bound<REAL, LP>(sqp[warp], q[warp][j][l]), 255);
                rec[row_cqp * LP + l] = (uint8_t)min(upper_bound<REAL, LP>(sq[warp][j], qp[warp][j][l]), 255);
            }
            __syncwarp();
        }
    }
}

// One Edge::UpdateMessage on the device (unit-level known-answer tests; sb_trws_update_message): one warp,
// source positions src (cones rooted there), destination positions dst, Di = the gamma-unscaled node sum,
// msg in / out.  Ranks and merge counts are built the way the table kernels build them.
template <typename REAL, int K, int KERN>
__global__ void gupdate_message_kernel(int L, const double *__restrict__ Di_in, const double *__restrict__ msg_in,
                                       const double *__restrict__ src, const double *__restrict__ dst, double alpha, double lambda,
                                       double gamma, double *__restrict__ msg_out, double *__restrict__ vmin_out)
{
    constexpr int LP = 32 * K;
    __shared__ REAL sh[2][LP];
    __shared__ trws::Pair<REAL> P[trws::scratch_pairs<K>()];
    const int lane = threadIdx.x;
    REAL Di[K], m[K], s[K], x[K];
    REAL smax = -Lim<REAL>::big(), xmax = -Lim<REAL>::big();
    for (int l = 0; l < L; l++) { smax = max(smax, (REAL)src[l]); xmax = max(xmax, (REAL)dst[l]); }
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int l = lane * K + k;
        Di[k] = l < L ? (REAL)Di_in[l] : REAL(0);
        m[k] = l < L ? (REAL)msg_in[l] : REAL(0);
        s[k] = l < L ? (REAL)src[l] : smax + REAL(1);
        x[k] = l < L ? (REAL)dst[l] : xmax + REAL(1);
        sh[0][l] = s[k];
        sh[1][l] = x[k];
    }
    if (lane == 0) {
        trws::Pair<REAL> t;
        t.a = Lim<REAL>::big();
        t.b = REAL(0);
        P[0] = t;
        P[trws::phys<K>(LP)] = t;
    }
    __syncwarp();
    uint8_t rk[K], cn[K];
    rank_rows<REAL, K>(sh[0], lane, rk);
    count_rows<REAL, K>(sh[1], sh[0], lane, cn);   // #{src <= dst[l]}
    REAL vmin;
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < K; k++) valid |= (lane * K + k < L ? 1u : 0u) << k;
    if constexpr (KERN == 1) vmin = trws::update_linear<REAL, K>((REAL)gamma, (REAL)alpha, (REAL)lambda, valid, lane, Di, m, s, rk, x, cn, P);
    else vmin = trws::update_quadratic<REAL, K>((REAL)gamma, (REAL)alpha, (REAL)lambda, valid, L, lane, Di, m, s, rk, x, cn, P);
#pragma unroll
    for (int k = 0; k < K; k++)
        if (lane * K + k < L) msg_out[lane * K + k] = (double)m[k];
    if (lane == 0) *vmin_out = (double)vmin;
}

} // namespace gtrws
} // namespace sb
