// ncc_volume.cu -- the NCC cost volume of dispmap_ncc.compute_ncc (dispmap_ncc.m:116-198) for 8-bit images and
// integer disparity levels, as ONE pass over the images for ALL levels, and the device-resident volume handle
// (sb_ncc_vol_*) that dispmap_ncc's sampling methods consume (dispmap_ncc.m:208-276) without the volume ever
// becoming host doubles.
//
//   ncc(r, c, d) = real( (sRT - mR sT - mT sR + n mT mR) / sqrt(sRR - 2 mR sR + n mR^2) / sqrt(sTT - 2 mT sT + n mT^2) )
//   with the sums over the (2p+1)^2 x 3 window (zero padded, conv2 'same'), T = image 2 shifted by d, n = 3 (2p+1)^2.
//
// With integer pixel values every sum is an integer, and the expression equals
//   Cn / sqrt(A B),   Cn = n sRT - sR sT,   A = n sRR - sR^2,   B = n sTT - sT^2          (all exact in 64-bit integers)
// so nothing cancels in floating point: A and its rsqrt depend on the reference pixel only, B and its rsqrt on the
// pixel of image 2 only (both precomputed once per image pair), and per level only sRT = box(R . T_d) is left.
//
// Kernel: lane <-> image row (32 rows per CTA, up to 32 - 2p of them outputs), the CTA marches along the columns.
// The horizontal box is a running sum per lane (one dp4a in, one out per step), the vertical box a warp prefix
// scan; warp w owns the levels w, w + 8, ... with their running sums in registers.  Image 2's columns (packed RGB,
// its window sum and rsqrt(B)) sit in a shared-memory ring that TMA tensor-map copies (cp.async.bulk.tensor.2d,
// SASS UTMALDG; out-of-image rows / columns arrive as zeros = the zero padding of conv2) refill 32 columns at a
// time, so both images are read from HBM once for all levels and the only HBM stream is the volume write.
#include "sb_common.h"
#include <cuda.h>
#include <vector>
#include <cmath>
#include <algorithm>
#include <cstring>

namespace sb {
namespace dm {

// defined in dispmap_kernels.cu
int ncc_volume_general(int H, int W, const double *d_im0, const double *d_im1, int D, const double *h_disps, const double *d_disps,
                       int patchsize, bool exact32, float *vol);

namespace {

constexpr int NV_WARPS = 8;        // warps per CTA: the levels are dealt round robin
constexpr int NV_LPW = 16;         // levels per warp (=> at most 128 levels per launch)
constexpr int NV_SLOTS = 8;        // ring of 32-column blocks of image 2
constexpr int NV_BLK = 32;
constexpr int NV_PMAX = 8;

// H x W x 3 doubles (MATLAB layout) -> one 32-bit word per pixel (B G R 0), leading dimension Hp
__global__ void pack_rgb_kernel(const double *__restrict__ im, int H, int W, int Hp, unsigned *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Hp * W) return;
    const int r = (int)(i % Hp), c = (int)(i / Hp);
    unsigned v = 0;
    if (r < H) {
        const size_t u = (size_t)c * H + r, plane = (size_t)H * W;
        v = (unsigned)im[u] | ((unsigned)im[plane + u] << 8) | ((unsigned)im[2 * plane + u] << 16);
    }
    out[i] = v;
}

// window sums of one image (zero padded): s1 = sum of the channel values, and rsq = 1 / sqrt(n s2 - s1^2) (0 for a flat window)
__global__ void window_stats_kernel(const unsigned *__restrict__ pk, int H, int W, int Hp, int p, unsigned *__restrict__ s1out,
                                    float *__restrict__ rsq)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Hp * W) return;
    const int r = (int)(i % Hp), c = (int)(i / Hp);
    unsigned s1 = 0, s2 = 0;
    if (r < H) {
        for (int dc = -p; dc <= p; dc++) {
            const int cc = c + dc;
            if (cc < 0 || cc >= W) continue;
            for (int dr = -p; dr <= p; dr++) {
                const int rr = r + dr;
                if (rr < 0 || rr >= H) continue;
                const unsigned v = pk[(size_t)cc * Hp + rr];
                s1 = __dp4a(v, 0x00010101u, s1);
                s2 = __dp4a(v, v, s2);
            }
        }
    }
    const long long n3 = 3LL * (2 * p + 1) * (2 * p + 1);
    const long long A = n3 * (long long)s2 - (long long)s1 * (long long)s1;
    s1out[i] = s1;
    rsq[i] = A > 0 ? rsqrtf((float)A) : 0.f;
}

__device__ __forceinline__ unsigned s_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Sum of v over the lanes lane - P .. lane + P (lane ids modulo 32: exact for P <= lane < 32 - P).  Windows of
// length 2^k are doubled up by shuffles and the 2P + 1 window is assembled from them (4-6 shuffles instead of the 7 of
// a prefix scan); the source lanes are worked out once.
template <int P> struct VBoxLanes {
    int up[5];     // (lane + 2^b) & 31
    int at[5];     // where the piece of length 2^b of the window starts, as a lane id
    __device__ __forceinline__ explicit VBoxLanes(int lane)
    {
        int start = -P;
#pragma unroll
        for (int b = 0; b < 5; b++) {
            up[b] = (lane + (1 << b)) & 31;
            at[b] = (lane + start) & 31;
            if ((2 * P + 1) & (1 << b)) start += 1 << b;
        }
    }
    __device__ __forceinline__ int box(int v) const
    {
        constexpr unsigned FULL = 0xffffffffu;
        constexpr int N = 2 * P + 1;
        int acc = 0, w = v;
#pragma unroll
        for (int b = 0; b < 5; b++) {
            if (N & (1 << b)) acc += __shfl_sync(FULL, w, at[b]);
            if ((N >> (b + 1)) == 0) break;
            w += __shfl_sync(FULL, w, up[b]);
        }
        return acc;
    }
};

struct NccArgs {
    const unsigned *pk0;      // [W][Hp] packed reference image
    const unsigned *sR;       // window sum of the reference image
    const float *rsA;
    int H, W, Hp, D, P;
    int dmax;
    int cols_per_cta;
    int disps[NV_WARPS * NV_LPW];
    float *vol;               // [D][W][H]
};

template <int P>
__global__ void __launch_bounds__(NV_WARPS * 32, 2)
ncc_levels_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmS,
                  const __grid_constant__ CUtensorMap tmB, const NccArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned *ringT = reinterpret_cast<unsigned *>(smem_raw);                  // [slot][col 32][row 32]
    unsigned *ringS = ringT + NV_SLOTS * NV_BLK * 32;
    float *ringB = reinterpret_cast<float *>(ringS + NV_SLOTS * NV_BLK * 32);
    __shared__ __align__(8) unsigned long long bars[NV_SLOTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the first row of a CTA is a multiple of 4: the tensor-map copies want a 16-byte aligned start along the rows
    constexpr int WIN = 2 * P + 1, P4 = (P + 3) & ~3, OUT = ((32 - P4 - P) / 4) * 4;
    const int H = a.H, W = a.W, Hp = a.Hp;
    const int r0 = (int)blockIdx.x * OUT - P4, r = r0 + lane;
    const int c0 = (int)blockIdx.y * a.cols_per_cta, c1 = min(W, c0 + a.cols_per_cta);
    const bool row_in = r >= 0 && r < H;
    const bool out_lane = lane >= P4 && lane < P4 + OUT && r < H;
    const long long n3 = 3LL * WIN * WIN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NV_SLOTS; s++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bars + s)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // blocks of 32 columns of image 2 (absolute column x = c_in - d), block xb = x >> 5 lives in slot xb & 7
    const int cin_first = max(c0 - P, 0), cin_last = c1 - 1 + P;    // (columns left of the image add nothing)
    const int j0 = cin_first >> 5, j1 = cin_last >> 5;
    const int xb_first = max(0, cin_first - a.dmax) >> 5;
    // (the addresses of the __grid_constant__ maps are taken HERE: inside a capturing lambda the compiler may hand out
    // the address of a local copy, and a tensor map in local memory is an illegal instruction for the copy engine)
    const unsigned long long pT = reinterpret_cast<unsigned long long>(&tmT), pS = reinterpret_cast<unsigned long long>(&tmS),
                             pB = reinterpret_cast<unsigned long long>(&tmB);
    auto load_block = [&, pT, pS, pB](int xb) {
        const int s = xb & (NV_SLOTS - 1);
        const unsigned bar = s_u32(bars + s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3 * NV_BLK * 32 * 4) : "memory");
        const int x0 = r0, x1 = xb * NV_BLK;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s_u32(ringT + s * NV_BLK * 32)), "l"(pT), "r"(x0), "r"(x1), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s_u32(ringS + s * NV_BLK * 32)), "l"(pS), "r"(x0), "r"(x1), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s_u32(ringB + s * NV_BLK * 32)), "l"(pB), "r"(x0), "r"(x1), "r"(bar) : "memory");
    };
    auto wait_block = [&](int xb) {
        const int s = xb & (NV_SLOTS - 1);
        const unsigned parity = (unsigned)(((xb - xb_first) >> 3) & 1);
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "W_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
            "@p bra D_%=;\n\t"
            "bra W_%=;\n\t"
            "D_%=:\n\t}" ::"r"(s_u32(bars + s)), "r"(parity) : "memory");
    };
    if (threadIdx.x == 0)
        for (int xb = xb_first; xb <= j0; xb++) load_block(xb);
    for (int xb = xb_first; xb <= j0; xb++) wait_block(xb);

    // this warp's levels (slots beyond the level count run on d = 0 and never store: no divergent paths in the loops)
    int dk[NV_LPW];
    int hs[NV_LPW];
#pragma unroll
    for (int k = 0; k < NV_LPW; k++) {
        const int li = warp + NV_WARPS * k;
        dk[k] = li < a.D ? a.disps[li] : 0;
        hs[k] = 0;
    }
    const int nlev = a.D > warp ? (a.D - warp + NV_WARPS - 1) / NV_WARPS : 0;
    // ring word of absolute column x for this lane: the ring holds 8 x 32 = 256 columns of 32 rows
    auto ring_idx = [&](int x) { return ((x & (NV_SLOTS * NV_BLK - 1)) << 5) + lane; };
    auto load_R = [&](int cin, unsigned &Rn, unsigned &Ro, unsigned &srv, float &rsa) {
        Rn = 0; Ro = 0; srv = 0; rsa = 0.f;
        if (row_in) {
            if (cin >= 0 && cin < W) Rn = __ldg(a.pk0 + (size_t)cin * Hp + r);
            const int co = cin - WIN;
            if (co >= cin_first && co >= 0 && co < W) Ro = __ldg(a.pk0 + (size_t)co * Hp + r);
            const int c = cin - P;
            if (c >= 0 && c < W) { srv = __ldg(a.sR + (size_t)c * Hp + r); rsa = __ldg(a.rsA + (size_t)c * Hp + r); }
        }
    };
    VBoxLanes<P> vl(lane);
    float *const vol_lane = a.vol + (size_t)warp * W * H + r;       // level `warp`, column 0, this lane's row
    const size_t lvl_stride = (size_t)NV_WARPS * W * H;
    unsigned nRn, nRo, nsr;
    float nrsa;
    load_R(cin_first, nRn, nRo, nsr, nrsa);
    for (int j = j0; j <= j1; j++) {
        // the block that holds column c_in itself (d = 0) must have landed; older blocks were waited for before
        if (j > j0) wait_block(j);
        __syncthreads();      // every warp has left block j - 1: its oldest ring block may be overwritten
        if (threadIdx.x == 0 && j + 1 <= j1) load_block(j + 1);
        const int cb = max(cin_first, j * NV_BLK), ce = min(cin_last, j * NV_BLK + NV_BLK - 1);
        for (int cin = cb; cin <= ce; cin++) {
            const unsigned Rn = nRn, Ro = nRo, srv = nsr;     // (Rn = 0 right of the image, Ro = 0 during the warm-up)
            const float rsa = nrsa;
            load_R(cin + 1, nRn, nRo, nsr, nrsa);
            // ---- horizontal running sums: column c_in enters the window, column c_in - (2P + 1) leaves
#pragma unroll
            for (int k = 0; k < NV_LPW; k++) {
                const int x = cin - dk[k], xo = x - WIN;
                const unsigned Tn = x >= 0 ? ringT[ring_idx(x)] : 0u;
                const unsigned To = xo >= 0 ? ringT[ring_idx(xo)] : 0u;
                hs[k] += (int)__dp4a(Rn, Tn, 0u) - (int)__dp4a(Ro, To, 0u);
            }
            const int c = cin - P;                       // output column of this step
            if (c < c0 || c >= c1) continue;             // (warp-uniform)
            float *const vcol = vol_lane + (size_t)c * H;
            if (c + P <= W - 1) {
#pragma unroll
                for (int k = 0; k < NV_LPW; k++) {
                    // vertical box over the lanes lane - P .. lane + P (wrap-around only reaches lanes that emit nothing)
                    const long long sRT = (long long)vl.box(hs[k]);
                    const int xc = c - dk[k];
                    const unsigned sT = ringS[ring_idx(xc)];
                    const float rsb = xc >= 0 ? ringB[ring_idx(xc)] : 0.f;      // left of the shifted image: ncc = 0
                    const long long Cn = n3 * sRT - (long long)srv * (long long)sT;
                    const float v = (float)Cn * rsa * rsb;
                    if (out_lane && k < nlev) vcol[(size_t)k * lvl_stride] = v;
                }
            } else {
                // the window leaves the image on the right: the sums of the shifted image are truncated at its last
                // column (x = W - 1 - d), so they are taken from the ring directly
#pragma unroll
                for (int k = 0; k < NV_LPW; k++) {    // (unrolled: the running sums must stay in registers)
                    if (k >= nlev) continue;
                    const int d = dk[k];
                    const long long sRT = (long long)vl.box(hs[k]);
                    const int xc = c - d;
                    float v = 0.f;
                    if (xc >= 0) {
                        unsigned t1 = 0, t2 = 0;
                        for (int xx = max(xc - P, 0); xx <= W - 1 - d; xx++) {
                            const unsigned *col = ringT + ((xx & (NV_SLOTS * NV_BLK - 1)) << 5);
#pragma unroll
                            for (int dr = -P; dr <= P; dr++) {
                                const int l2 = lane + dr;
                                const unsigned tv = (l2 >= 0 && l2 < 32) ? col[l2] : 0u;
                                t1 = __dp4a(tv, 0x00010101u, t1);
                                t2 = __dp4a(tv, tv, t2);
                            }
                        }
                        const long long B = n3 * (long long)t2 - (long long)t1 * (long long)t1;
                        const float rsb = B > 0 ? rsqrtf((float)B) : 0.f;
                        const long long Cn = n3 * sRT - (long long)srv * (long long)t1;
                        v = (float)Cn * rsa * rsb;
                    }
                    if (out_lane) vcol[(size_t)k * lvl_stride] = v;
                }
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [W][Hp] array of 32-bit words, boxes of 32 rows x 32 columns
void make_map(CUtensorMap *m, void *base, int Hp, int W)
{
    EncodeTiledFn fn = encode_fn();
    SB_REQUIRE(fn, SB_ECUDA, "sb_ncc_volume: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)Hp, (cuuint64_t)W};
    const cuuint64_t gstride[1] = {(cuuint64_t)Hp * 4};
    const cuuint32_t box[2] = {32, NV_BLK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SB_REQUIRE(rc == CUDA_SUCCESS, SB_ECUDA, "sb_ncc_volume: cuTensorMapEncodeTiled failed (%d)", (int)rc);
}

template <int P> void launch_levels(dim3 grid, size_t smem, const CUtensorMap &t, const CUtensorMap &s, const CUtensorMap &b, const NccArgs &a)
{
    SB_CUDA(cudaFuncSetAttribute(ncc_levels_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ncc_levels_kernel<P><<<grid, NV_WARPS * 32, smem>>>(t, s, b, a);
}

} // namespace

// The volume of dispmap_ncc.compute_ncc as fp32 [D][W][H] on the device.  d_im0 / d_im1: H x W x 3 doubles on the device.
// Returns the milliseconds the volume kernels took (CUDA events); *fast tells which path ran.
double compute_ncc_volume(int H, int W, const double *d_im0, const double *d_im1, const double *h_im0, const double *h_im1, int D,
                          const double *h_disps, const double *d_disps, int patchsize, float *vol, bool *fast)
{
    const long long N = (long long)H * W;
    // exact path: 8-bit integer images and integer disparities
    bool exact = true;
    int dmax = 0;
    for (int i = 0; i < D && exact; i++) {
        exact = h_disps[i] == std::floor(h_disps[i]) && h_disps[i] >= 0 && h_disps[i] < (double)W;
        if (exact) dmax = std::max(dmax, (int)h_disps[i]);
    }
    for (long long i = 0; i < N * 3 && exact; i++)
        exact = h_im0[i] >= 0 && h_im0[i] <= 255 && h_im0[i] == std::floor(h_im0[i]) && h_im1[i] >= 0 && h_im1[i] <= 255 &&
                h_im1[i] == std::floor(h_im1[i]);
    const bool exact32 = exact && 3.0 * (2 * patchsize + 1) * (2 * patchsize + 1) * 65025.0 < 16777216.0;
    const bool use_fast = exact && !getenv("SB_NCC_GENERAL") && patchsize >= 1 && patchsize <= NV_PMAX && H >= 1 &&
                          dmax + 2 * patchsize + 1 <= 5 * NV_BLK && encode_fn() != nullptr;
    if (fast) *fast = use_fast;
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0));
    SB_CUDA(cudaEventCreate(&e1));
    SB_CUDA(cudaEventRecord(e0, 0));
    if (!use_fast) {
        ncc_volume_general(H, W, d_im0, d_im1, D, h_disps, d_disps, patchsize, exact32, vol);
    } else {
        const int Hp = (H + 3) & ~3;      // 16-byte row pitch for the tensor maps
        const size_t np = (size_t)Hp * W;
        DevBuf<unsigned> pk0(np), pk1(np), sR(np), sT(np);
        DevBuf<float> rsA(np), rsB(np);
        const unsigned nb = (unsigned)((np + 255) / 256);
        pack_rgb_kernel<<<nb, 256>>>(d_im0, H, W, Hp, pk0.p);
        pack_rgb_kernel<<<nb, 256>>>(d_im1, H, W, Hp, pk1.p);
        window_stats_kernel<<<nb, 256>>>(pk0.p, H, W, Hp, patchsize, sR.p, rsA.p);
        window_stats_kernel<<<nb, 256>>>(pk1.p, H, W, Hp, patchsize, sT.p, rsB.p);
        SB_CUDA(cudaGetLastError());
        count_launch(4);
        CUtensorMap tmT, tmS, tmB;
        make_map(&tmT, pk1.p, Hp, W);
        make_map(&tmS, sT.p, Hp, W);
        make_map(&tmB, rsB.p, Hp, W);
        const int P = patchsize, P4 = (P + 3) & ~3, OUT = ((32 - P4 - P) / 4) * 4;
        const int groups = (H + OUT - 1) / OUT;
        // ONE wave of CTAs, two per SM (a second, mostly empty wave would double the time); a column chunk re-reads
        // 2p + 1 + dmax columns of warm-up
        int num_sms = 148, dev = 0;
        SB_CUDA(cudaGetDevice(&dev));
        SB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        int chunks = std::max(1, (2 * num_sms) / groups);
        int cpc = std::max(64, (W + chunks - 1) / chunks);
        cpc = (cpc + 31) & ~31;
        chunks = (W + cpc - 1) / cpc;
        const size_t smem = (size_t)3 * NV_SLOTS * NV_BLK * 32 * 4;
        for (int l0 = 0; l0 < D; l0 += NV_WARPS * NV_LPW) {
            NccArgs a;
            std::memset(&a, 0, sizeof(a));
            a.pk0 = pk0.p; a.sR = sR.p; a.rsA = rsA.p; a.H = H; a.W = W; a.Hp = Hp; a.P = P;
            a.D = std::min(NV_WARPS * NV_LPW, D - l0);
            a.dmax = 0;
            for (int i = 0; i < a.D; i++) { a.disps[i] = (int)h_disps[l0 + i]; a.dmax = std::max(a.dmax, a.disps[i]); }
            a.cols_per_cta = cpc;
            a.vol = vol + (size_t)l0 * N;
            dim3 grid(groups, chunks);
            switch (P) {
            case 1: launch_levels<1>(grid, smem, tmT, tmS, tmB, a); break;
            case 2: launch_levels<2>(grid, smem, tmT, tmS, tmB, a); break;
            case 3: launch_levels<3>(grid, smem, tmT, tmS, tmB, a); break;
            case 4: launch_levels<4>(grid, smem, tmT, tmS, tmB, a); break;
            case 5: launch_levels<5>(grid, smem, tmT, tmS, tmB, a); break;
            case 6: launch_levels<6>(grid, smem, tmT, tmS, tmB, a); break;
            case 7: launch_levels<7>(grid, smem, tmT, tmS, tmB, a); break;
            default: launch_levels<8>(grid, smem, tmT, tmS, tmB, a); break;
            }
            SB_CUDA(cudaGetLastError());
            count_launch();
        }
        SB_CUDA(cudaDeviceSynchronize());   // the temporaries above are released at scope exit
    }
    SB_CUDA(cudaEventRecord(e1, 0));
    SB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

} // namespace dm
} // namespace sb
