// gtrws_solve.cu -- host driver of the grid-native TRW-S path (sb_trws_grid_*,
// include/stereo_b200.h).  Same solver semantics as solve_mrf() of cpp/trws_mex.cpp:27-147
// (order the nodes, iterate forward / backward sweeps under the stopping rule of
// cpp/trw-s/minimize.cpp:97-112, return labels / energy / bound / iterations), but the problem
// enters the way dispmap_super.simultaneous_fusion HOLDS it -- L plane proposals per pixel, a unary
// slab per proposal and one weight per term (dispmap_super.m:158-188) -- instead of the L x E q / qprim
// arrays it derives from them, and lives in the 45-bytes-per-label-and-node layout of
// gtrws_kernels.cuh, row-banded over the GPUs of a box.
#include "sb_common.h"
#include "trws_order.h"
#include "gtrws_plan.h"
#include "gtrws_kernels.cuh"
#include "gtrws_launch.h"
#include <vector>
#include <map>
#include <tuple>
#include <memory>
#include <mutex>
#include <chrono>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <unistd.h>

namespace sb {
namespace gtrws {

const GOps *gops_for_labels(int L)
{
    static const GOps *table[] = {&gops_k1, &gops_k2, &gops_k3, &gops_k4, &gops_k6, &gops_k8};
    for (const GOps *k : table)
        if (32 * k->K >= L) return k;
    return nullptr;
}

namespace {

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

struct Ctrl {
    int ticket;
    int ring_smid;
    double acc[2];
};

// Plain cudaMalloc buffers: the big slabs are allocated once per solver, and the message / selected-
// position arrays must be exportable through CUDA IPC.
template <typename T> struct RawBuf {
    T *p = nullptr;
    size_t n = 0;
    RawBuf() {}
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    ~RawBuf() { if (p) cudaFree(p); }
    void alloc(size_t count)
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = count;
        if (count) SB_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    size_t bytes() const { return n * sizeof(T); }
};

struct GridPlanDev {
    int S[2] = {0, 0};
    int save_slots = 0;
    DevBuf<GSeg> segs[2];
    DevBuf<int32_t> seg_ptr[2], strip_len[2], is_ring[2];
};

std::shared_ptr<GridPlanDev> grid_plan(int dev, int H, int W, int rank, int world, int blocks)
{
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int, int, int>, std::shared_ptr<GridPlanDev>> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple(dev, H, W, rank, world, blocks);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    auto gp = std::make_shared<GridPlanDev>();
    for (int pass = 0; pass < 2; pass++) {
        GPassPlan plan;
        build_gpass_plan(H, W, pass, world > 1 ? rank : -1, world, blocks, plan);
        gp->S[pass] = (int)plan.strip_len.size();
        gp->save_slots = std::max(gp->save_slots, (int)plan.save_slots);
        gp->segs[pass].alloc(std::max<size_t>(plan.segs.size(), 1));
        gp->seg_ptr[pass].alloc(plan.seg_ptr.size());
        gp->strip_len[pass].alloc(std::max<size_t>(plan.strip_len.size(), 1));
        gp->is_ring[pass].alloc(std::max<size_t>(plan.is_ring.size(), 1));
        SB_CUDA(cudaMemcpy(gp->is_ring[pass].p, plan.is_ring.data(), plan.is_ring.size() * 4, cudaMemcpyHostToDevice));
        SB_CUDA(cudaMemcpy(gp->segs[pass].p, plan.segs.data(), plan.segs.size() * sizeof(GSeg), cudaMemcpyHostToDevice));
        SB_CUDA(cudaMemcpy(gp->seg_ptr[pass].p, plan.seg_ptr.data(), plan.seg_ptr.size() * 4, cudaMemcpyHostToDevice));
        SB_CUDA(cudaMemcpy(gp->strip_len[pass].p, plan.strip_len.data(), plan.strip_len.size() * 4, cudaMemcpyHostToDevice));
    }
    if (cache.size() >= 8) cache.clear();
    cache[key] = gp;
    return gp;
}

// ---------------------------------------------------------------- set-up kernels (label-count agnostic)

// planes: nl proposals, each 4 x N doubles ([a; b; c; d0] per MATLAB node u = r + H c); unary nl x N.
// own = (-(a x + b y + d0) / c - d_min) / d_step at the node's own point (x, y) = (c + 1, r + 1)
// (dispmap_super.m:318-328, dispmap_globalstereo.m:336-345), gx / gy = its change per column / row.
template <typename REAL>
__global__ void gfill_labels_kernel(const double *__restrict__ planes, const double *__restrict__ unary, int nl, int l0,
                                    int H, int W, int r_base, int rows, int c_base, int Wl, int LP, double d_min, double d_step,
                                    REAL *__restrict__ nodeF, int *bad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long Nloc = (long long)rows * Wl;
    if (t >= Nloc * nl) return;
    const int ll = (int)(t % nl);
    const long long v = t / nl;
    const int r = r_base + (int)(v / Wl), c = c_base + (int)(v % Wl);
    if (c < 0 || c >= W) return;      // unused halo column at the image border
    const long long u = r + (long long)H * c;
    const long long N = (long long)H * W;
    const double *pl = planes + ((long long)ll * N + u) * 4;
    const double a = pl[0], b = pl[1], cc = pl[2], d0 = pl[3];
    const double un = unary[(long long)ll * N + u];
    if (cc == 0.0) atomicOr(bad, 1);
    const double own = (-(a * (double)(c + 1) + b * (double)(r + 1) + d0) / cc - d_min) / d_step;
    const double gx = -a / cc / d_step, gy = -b / cc / d_step;
    if (!(own == own) || !(gx == gx) || !(gy == gy)) atomicOr(bad, 2);
    if (!(un == un)) atomicOr(bad, 4);
    REAL *rec = nodeF + v * 4 * LP + (l0 + ll);
    rec[NF_D * LP] = (REAL)un;
    rec[NF_GX * LP] = (REAL)gx;
    rec[NF_OWN * LP] = (REAL)own;
    rec[NF_GY * LP] = (REAL)gy;
}

// alphas in the reference's term order (dispmap_super.m:284-294: vertical down, vertical up, horizontal
// right, horizontal left, each column-major over the start node) -> alpha[pair][j]
template <typename REAL>
__global__ void gweights_kernel(const double *__restrict__ alphas, int H, int W, int r_base, int rows, int c_base, int Wl,
                                REAL *__restrict__ alpha)
{
    const long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long Nloc = (long long)rows * Wl;
    if (pr >= 2 * Nloc) return;
    const long long v = pr >> 1;
    const int dirn = (int)(pr & 1);
    const int r = r_base + (int)(v / Wl), c = c_base + (int)(v % Wl);
    const long long nV = (long long)(H - 1) * W, nH = (long long)H * (W - 1);
    REAL a0 = REAL(0), a1 = REAL(0);
    if (c < 0 || c >= W) {
        // unused halo column at the image border
    } else if (dirn == 0) {
        if (r + 1 < H) {
            const long long e = (long long)c * (H - 1) + r;
            a0 = (REAL)alphas[e];
            a1 = (REAL)alphas[nV + e];
        }
    } else if (c + 1 < W) {
        const long long e = 2 * nV + (long long)c * H + r;
        a0 = (REAL)alphas[e];
        a1 = (REAL)alphas[nH + e];
    }
    alpha[pr * 2] = a0;
    alpha[pr * 2 + 1] = a1;
}

// rounded labels of the rows this rank sweeps -> doubles, 1-based, MATLAB node order (trws_mex.cpp:134-139)
// (one column block: owned columns [c_lo, c_hi), stored from column c_base with Wl columns per row)
__global__ void glabels_kernel(const int32_t *__restrict__ sol, int H, int c_lo, int c_hi, int c_base, int Wl, double *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int wo = c_hi - c_lo;
    const long long n = (long long)H * wo;
    if (t >= n) return;
    const int r = (int)(t / wo), c = c_lo + (int)(t % wo);
    out[r + (long long)H * c] = (double)(sol[(long long)r * Wl + (c - c_base)] + 1);
}

// one stored label plane back out (inspection / parity tests): 4 x N doubles, rows not stored -> 0
template <typename REAL>
__global__ void gget_label_kernel(const REAL *__restrict__ nodeF, int l, int H, int W, int r_base, int rows, int c_base, int Wl, int LP,
                                  double *__restrict__ out)
{
    const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= (long long)rows * Wl) return;
    const int r = r_base + (int)(v / Wl), c = c_base + (int)(v % Wl);
    if (c < 0 || c >= W) return;
    const long long u = r + (long long)H * c, N = (long long)H * W;
    const REAL *rec = nodeF + v * 4 * LP + l;
    out[u] = (double)rec[NF_D * LP];
    out[N + u] = (double)rec[NF_OWN * LP];
    out[2 * N + u] = (double)rec[NF_GX * LP];
    out[3 * N + u] = (double)rec[NF_GY * LP];
}

template <typename REAL>
__global__ void gget_weights_kernel(const REAL *__restrict__ alpha, int H, int W, int r_base, int rows, int c_base, int Wl,
                                    double *__restrict__ out)
{
    const long long pr = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pr >= 2LL * rows * Wl) return;
    const long long v = pr >> 1;
    const int dirn = (int)(pr & 1);
    const int r = r_base + (int)(v / Wl), c = c_base + (int)(v % Wl);
    const long long nV = (long long)(H - 1) * W, nH = (long long)H * (W - 1);
    if (c < 0 || c >= W) return;
    if (dirn == 0) {
        if (r + 1 < r_base + rows) {
            const long long e = (long long)c * (H - 1) + r;
            out[e] = (double)alpha[pr * 2];
            out[nV + e] = (double)alpha[pr * 2 + 1];
        }
    } else if (c + 1 < c_base + Wl && c + 1 < W) {
        const long long e = 2 * nV + (long long)c * H + r;
        out[e] = (double)alpha[pr * 2];
        out[nH + e] = (double)alpha[pr * 2 + 1];
    }
}

// Seeded synthetic problem generated in place (bench.py at sizes whose inputs do not fit a host):
// label l of every pixel is a plane of a piecewise-planar field over a rectangular segmentation
// (cell size and plane parameters hashed from (seed, l, cell)), every fourth label fronto-parallel,
// the last label a per-pixel mix ("current assignment", dispmap_super.m:158); unary ~ U(0, log 2);
// weights 2 * {108, 9} with probability {0.8, 0.2} per neighbour pair (dispmap_globalstereo.m:400-403).
__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ unsigned hash3(unsigned seed, unsigned a, unsigned b, unsigned c)
{
    return hash32(hash32(hash32(seed ^ (a * 0x9e3779b9u)) ^ (b * 0x85ebca6bu)) ^ (c * 0xc2b2ae35u));
}
__device__ __forceinline__ float u01(unsigned h) { return (float)(h >> 8) * (1.0f / 16777216.0f); }

template <typename REAL>
__global__ void gsynth_kernel(unsigned seed, int kern, int Hs, int Ws, int r_off, int c_off, int Wloc, int Wgrid, int L, int r_base, int rows, int c_base, int LP,
                              REAL *__restrict__ nodeF, REAL *__restrict__ alpha)
{
    // (Hs, Ws): the grid the synthetic scene is defined on; the solver's own grid is the window of it that
    // starts at (r_off, c_off) (the whole of it for the real run, a crop for the CPU-baseline sample)
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long Nloc = (long long)rows * Wloc;
    if (t >= Nloc * L) return;
    const int l = (int)(t % L);
    const long long v = t / L;
    const int cw = c_base + (int)(v % Wloc);     // column within the solver's own grid
    if (cw < 0 || cw >= Wgrid) return;           // unused halo column at the image border
    const int r = r_off + r_base + (int)(v / Wloc), c = c_off + cw;
    const int H = Hs, W = Ws;
    auto plane_of = [&](int lab, float &own, float &gx, float &gy) {
        if (lab % 4 == 0) {
            own = ((float)lab + 0.5f) / (float)L; gx = 0.f; gy = 0.f;
            return;
        }
        const unsigned hs = hash3(seed, 0x51u, (unsigned)lab, 0u);
        const int cw = 16 + (int)(hs & 127u), ch = 16 + (int)((hs >> 8) & 127u);   // cell size of this proposal
        const int cx = c / cw, cy = r / ch;
        const unsigned hc = hash3(seed, 0x52u + (unsigned)lab, (unsigned)cx, (unsigned)cy);
        gx = (u01(hash32(hc ^ 1u)) - 0.5f) * 0.16f / (float)W;
        gy = (u01(hash32(hc ^ 2u)) - 0.5f) * 0.16f / (float)H;
        const float d = u01(hash32(hc ^ 3u));
        own = d + gx * ((float)c - ((float)cx + 0.5f) * (float)cw) + gy * ((float)r - ((float)cy + 0.5f) * (float)ch);
    };
    float own, gx, gy;
    if (l == L - 1 && L > 2) {
        const int pick = (int)(hash3(seed, 0x53u, (unsigned)r, (unsigned)c) % (unsigned)(L - 1));
        plane_of(pick, own, gx, gy);
    } else {
        plane_of(l, own, gx, gy);
    }
    REAL *rec = nodeF + v * 4 * LP + l;
    rec[NF_D * LP] = (REAL)(u01(hash3(seed, 0x54u + (unsigned)l, (unsigned)r, (unsigned)c)) * 0.69314718f);
    rec[NF_GX * LP] = (REAL)gx;
    rec[NF_OWN * LP] = (REAL)own;
    rec[NF_GY * LP] = (REAL)gy;
    if (l < 2) {
        // l = 0: pair (v, down), l = 1: pair (v, right); both terms of a pair share the weight
        const bool exists = true;   // pairs beyond the solver's own grid are never read
        float w = (u01(hash3(seed, 0x55u + (unsigned)l, (unsigned)r, (unsigned)c)) < 0.8f ? 108.f : 9.f) * 2.f;
        if (kern == 2) w = w / 0.02f;
        if (!exists) w = 0.f;
        alpha[(v * 2 + l) * 2] = (REAL)w;
        alpha[(v * 2 + l) * 2 + 1] = (REAL)w;
    }
}

struct SolverBase {
    virtual ~SolverBase() {}
    virtual void set_labels(int l0, int nl, const double *planes, const double *unary, double d_min, double d_step) = 0;
    virtual void set_weights(const double *alphas) = 0;
    virtual void synth(uint64_t seed, int Hs, int Ws, int r_off, int c_off) = 0;
    virtual void finalize() = 0;
    virtual void get_label(int l, double *out) = 0;
    virtual void get_weights(double *out) = 0;
    virtual void reset() = 0;
    virtual void minimize(double maxiter, double max_relgap, double *energy, double *lb, double *iters, sb_trws_timing *timing) = 0;
    virtual void labels(double *out) = 0;
    virtual void ipc_export(unsigned char *out) = 0;
    virtual void ipc_attach(const unsigned char *up, const unsigned char *down) = 0;
    virtual void run_one_pass(int pass, int mode, double *acc) = 0;
    virtual void launch_one_pass(int pass, int mode) = 0;
    virtual int wait_all(double *acc, int max_passes) = 0;
    virtual void *raw_ptr(int which) = 0;
    virtual void attach_local(SolverBase *up, SolverBase *down, int share) = 0;
    virtual void info(int64_t *out) = 0;
    virtual int latency_build() const = 0;
    virtual void counters(double *out) = 0;
    double setup_ms = 0;
};

template <typename REAL>
struct Solver : SolverBase {
    int kernel, H, W, L, K, LP, precision, rank, world;
    Band band;
    int rows, Wl;      // rows and columns stored per column block of this rank
    int blocks = 1;    // column blocks per rank (block-cyclic bands)
    int64_t blk_nodes = 0;
    int64_t Nloc, N, E;
    bool fuse, finalized = false;
    const GOps *ops;
    cudaStream_t stream = 0;
    RawBuf<REAL> dNodeF, dMsg, dAlpha;
    RawBuf<uint8_t> dNodeB, dPairB;
    RawBuf<unsigned long long> dSelBox, dSave;
    RawBuf<int32_t> dSol;
    RawBuf<unsigned char> dCtrl;
    RawBuf<long long> dProf;
    int *rec_host = nullptr;   // SB_TRWS_RECORD flight recorder (host-mapped)
    int rec_ctas = 0;
    double watchdog_ms = 30000.0;
    std::shared_ptr<GridPlanDev> plan;
    void *peer_ptr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    GProblem<REAL> P;
    int grid_fwd = 1, grid_bwd = 1;
    int lat = 0;               // 1: latency build of the sweep (at most two CTAs per SM, no spills, early operands)
    bool isolate = false;
    unsigned launch_epoch = 0, pass_counter = 0;
    Ctrl *hc = nullptr;   // pinned
    double kernel_ms = 0, total_kernel_ms = 0;
    int64_t kernel_count = 0, total_kernel_count = 0;

    Solver(int kernel_, int H_, int W_, int L_, double tol, const sb_trws_options &opt, int rank_, int world_)
        : kernel(kernel_), H(H_), W(W_), L(L_), rank(rank_), world(world_)
    {
        const double t0 = now_ms();
        precision = sizeof(REAL) == 8 ? SB_F64 : SB_F32;
        fuse = opt.fuse_rounding != 0;
        ops = gops_for_labels(L);
        SB_REQUIRE(ops, SB_EUNSUP, "sb_trws_grid: %d labels exceed SB_MAX_LABELS=%d", L, SB_MAX_LABELS);
        K = ops->K;
        LP = 32 * K;
        N = (int64_t)H * W;
        E = 2 * ((int64_t)(H - 1) * W + (int64_t)H * (W - 1));
        SB_REQUIRE(world == 1 || world <= W / 4, SB_EUNSUP, "sb_trws_grid: at least four columns per rank");
        blocks = world > 1 ? (opt.col_blocks > 0 ? opt.col_blocks : default_col_blocks(W, world)) : 1;
        SB_REQUIRE(world == 1 || world * blocks <= W / 4, SB_EUNSUP, "sb_trws_grid: at least four columns per column block");
        band = band_window(H, W, world > 1 ? rank : -1, world, blocks);
        SB_REQUIRE(band.nblocks >= 1, SB_EUNSUP, "sb_trws_grid: rank %d has no column block", rank);
        rows = H;
        Wl = band.Wl;
        blk_nodes = (int64_t)H * Wl;
        Nloc = band.nodes();
        // several ranks may live in one process (sb_trws_grid_attach_local): their sweeps must run concurrently
        if (world > 1) SB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        SB_REQUIRE(Nloc < (1LL << 30), SB_EUNSUP, "sb_trws_grid: too many nodes per rank");
        int dev = 0, num_sms = 0;
        SB_CUDA(cudaGetDevice(&dev));
        SB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        plan = grid_plan(dev, H, W, rank, world, blocks);

        dNodeF.alloc((size_t)Nloc * 4 * LP);
        dNodeB.alloc((size_t)Nloc * LP);
        dMsg.alloc((size_t)Nloc * 4 * LP);
        dPairB.alloc((size_t)Nloc * 2 * 6 * LP);
        dAlpha.alloc((size_t)Nloc * 4);
        dSelBox.alloc((size_t)Nloc * 4);
        dSol.alloc((size_t)Nloc);
        dCtrl.alloc(sizeof(Ctrl));
        dSave.alloc((size_t)std::max(plan->save_slots, 1) * LP * (sizeof(REAL) / 4));
        SB_CUDA(cudaMemsetAsync(dSave.p, 0, dSave.bytes(), stream));
        SB_CUDA(cudaMemsetAsync(dNodeF.p, 0, dNodeF.bytes(), stream));
        SB_CUDA(cudaMemsetAsync(dAlpha.p, 0, dAlpha.bytes(), stream));
        SB_CUDA(cudaMemsetAsync(dSelBox.p, 0, dSelBox.bytes(), stream));
        SB_CUDA(cudaMallocHost((void **)&hc, sizeof(Ctrl)));
        SB_CUDA(cudaMallocHost((void **)&hq, sizeof(Ctrl) * MAX_PENDING));

        std::memset(&P, 0, sizeof(P));
        P.H = H; P.W = Wl; P.L = L; P.LP = LP; P.rows = rows; P.Nloc = Nloc;
        P.nodeF = dNodeF.p; P.nodeB = dNodeB.p; P.msg = dMsg.p; P.pairB = dPairB.p; P.alpha = dAlpha.p;
        P.selbox = dSelBox.p; P.lambda = (REAL)tol;
        P.S = plan->S[0]; P.world = world; P.sol = dSol.p; P.save = dSave.p;
        Ctrl *ctrl = reinterpret_cast<Ctrl *>(dCtrl.p);
        P.ticket = &ctrl->ticket; P.acc = ctrl->acc; P.ring_smid = &ctrl->ring_smid;

        if ((getenv("SB_TRWS_PROFILE") || getenv("SB_TRWS_RECORD")) && !DIAG)
            fprintf(stderr, "[stereo_b200] SB_TRWS_PROFILE / SB_TRWS_RECORD need the diagnostics build of the grid sweep: "
                            "make -C stereo_b200/csrc clean all EXTRA=-DSB_GTRWS_DIAG=1\n");
        if (getenv("SB_TRWS_PROFILE")) {
            dProf.alloc(144);
            SB_CUDA(cudaMemsetAsync(dProf.p, 0, dProf.bytes(), stream));
            P.prof = dProf.p;
            P.prof_warp = (atoi(getenv("SB_TRWS_PROFILE")) - 1) & 3;
        }
        if (const char *e = getenv("SB_TRWS_WATCHDOG_MS")) watchdog_ms = atof(e);
        P.poll_ns = 100;
        if (const char *e = getenv("SB_GTRWS_POLL_NS")) P.poll_ns = (unsigned)std::max(0, atoi(e));

        if (getenv("SB_TRWS_RECORD")) {
            rec_ctas = 1024;
            SB_CUDA(cudaHostAlloc((void **)&rec_host, (size_t)rec_ctas * 6 * 4 * sizeof(int), cudaHostAllocMapped));
            SB_CUDA(cudaHostGetDevicePointer((void **)&P.rec, rec_host, 0));
        }
        // Throughput or latency build?  A pass cannot end before the DAG's critical path (the ring chain, then H + W
        // dependent node steps; across ranks also the pipeline fill): when the walkers have few node steps each next
        // to that, the build with the shorter node step wins (two walkers per SM, operands taken ahead of the chain),
        // where a large grid on one GPU wants as many walkers as fit.  Measured on one B200 (ms per pass, throughput /
        // latency build): 375x450x64 2.60 / 2.44, 540x960x128 5.68 / 5.48, 1980x360x192 12.2 / 11.2, 1080x1920x128
        // 13.7 / 15.3, 1980x2880x192 38.1 / 41.5 -- the break-even is near 3 (H + W) node steps per SM.
        lat = (double)Nloc / num_sms <= (world > 1 ? 6.0 : 3.0) * (double)(H + W);
        if (opt.latency_mode) lat = opt.latency_mode > 0;
        if (const char *e = getenv("SB_GTRWS_LAT")) lat = atoi(e) != 0;
        auto grid_for = [&](int pass) {
            const int bps = ops->blocks_per_sm(precision, kernel, pass, lat);
            SB_REQUIRE(bps >= 1, SB_ECUDA, "sb_trws_grid: sweep kernel does not fit on an SM");
            long long g = (long long)bps * num_sms;
            if (const char *e = getenv("SB_GTRWS_CTAS_PER_SM")) g = std::max(1, std::min(bps, atoi(e))) * (long long)num_sms;
            const int Sp = plan->S[pass == PASS_FWD ? 0 : 1];
            if (g > Sp) g = Sp;
            // Every SM gets the same number of walkers.  (Rounding the grid down to "even waves" -- the fewest walkers that
            // keep ceil(S / g) -- leaves some SMs with one walker fewer; those walkers run faster than the rows they
            // follow and only wait, and the pass was 1-3.5 % slower at every shape measured.  SB_GTRWS_EVEN_WAVES=1.)
            if (getenv("SB_GTRWS_EVEN_WAVES") && g >= 1) {
                const long long rounds = (Sp + g - 1) / g;
                g = (Sp + rounds - 1) / rounds;
            }
            if (rec_host && g > rec_ctas) g = rec_ctas;
            // two strip walkers must be in flight: (H-3,1) waits for (H-2,1) (trws_order.cpp)
            SB_REQUIRE(Sp < 2 || g >= 2, SB_ECUDA, "sb_trws_grid: fewer than two resident CTAs");
            return (int)std::max<long long>(g, 1);
        };
        grid_fwd = grid_for(PASS_FWD);
        grid_bwd = grid_for(PASS_BWD);
        // the ring chain gets an SM of its own when there are CTAs to spare (at least three per SM)
        isolate = grid_fwd >= 3 * num_sms && !getenv("SB_GTRWS_NO_ISOLATE");
        reset();
        SB_CUDA(cudaStreamSynchronize(stream));
        setup_ms = now_ms() - t0;
    }

    ~Solver() override
    {
        for (int d = 0; d < 2; d++)
            for (int a = 0; a < 2; a++)
                if (peer_ptr[d][a]) cudaIpcCloseMemHandle(peer_ptr[d][a]);
        if (stream) { cudaStreamSynchronize(stream); }
        if (hc) cudaFreeHost(hc);
        if (hq) cudaFreeHost(hq);
        for (const Pending &pd : pending) { cudaEventDestroy(pd.e0); cudaEventDestroy(pd.e1); }
        for (cudaEvent_t e : event_pool) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }

    void set_labels(int l0, int nl, const double *planes, const double *unary, double d_min, double d_step) override
    {
        SB_REQUIRE(l0 >= 0 && nl >= 1 && l0 + nl <= L, SB_EINVAL, "sb_trws_grid_set_labels: labels [%d, %d) outside [0, %d)", l0, l0 + nl, L);
        SB_REQUIRE(planes && unary, SB_EINVAL, "sb_trws_grid_set_labels: null pointer");
        SB_REQUIRE(d_step != 0.0, SB_EINVAL, "sb_trws_grid_set_labels: d_step == 0");
        const double t0 = now_ms();
        // proposals are uploaded in groups that keep the staging buffers below ~512 MB
        const int grp = (int)std::max<int64_t>(1, std::min<int64_t>(nl, (512LL << 20) / (40 * N)));
        DevBuf<double> dPl((size_t)grp * 4 * N), dUn((size_t)grp * N);
        DevBuf<int> dBad(1);
        SB_CUDA(cudaMemsetAsync(dBad.p, 0, sizeof(int), stream));
        for (int g0 = 0; g0 < nl; g0 += grp) {
            const int ng = std::min(grp, nl - g0);
            // (cudaMemcpyDefault: the proposals may already live on the device -- a fusion loop that keeps its plane
            // fields resident passes device pointers)
            SB_CUDA(cudaMemcpyAsync(dPl.p, planes + (size_t)g0 * 4 * N, (size_t)ng * 4 * N * 8, cudaMemcpyDefault, stream));
            SB_CUDA(cudaMemcpyAsync(dUn.p, unary + (size_t)g0 * N, (size_t)ng * N * 8, cudaMemcpyDefault, stream));
            const long long tot = blk_nodes * ng;
            for (int b = 0; b < band.nblocks; b++) {
                gfill_labels_kernel<REAL><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(
                    dPl.p, dUn.p, ng, l0 + g0, H, W, 0, rows, band.c_base(b), Wl, LP, d_min, d_step, dNodeF.p + (size_t)b * blk_nodes * 4 * LP, dBad.p);
                SB_CUDA(cudaGetLastError());
                count_launch();
            }
        }
        int bad = 0;
        SB_CUDA(cudaMemcpyAsync(&bad, dBad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
        SB_REQUIRE(!(bad & 1), SB_EINVAL, "Infinite disparity");                 // dispmap_super.m:321-323
        SB_REQUIRE(!(bad & 2), SB_EINVAL, "q contains NaN");                     // trws.m:9-15
        SB_REQUIRE(!(bad & 4), SB_EINVAL, "sb_trws_grid_set_labels: unary contains NaN");
        finalized = false;
        setup_ms += now_ms() - t0;
    }

    void set_weights(const double *alphas) override
    {
        SB_REQUIRE(alphas, SB_EINVAL, "sb_trws_grid_set_weights: null pointer");
        const double t0 = now_ms();
        DevBuf<double> dA((size_t)E);
        SB_CUDA(cudaMemcpyAsync(dA.p, alphas, (size_t)E * 8, cudaMemcpyDefault, stream));
        for (int b = 0; b < band.nblocks; b++) {
            gweights_kernel<REAL><<<(unsigned)((2 * blk_nodes + 255) / 256), 256, 0, stream>>>(dA.p, H, W, 0, rows, band.c_base(b), Wl,
                                                                                              dAlpha.p + (size_t)b * blk_nodes * 4);
            SB_CUDA(cudaGetLastError());
            count_launch();
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        setup_ms += now_ms() - t0;
    }

    void synth(uint64_t seed, int Hs, int Ws, int r_off, int c_off) override
    {
        const double t0 = now_ms();
        if (Hs <= 0 || Ws <= 0) { Hs = H; Ws = W; r_off = 0; c_off = 0; }
        SB_REQUIRE(r_off >= 0 && c_off >= 0 && r_off + H <= Hs && c_off + W <= Ws, SB_EINVAL,
                   "sb_trws_grid_synth: window (%d,%d)+%dx%d outside the %dx%d scene", r_off, c_off, H, W, Hs, Ws);
        const long long tot = blk_nodes * L;
        for (int b = 0; b < band.nblocks; b++) {
            gsynth_kernel<REAL><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(
                (unsigned)(seed ^ (seed >> 32)), kernel, Hs, Ws, r_off, c_off, Wl, W, L, 0, rows, band.c_base(b), LP,
                dNodeF.p + (size_t)b * blk_nodes * 4 * LP, dAlpha.p + (size_t)b * blk_nodes * 4);
            SB_CUDA(cudaGetLastError());
            count_launch();
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        finalized = false;
        setup_ms += now_ms() - t0;
    }

    void finalize() override
    {
        const double t0 = now_ms();
        GTablesLaunch tl;
        tl.precision = precision; tl.nodeF = dNodeF.p; tl.nodeB = dNodeB.p; tl.pairB = dPairB.p;
        // (the blocks are stacked: to the table kernels they are one grid of nblocks * H rows; the pairs that would join
        // two blocks are never read)
        tl.W = Wl; tl.rows = rows * band.nblocks; tl.L = L; tl.stream = stream;
        ops->tables(tl);
        SB_CUDA(cudaStreamSynchronize(stream));
        finalized = true;
        setup_ms += now_ms() - t0;
    }

    void get_label(int l, double *out) override
    {
        SB_REQUIRE(l >= 0 && l < L && out, SB_EINVAL, "sb_trws_grid_get_label: bad arguments");
        DevBuf<double> d((size_t)4 * N);
        SB_CUDA(cudaMemsetAsync(d.p, 0, d.bytes(), stream));
        for (int b = 0; b < band.nblocks; b++) {
            gget_label_kernel<REAL><<<(unsigned)((blk_nodes + 255) / 256), 256, 0, stream>>>(dNodeF.p + (size_t)b * blk_nodes * 4 * LP, l, H, W, 0, rows,
                                                                                            band.c_base(b), Wl, LP, d.p);
            SB_CUDA(cudaGetLastError());
            count_launch();
        }
        SB_CUDA(cudaMemcpyAsync(out, d.p, d.bytes(), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
    }

    void get_weights(double *out) override
    {
        SB_REQUIRE(out, SB_EINVAL, "sb_trws_grid_get_weights: null pointer");
        DevBuf<double> d((size_t)E);
        SB_CUDA(cudaMemsetAsync(d.p, 0, d.bytes(), stream));
        for (int b = 0; b < band.nblocks; b++) {
            gget_weights_kernel<REAL><<<(unsigned)((2 * blk_nodes + 255) / 256), 256, 0, stream>>>(dAlpha.p + (size_t)b * blk_nodes * 4, H, W, 0, rows,
                                                                                                  band.c_base(b), Wl, d.p);
            SB_CUDA(cudaGetLastError());
            count_launch();
        }
        SB_CUDA(cudaMemcpyAsync(out, d.p, d.bytes(), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
    }

    // ZeroMessages (MRFEnergy.cpp:115-131): +0.0 everywhere = pass-counter parity 0
    void reset() override
    {
        SB_CUDA(cudaMemsetAsync(dSol.p, 0, dSol.bytes(), stream));
        SB_CUDA(cudaMemsetAsync(dMsg.p, 0, dMsg.bytes(), stream));
        // (a neighbouring rank writes into these arrays: the caller's barrier must find them cleared)
        if (world > 1) SB_CUDA(cudaStreamSynchronize(stream));
        pass_counter = 0;
    }

    // One pass = memset of the control block, the persistent sweep launch, the control block back to a pinned slot.
    // launch_pass only enqueues (the ranks of a multi-GPU run are not separated by host round trips: a message word
    // validates itself, so a rank simply runs ahead until it needs a word its neighbour has not written yet);
    // wait_passes synchronises, with the watchdog, and accounts the launch times.
    struct Pending { cudaEvent_t e0, e1; };
    std::vector<Pending> pending;            // launched, not yet waited for
    std::vector<cudaEvent_t> event_pool;
    Ctrl *hq = nullptr;                      // pinned: control blocks of the pending passes
    static constexpr int MAX_PENDING = 256;
    int last_grid = 0, last_pass = 0, last_mode = 0;

    cudaEvent_t get_event()
    {
        if (!event_pool.empty()) { cudaEvent_t e = event_pool.back(); event_pool.pop_back(); return e; }
        cudaEvent_t e;
        SB_CUDA(cudaEventCreate(&e));
        return e;
    }

    void launch_pass(int pass, int mode)
    {
        SB_REQUIRE(finalized, SB_EINVAL, "sb_trws_grid: call sb_trws_grid_finalize after the labels are set");
        SB_REQUIRE((int)pending.size() < MAX_PENDING, SB_EINVAL, "sb_trws_grid: too many passes in flight (wait first)");
        SB_CUDA(cudaMemsetAsync(dCtrl.p, 0, dCtrl.bytes(), stream));
        const bool sends = pass == PASS_BWD || (mode & MODE_SEND);
        if (sends) ++pass_counter;
        P.tag = pass_counter & 1u;
        P.epoch = ++launch_epoch;
        P.mode = mode;
        const int pi = pass == PASS_FWD ? 0 : 1;
        if (dProf.p) P.prof = dProf.p + 72 * pi;
        P.head_strip = pass == PASS_FWD ? (world > 1 ? -1 : 1) : 0;
        if (const char *e = getenv("SB_PROF_STRIP")) P.head_strip = atoi(e);
        P.segs = plan->segs[pi].p;
        P.seg_ptr = plan->seg_ptr[pi].p;
        P.strip_len = plan->strip_len[pi].p;
        P.is_ring = plan->is_ring[pi].p;
        P.S = plan->S[pi];
        P.isolate_ring = (pass == PASS_FWD && isolate) ? 1 : 0;
        GSweepLaunch sl;
        sl.precision = precision; sl.kern = kernel; sl.pass = pass; sl.problem = &P;
        sl.grid = pass == PASS_FWD ? grid_fwd : grid_bwd; sl.stream = stream; sl.lat = lat;
        last_grid = sl.grid; last_pass = pass; last_mode = mode;
        if (rec_host) std::memset(rec_host, 0xff, (size_t)rec_ctas * 6 * 4 * sizeof(int));
        Pending pd;
        pd.e0 = get_event();
        pd.e1 = get_event();
        SB_CUDA(cudaEventRecord(pd.e0, stream));
        ops->sweep(sl);
        SB_CUDA(cudaEventRecord(pd.e1, stream));
        SB_CUDA(cudaMemcpyAsync(hq + pending.size(), dCtrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
        pending.push_back(pd);
    }

    // acc: 2 doubles (energy, lower-bound contribution) per pending pass, in launch order (may be null)
    void wait_passes(double *acc)
    {
        // watchdog: a sweep that does not finish (a dependency that never arrives) cannot be cancelled --
        // report where it stands and leave, instead of hanging the caller forever
        const double t_start = now_ms();
        const double limit = watchdog_ms * (double)std::max<size_t>(pending.size(), 1);
        cudaError_t q;
        while ((q = cudaStreamQuery(stream)) == cudaErrorNotReady && now_ms() - t_start < limit) usleep(50);
        if (q == cudaErrorNotReady) {
            fprintf(stderr, "[sb_trws_grid] rank %d/%d: %zu pass(es) in flight (last: pass=%d mode=%d epoch=%u tag=%u grid=%d) did not finish within %.0f ms: HANG\n",
                    rank, world, pending.size(), last_pass, last_mode, P.epoch, P.tag, last_grid, limit);
            if (rec_host)
                for (int c = 0; c < last_grid && c < rec_ctas; c++) {
                    fprintf(stderr, "[sb record] cta %d:", c);
                    for (int w = 0; w < 6; w++) {
                        const int *r = rec_host + ((size_t)c * 6 + w) * 4;
                        fprintf(stderr, " w%d(strip %d step %d ph %d x%x)", w, r[0], r[1], r[2], r[3]);
                    }
                    fprintf(stderr, "\n");
                }
            fflush(stderr);
            _exit(3);
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        for (size_t i = 0; i < pending.size(); i++) {
            float ms = 0;
            SB_CUDA(cudaEventElapsedTime(&ms, pending[i].e0, pending[i].e1));
            kernel_ms += ms;
            kernel_count++;
            total_kernel_ms += ms;
            total_kernel_count++;
            event_pool.push_back(pending[i].e0);
            event_pool.push_back(pending[i].e1);
            if (acc) { acc[2 * i] = hq[i].acc[0]; acc[2 * i + 1] = hq[i].acc[1]; }
        }
        if (!pending.empty()) *hc = hq[pending.size() - 1];
        pending.clear();
    }

    void run_pass(int pass, int mode)
    {
        launch_pass(pass, mode);
        wait_passes(nullptr);
    }

    void ipc_export(unsigned char *out) override
    {
        SB_REQUIRE(world > 1, SB_EINVAL, "sb_trws_grid_ipc_export: the solver was not created for several ranks");
        cudaIpcMemHandle_t h[2];
        SB_CUDA(cudaIpcGetMemHandle(&h[0], dMsg.p));
        SB_CUDA(cudaIpcGetMemHandle(&h[1], dSelBox.p));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        std::memcpy(out, h, sizeof(h));
    }

    void set_peer_geometry(int d, int peer_rank)
    {
        // Every rank stores its blocks with the same pitch, so my node (b, r, cl) is the peer's (b', r, cl -+ wb): a send
        // to the LEFT leaves through my halo column 0 = the peer's last owned column wb, a send to the RIGHT through my
        // column wb = the peer's halo column 0; b' = b except where the block sequence wraps around the ranks.
        (void)peer_rank;
        P.peer_dW[d] = 0;
        if (d == 0) P.peer_dc[d] = (rank == 0 ? -blk_nodes : 0) + band.wb;
        else P.peer_dc[d] = (rank == world - 1 ? blk_nodes : 0) - band.wb;
    }

    void *raw_ptr(int which) override { return which == 0 ? (void *)dMsg.p : (void *)dSelBox.p; }

    // neighbours that live in the SAME process and on the same device (several ranks on one GPU: how the
    // one-GPU test box exercises the banded sweep).  `share` solvers run their sweeps concurrently, so each
    // takes its part of the resident-CTA budget.
    void attach_local(SolverBase *up, SolverBase *down, int share) override
    {
        SB_REQUIRE(world > 1, SB_EINVAL, "sb_trws_grid_attach_local: the solver was not created for several ranks");
        SolverBase *src[2] = {up, down};
        for (int d = 0; d < 2; d++) {
            if (!src[d]) continue;
            const int peer_rank = (d == 0 ? rank - 1 + world : rank + 1) % world;    // the block sequence wraps around the ranks
            P.peer_msg[d] = static_cast<REAL *>(src[d]->raw_ptr(0));
            P.peer_selbox[d] = static_cast<unsigned long long *>(src[d]->raw_ptr(1));
            set_peer_geometry(d, peer_rank);
        }
        if (share > 1) {
            grid_fwd = std::max(2, grid_fwd / share);
            grid_bwd = std::max(2, grid_bwd / share);
            isolate = false;
        }
    }

    void ipc_attach(const unsigned char *up, const unsigned char *down) override
    {
        SB_REQUIRE(world > 1, SB_EINVAL, "sb_trws_grid_ipc_attach: the solver was not created for several ranks");
        const unsigned char *src[2] = {up, down};
        for (int d = 0; d < 2; d++) {
            if (!src[d]) continue;
            const int peer_rank = (d == 0 ? rank - 1 + world : rank + 1) % world;    // the block sequence wraps around the ranks
            void *ptr[2];
            if (d == 1 && src[0] && std::memcmp(src[0], src[1], 2 * sizeof(cudaIpcMemHandle_t)) == 0) {
                // two ranks: the neighbour on both sides is the same process, and a handle opens only once
                ptr[0] = peer_ptr[0][0]; ptr[1] = peer_ptr[0][1];
            } else {
                cudaIpcMemHandle_t h[2];
                std::memcpy(h, src[d], sizeof(h));
                for (int a = 0; a < 2; a++)
                    SB_CUDA(cudaIpcOpenMemHandle(&peer_ptr[d][a], h[a], cudaIpcMemLazyEnablePeerAccess));
                ptr[0] = peer_ptr[d][0]; ptr[1] = peer_ptr[d][1];
            }
            P.peer_msg[d] = static_cast<REAL *>(ptr[0]);
            P.peer_selbox[d] = static_cast<unsigned long long *>(ptr[1]);
            set_peer_geometry(d, peer_rank);
        }
    }

    void run_one_pass(int pass, int mode, double *acc) override
    {
        run_pass(pass == 0 ? PASS_FWD : PASS_BWD, mode);
        acc[0] = hc->acc[0];
        acc[1] = hc->acc[1];
    }
    void launch_one_pass(int pass, int mode) override { launch_pass(pass == 0 ? PASS_FWD : PASS_BWD, mode); }
    int wait_all(double *acc, int max_passes) override
    {
        const int n = (int)pending.size();
        SB_REQUIRE(!acc || max_passes >= n, SB_EINVAL, "sb_trws_grid_wait: %d passes in flight, room for %d", n, max_passes);
        wait_passes(acc);
        return n;
    }

    // minimize.cpp:31-113 (same driver as trws_solve.cu)
    void minimize(double maxiter, double max_relgap, double *energy_out, double *lb_out, double *iters_out,
                  sb_trws_timing *timing) override
    {
        SB_REQUIRE(world == 1, SB_EINVAL, "sb_trws_grid_minimize: a banded solver is driven pass by pass (sb_trws_grid_pass)");
        const int64_t launches0 = g_launches.load();
        kernel_ms = 0;
        kernel_count = 0;
        EventTimer timer(stream);
        timer.start();
        const int iter_max = (int)maxiter;
        double energy = 0, lb = 0;
        int iterations = 0;
        for (int it = 1;; it++) {
            run_pass(PASS_FWD, MODE_SEND | ((fuse && it > 1) ? MODE_ROUND : 0));
            if (fuse && it > 1) {
                energy = hc->acc[0];
                if ((energy - lb) / energy < max_relgap) { iterations = it - 1; break; }
            }
            run_pass(PASS_BWD, 0);
            lb = hc->acc[1];
            if (!fuse || it >= iter_max) {
                run_pass(PASS_FWD, MODE_ROUND);
                energy = hc->acc[0];
                if (it >= iter_max || (energy - lb) / energy < max_relgap) { iterations = it; break; }
            }
        }
        const double solve_ms = timer.stop_ms();
        if (P.prof) {
            long long h[144];
            SB_CUDA(cudaMemcpy(h, dProf.p, sizeof(h), cudaMemcpyDeviceToHost));
            SB_CUDA(cudaMemsetAsync(dProf.p, 0, sizeof(h), stream));
            for (int g = 0; g < 6; g++) {
                const long long *q = h + 24 * g;
                const double nt = q[7] ? (double)q[7] : 1, nh = q[16] ? (double)q[16] : 1;
                fprintf(stderr, "[sb gprofile] %s %s term (%lld steps): waitFULL=%.0f waitstage=%.0f total+round=%.0f operands=%.0f update=%.0f "
                        "stores=%.0f tail=%.0f | helper (%lld): waitfree=%.0f waitstage=%.0f static=%.0f msgpoll=%.0f roundpoll=%.0f write=%.0f gate=%.0f issuework=%.0f cyc/step, %.2f poll retries/step\n",
                        g >= 3 ? "bwd" : "fwd", (g % 3) == 0 ? "ring" : (g % 3) == 1 ? "rows" : "head row", q[7], q[0] / nt, q[1] / nt, q[2] / nt, q[3] / nt, q[4] / nt, q[5] / nt, q[6] / nt, q[16],
                        q[8] / nh, q[9] / nh, q[10] / nh, q[11] / nh, q[12] / nh, q[13] / nh, q[14] / nh, q[15] / nh, q[17] / nh);
            }
        }
        *energy_out = energy;
        *lb_out = lb;
        *iters_out = (double)iterations;
        if (timing) {
            std::memset(timing, 0, sizeof(*timing));
            timing->setup_ms = setup_ms;
            timing->solve_ms = solve_ms;
            timing->sweep_ms_avg = iterations ? solve_ms / iterations : 0;
            timing->kernel_launches = g_launches.load() - launches0;
            timing->sweep_kernel_ms = kernel_ms;
            timing->sweep_kernel_launches = kernel_count;
        }
    }

    void labels(double *out) override
    {
        DevBuf<double> d((size_t)N);
        SB_CUDA(cudaMemsetAsync(d.p, 0, d.bytes(), stream));
        for (int b = 0; b < band.nblocks; b++) {
            const long long n = (long long)H * (band.c_hi(b) - band.c_lo(b));
            if (n <= 0) continue;
            glabels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dSol.p + (size_t)b * blk_nodes, H, band.c_lo(b), band.c_hi(b), band.c_base(b),
                                                                            Wl, d.p);
            SB_CUDA(cudaGetLastError());
            count_launch();
        }
        SB_CUDA(cudaMemcpyAsync(out, d.p, d.bytes(), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
    }

    void counters(double *out) override
    {
        out[0] = total_kernel_ms;
        out[1] = (double)total_kernel_count;
        out[2] = setup_ms;
    }

    int latency_build() const override { return lat; }
    void info(int64_t *out) override
    {
        out[0] = (int64_t)(dNodeF.bytes() + dNodeB.bytes() + dMsg.bytes() + dPairB.bytes() + dAlpha.bytes() + dSelBox.bytes() + dSol.bytes());
        out[1] = Nloc;
        out[2] = band.c_lo(0);
        out[3] = band.c_hi(0);
        out[4] = grid_fwd;
        out[5] = grid_bwd;
        out[6] = (int64_t)ops->smem_bytes(precision);
        out[7] = LP;
    }
};

} // namespace
} // namespace gtrws
} // namespace sb

extern "C" {

struct sb_trws_grid {
    sb::gtrws::SolverBase *impl;
};

int sb_trws_grid_create(int kernel, int H, int W, int L, double tol, const sb_trws_options *opt_in, int rank, int world,
                        sb_trws_grid **out)
{
    return sb::guarded([&] {
        SB_REQUIRE(out, SB_EINVAL, "sb_trws_grid_create: null output");
        *out = nullptr;
        SB_REQUIRE(world >= 1 && rank >= 0 && rank < world, SB_EINVAL, "sb_trws_grid_create: bad rank / world");
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unsupported kernel");   // trws_mex.cpp:156-163
        SB_REQUIRE(L >= 1 && H >= 1 && W >= 1, SB_EINVAL, "sb_trws_grid_create: bad sizes H=%d W=%d L=%d", H, W, L);
        SB_REQUIRE(H >= 4 && W >= 4, SB_EUNSUP, "sb_trws_grid_create: the grid-native path needs H, W >= 4 (use sb_trws_solve)");
        sb_trws_options opt;
        if (opt_in) opt = *opt_in; else sb_trws_default_options(&opt);
        SB_REQUIRE(opt.precision == SB_F32 || opt.precision == SB_F64, SB_EINVAL, "sb_trws_grid_create: bad precision");
        sb::require_device();
        sb::gtrws::SolverBase *impl;
        if (opt.precision == SB_F64) impl = new sb::gtrws::Solver<double>(kernel, H, W, L, tol, opt, rank, world);
        else impl = new sb::gtrws::Solver<float>(kernel, H, W, L, tol, opt, rank, world);
        *out = new sb_trws_grid{impl};
    });
}

#define SB_GRID_ENTRY(name, check, call)                                                     \
    return sb::guarded([&] {                                                                 \
        SB_REQUIRE(g && g->impl && (check), SB_EINVAL, name ": null pointer");               \
        call;                                                                                \
    })

int sb_trws_grid_set_labels(sb_trws_grid *g, int l0, int nl, const double *planes, const double *unary, double d_min, double d_step)
{
    SB_GRID_ENTRY("sb_trws_grid_set_labels", true, g->impl->set_labels(l0, nl, planes, unary, d_min, d_step));
}
int sb_trws_grid_set_weights(sb_trws_grid *g, const double *alphas)
{
    SB_GRID_ENTRY("sb_trws_grid_set_weights", true, g->impl->set_weights(alphas));
}
int sb_trws_grid_synth(sb_trws_grid *g, uint64_t seed)
{
    SB_GRID_ENTRY("sb_trws_grid_synth", true, g->impl->synth(seed, 0, 0, 0, 0));
}
int sb_trws_grid_synth_window(sb_trws_grid *g, uint64_t seed, int scene_H, int scene_W, int r_off, int c_off)
{
    SB_GRID_ENTRY("sb_trws_grid_synth_window", true, g->impl->synth(seed, scene_H, scene_W, r_off, c_off));
}
int sb_trws_grid_finalize(sb_trws_grid *g)
{
    SB_GRID_ENTRY("sb_trws_grid_finalize", true, g->impl->finalize());
}
int sb_trws_grid_get_label(sb_trws_grid *g, int l, double *out)
{
    SB_GRID_ENTRY("sb_trws_grid_get_label", out, g->impl->get_label(l, out));
}
int sb_trws_grid_get_weights(sb_trws_grid *g, double *alphas)
{
    SB_GRID_ENTRY("sb_trws_grid_get_weights", alphas, g->impl->get_weights(alphas));
}
int sb_trws_grid_reset(sb_trws_grid *g)
{
    SB_GRID_ENTRY("sb_trws_grid_reset", true, g->impl->reset());
}
int sb_trws_grid_minimize(sb_trws_grid *g, double maxiter, double max_relgap, double *energy, double *lower_bound,
                          double *iterations, sb_trws_timing *timing)
{
    SB_GRID_ENTRY("sb_trws_grid_minimize", energy && lower_bound && iterations,
                  g->impl->minimize(maxiter, max_relgap, energy, lower_bound, iterations, timing));
}
int sb_trws_grid_get_labels(sb_trws_grid *g, double *labels)
{
    SB_GRID_ENTRY("sb_trws_grid_get_labels", labels, g->impl->labels(labels));
}
int sb_trws_grid_ipc_export(sb_trws_grid *g, unsigned char *handles)
{
    SB_GRID_ENTRY("sb_trws_grid_ipc_export", handles, g->impl->ipc_export(handles));
}
int sb_trws_grid_ipc_attach(sb_trws_grid *g, const unsigned char *up, const unsigned char *down)
{
    SB_GRID_ENTRY("sb_trws_grid_ipc_attach", true, g->impl->ipc_attach(up, down));
}
int sb_trws_grid_pass(sb_trws_grid *g, int pass, int mode, double *acc)
{
    SB_GRID_ENTRY("sb_trws_grid_pass", acc && (pass == 0 || pass == 1) && mode >= 0 && mode <= 3,
                  g->impl->run_one_pass(pass, mode, acc));
}
int sb_trws_grid_launch_pass(sb_trws_grid *g, int pass, int mode)
{
    SB_GRID_ENTRY("sb_trws_grid_launch_pass", (pass == 0 || pass == 1) && mode >= 0 && mode <= 3, g->impl->launch_one_pass(pass, mode));
}
int sb_trws_grid_wait(sb_trws_grid *g, double *acc, int max_passes, int *n_passes)
{
    SB_GRID_ENTRY("sb_trws_grid_wait", true, { const int n = g->impl->wait_all(acc, max_passes); if (n_passes) *n_passes = n; });
}
int sb_trws_grid_attach_local(sb_trws_grid *g, sb_trws_grid *up, sb_trws_grid *down, int share)
{
    SB_GRID_ENTRY("sb_trws_grid_attach_local", true, g->impl->attach_local(up ? up->impl : nullptr, down ? down->impl : nullptr, share));
}
int sb_trws_grid_counters(sb_trws_grid *g, double *out)
{
    SB_GRID_ENTRY("sb_trws_grid_counters", out, g->impl->counters(out));
}
int sb_trws_grid_info(sb_trws_grid *g, int64_t *info)
{
    SB_GRID_ENTRY("sb_trws_grid_info", info, g->impl->info(info));
}
int sb_trws_grid_latency_mode(sb_trws_grid *g, int *on)
{
    SB_GRID_ENTRY("sb_trws_grid_latency_mode", on, *on = g->impl->latency_build());
}
void sb_trws_grid_destroy(sb_trws_grid *g)
{
    if (!g) return;
    delete g->impl;
    delete g;
}

// One UpdateMessage on the device (unit-level known-answer tests against tests/golden/update_message.npz)
int sb_trws_update_message(int kernel, int L, const double *Di, const double *msg, const double *src_pos, const double *dst_pos,
                           double alpha, double lambda, double gamma, int precision, double *msg_out, double *vmin_out)
{
    return sb::guarded([&] {
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unsupported kernel");
        SB_REQUIRE(Di && msg && src_pos && dst_pos && msg_out && vmin_out && L >= 1, SB_EINVAL, "sb_trws_update_message: bad arguments");
        const sb::gtrws::GOps *ops = sb::gtrws::gops_for_labels(L);
        SB_REQUIRE(ops, SB_EUNSUP, "sb_trws_update_message: %d labels exceed SB_MAX_LABELS=%d", L, SB_MAX_LABELS);
        sb::require_device();
        sb::DevBuf<double> d((size_t)5 * L + 1);
        const double *src[4] = {Di, msg, src_pos, dst_pos};
        for (int i = 0; i < 4; i++) SB_CUDA(cudaMemcpy(d.p + (size_t)i * L, src[i], (size_t)L * 8, cudaMemcpyHostToDevice));
        sb::gtrws::GUpdateLaunch a;
        a.precision = precision; a.kern = kernel; a.L = L;
        a.Di = d.p; a.msg = d.p + L; a.src = d.p + 2 * L; a.dst = d.p + 3 * L;
        a.alpha = alpha; a.lambda = lambda; a.gamma = gamma;
        a.msg_out = d.p + 4 * L; a.vmin_out = d.p + 5 * L; a.stream = 0;
        ops->update_message(a);
        SB_CUDA(cudaMemcpy(msg_out, a.msg_out, (size_t)L * 8, cudaMemcpyDeviceToHost));
        SB_CUDA(cudaMemcpy(vmin_out, a.vmin_out, 8, cudaMemcpyDeviceToHost));
    });
}

// host-only: schedule statistics of the grid-native plan (tests)
int sb_trws_grid_plan_stats(int H, int W, int rank, int world, int64_t *stats)
{
    return sb::guarded([&] {
        SB_REQUIRE(stats, SB_EINVAL, "sb_trws_grid_plan_stats: null pointer");
        for (int pass = 0; pass < 2; pass++) {
            sb::gtrws::GPassPlan plan;
            sb::gtrws::build_gpass_plan(H, W, pass, world > 1 ? rank : -1, world, 0, plan);
            int64_t steps = 0, nodes = 0, two = 0, peer_up = 0, peer_down = 0;
            for (const auto &s : plan.segs) {
                steps += s.n;
                if (!(s.flags & sb::gtrws::GF_DEFERRED)) nodes += s.n;
                if (s.flags & sb::gtrws::GF_SAVE) two += s.n;
                for (int d = 0; d < 4; d++) {
                    const int role = (int)((s.roles >> (4 * d)) & 15u);
                    if (role != sb::gtrws::ROLE_SEND0 && role != sb::gtrws::ROLE_SEND1) continue;
                    if (s.peer[d] == 1) peer_up += s.n;
                    if (s.peer[d] == 2) peer_down += s.n;
                }
            }
            stats[0 + 6 * pass] = (int64_t)plan.strip_len.size();
            stats[1 + 6 * pass] = (int64_t)plan.segs.size();
            stats[2 + 6 * pass] = steps;
            stats[3 + 6 * pass] = nodes;
            stats[4 + 6 * pass] = two;
            stats[5 + 6 * pass] = peer_up * 1000000 + peer_down;
        }
    });
}

} // extern "C"
