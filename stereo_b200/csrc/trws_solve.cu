// trws_solve.cu -- host driver of the TRW-S path behind sb_trws_solve
// (include/stereo_b200.h).  Mirrors solve_mrf() of cpp/trws_mex.cpp:27-147:
// validate, build the graph data, order the nodes, iterate forward / backward
// sweeps with the reference's stopping rule (cpp/trw-s/minimize.cpp:97-112),
// return labels (1-based), energy, lower bound and iteration count.
#include "sb_common.h"
#include "trws_order.h"
#include "trws_kernels.cuh"
#include "trws_launch.h"
#include <vector>
#include <unistd.h>
#include <map>
#include <tuple>
#include <memory>
#include <mutex>
#include <chrono>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace sb {
namespace trws {

const KOps *kops_for_labels(int L)
{
    static const KOps *table[] = {&kops_k1, &kops_k2, &kops_k3, &kops_k4, &kops_k6, &kops_k8};
    for (const KOps *k : table)
        if (32 * k->K >= L) return k;
    return nullptr;
}

namespace {

struct Ctrl {
    int ticket;
    int pad;
    double acc[2];
    // followed by int32 progress[S] (zeroed together with the rest before each launch)
};

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// The sweep schedule depends only on the grid shape: built once per (device, H, W) and kept
// resident (the reference re-runs its O(N * perimeter) SetAutomaticOrdering on every call,
// ordering.cpp:7-157).
struct GridPlan {
    int S = 0;          // strips this rank sweeps
    int S_global = 0;   // strips of the whole schedule
    DevBuf<Segment> segs[2];
    DevBuf<int32_t> seg_ptr[2];
    DevBuf<int32_t> strip_len, strip_gid;
};

std::shared_ptr<GridPlan> grid_plan(int dev, int H, int W, int rank, int world)
{
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int, int>, std::shared_ptr<GridPlan>> cache;
    std::lock_guard<std::mutex> lock(mu);
    const auto key = std::make_tuple(dev, H, W, rank, world);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<int32_t> order;
    SB_REQUIRE(grid_ordering(H, W, order), SB_EINVAL,
               "sb_trws_solve: %dx%d grid has no valid automatic ordering (the reference crashes on it)", H, W);
    std::vector<uint8_t> info;
    build_node_info(H, W, order, info);
    Schedule sched;
    build_schedule(H, W, order, sched, world);
    auto gp = std::make_shared<GridPlan>();
    gp->S_global = (int)sched.strip_ptr.size() - 1;
    for (int pass = 0; pass < 2; pass++) {
        PassPlan plan;
        build_pass_plan(H, W, info, sched, pass, plan, world > 1 ? rank : -1);
        if (pass == 0) {
            gp->S = (int)plan.strips.size();
            gp->strip_len.alloc(std::max<size_t>(plan.strips.size(), 1));
            gp->strip_gid.alloc(std::max<size_t>(plan.strips.size(), 1));
            SB_CUDA(cudaMemcpy(gp->strip_len.p, plan.strip_len.data(), plan.strip_len.size() * 4, cudaMemcpyHostToDevice));
            SB_CUDA(cudaMemcpy(gp->strip_gid.p, plan.strips.data(), plan.strips.size() * 4, cudaMemcpyHostToDevice));
        }
        gp->segs[pass].alloc(std::max<size_t>(plan.segs.size(), 1));
        gp->seg_ptr[pass].alloc(plan.seg_ptr.size());
        SB_CUDA(cudaMemcpy(gp->segs[pass].p, plan.segs.data(), plan.segs.size() * sizeof(Segment), cudaMemcpyHostToDevice));
        SB_CUDA(cudaMemcpy(gp->seg_ptr[pass].p, plan.seg_ptr.data(), plan.seg_ptr.size() * 4, cudaMemcpyHostToDevice));
    }
    if (cache.size() >= 8) cache.clear();   // bounded: shapes rarely change within a session
    cache[key] = gp;
    return gp;
}

// Pinned result slots are recycled: cudaMallocHost / cudaFreeHost cost milliseconds (and
// synchronise the device), far too much per solve.
struct PinnedPool {
    std::mutex mu;
    std::vector<void *> free_list;
    void *get(size_t bytes)
    {
        {
            std::lock_guard<std::mutex> lock(mu);
            if (!free_list.empty()) {
                void *p = free_list.back();
                free_list.pop_back();
                return p;
            }
        }
        void *p = nullptr;
        SB_CUDA(cudaMallocHost(&p, std::max<size_t>(bytes, 256)));
        return p;
    }
    void put(void *p)
    {
        std::lock_guard<std::mutex> lock(mu);
        free_list.push_back(p);
    }
};
PinnedPool &pinned_pool()
{
    static PinnedPool pool;
    return pool;
}

// Type-erased solver object behind the sb_trws_solver handle.
struct SolverBase {
    virtual ~SolverBase() {}
    virtual void reset() = 0;
    virtual void minimize(double maxiter, double max_relgap, double *energy, double *lb, double *iters,
                          sb_trws_timing *timing) = 0;
    virtual void labels(double *out) = 0;
    // row-banded multi-GPU (include/stereo_b200.h, "several GPUs")
    virtual void ipc_export(unsigned char *out) = 0;
    virtual void ipc_attach(const unsigned char *up, const unsigned char *down) = 0;
    virtual void run_one_pass(int pass, int mode, double *acc) = 0;
    double setup_ms = 0;
};

template <typename REAL>
struct Solver : SolverBase {
    int kernel, L, H, W, K, LP, S = 0, precision;
    int64_t N, E;
    bool fuse;
    const KOps *ops;
    cudaStream_t stream = 0;
    DevBuf<REAL> dD, dMsg, dPosQ, dPosQp, dAlpha;
    DevBuf<uint8_t> dRankQ, dRankQp, dCntQ, dCntQp;
    std::shared_ptr<GridPlan> plan;
    DevBuf<REAL> dSelPos;
    DevBuf<unsigned long long> dMbox, dSelBox;
    int rank = 0, world = 1;
    // world > 1: message / mailbox arrays come from cudaMalloc so they can be opened by the
    // neighbouring ranks through CUDA IPC
    REAL *ipc_msg = nullptr;
    unsigned long long *ipc_mbox = nullptr, *ipc_selbox = nullptr;
    void *peer_ptr[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    DevBuf<int32_t> dSol;
    DevBuf<unsigned char> dCtrl;   // Ctrl + progress[S]
    DevBuf<long long> dProf;
    Problem<REAL> P;
    int grid_fwd = 1, grid_bwd = 1, wpb = 1, epoch = 0, nhw = 2;
    int *rec_host = nullptr;     // SB_TRWS_RECORD flight recorder (host-mapped)
    int rec_ctas = 0;
    unsigned launch_epoch = 0;   // never reset: tags of earlier launches must not validate
    Ctrl *hc = nullptr; // pinned
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double kernel_ms = 0;
    int64_t kernel_count = 0;

    Solver(int kernel_, int L_, int64_t N_, int64_t E_, int H_, int W_, const double *unary, const double *q,
           const double *qprim, const double *alphas, double tol, const sb_trws_options &opt, int rank_ = 0,
           int world_ = 1)
        : kernel(kernel_), L(L_), H(H_), W(W_), N(N_), E(E_), rank(rank_), world(world_)
    {
        // a constructor that throws does not run the destructor: release what was acquired (pinned slot, events,
        // IPC allocations) before the rejection ("q contains NaN", a CUDA error) reaches the caller
        try {
            init(unary, q, qprim, alphas, tol, opt);
        } catch (...) {
            release();
            throw;
        }
    }

    void init(const double *unary, const double *q, const double *qprim, const double *alphas, double tol, const sb_trws_options &opt)
    {
        SB_REQUIRE(world == 1 || sizeof(REAL) == 4, SB_EUNSUP, "row-banded sweeps run in fp32 only");
        precision = sizeof(REAL) == 8 ? SB_F64 : SB_F32;
        fuse = opt.fuse_rounding != 0;
        ops = kops_for_labels(L);
        SB_REQUIRE(ops, SB_EUNSUP, "sb_trws_solve: %d labels exceed SB_MAX_LABELS=%d", L, SB_MAX_LABELS);
        K = ops->K;
        LP = 32 * K;
        const double t0 = now_ms();
        int dev = 0, num_sms = 0;
        SB_CUDA(cudaGetDevice(&dev));
        SB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));

        // ---- host graph logic: ordering + dispatch schedule (cached per grid shape)
        plan = grid_plan(dev, H, W, rank, world);
        S = plan->S;

        // ---- device state
        dD.alloc((size_t)N * LP); dPosQ.alloc((size_t)E * LP); dPosQp.alloc((size_t)E * LP);
        if (world == 1) dMsg.alloc((size_t)E * LP);
        dAlpha.alloc((size_t)E);
        dRankQ.alloc((size_t)E * LP); dRankQp.alloc((size_t)E * LP); dCntQ.alloc((size_t)E * LP); dCntQp.alloc((size_t)E * LP);
        dSol.alloc((size_t)N);
        dSelPos.alloc((size_t)std::max<int64_t>(E, 1));
        const size_t mbox_words = (size_t)std::max<int64_t>(E, 1) * LP, sel_words = (size_t)std::max<int64_t>(E, 1);
        if (world > 1) {
            SB_CUDA(cudaMalloc((void **)&ipc_msg, (size_t)E * LP * sizeof(REAL)));
            SB_CUDA(cudaMalloc((void **)&ipc_mbox, mbox_words * 8));
            SB_CUDA(cudaMalloc((void **)&ipc_selbox, sel_words * 8));
            SB_CUDA(cudaMemsetAsync(ipc_mbox, 0, mbox_words * 8, stream));
            SB_CUDA(cudaMemsetAsync(ipc_selbox, 0, sel_words * 8, stream));
        } else if (sizeof(REAL) == 4) {
            dMbox.alloc(mbox_words);
            dSelBox.alloc(sel_words);
            SB_CUDA(cudaMemsetAsync(dMbox.p, 0, dMbox.bytes(), stream));
            SB_CUDA(cudaMemsetAsync(dSelBox.p, 0, dSelBox.bytes(), stream));
        }
        dCtrl.alloc(sizeof(Ctrl) + (size_t)plan->S_global * 4);
        DevBuf<int> dBad(1);
        hc = static_cast<Ctrl *>(pinned_pool().get(sizeof(Ctrl)));
        SB_CUDA(cudaEventCreate(&ev0));
        SB_CUDA(cudaEventCreate(&ev1));

        SB_CUDA(cudaMemsetAsync(dSelPos.p, 0, dSelPos.bytes(), stream));
        SB_CUDA(cudaMemsetAsync(dBad.p, 0, sizeof(int), stream));
        {
            // raw doubles -> padded REAL arrays + rank / merge-count tables
            DevBuf<double> raw((size_t)N * L);
            SB_CUDA(cudaMemcpyAsync(raw.p, unary, raw.bytes(), cudaMemcpyHostToDevice, stream));
            const long long tot = (long long)N * LP;
            convert_unary_kernel<REAL><<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(raw.p, dD.p, L, LP, N);
            SB_CUDA(cudaGetLastError());
            count_launch();
            SB_CUDA(cudaStreamSynchronize(stream));
        }
        if (E > 0) {
            DevBuf<double> rawa((size_t)E);
            SB_CUDA(cudaMemcpyAsync(rawa.p, alphas, rawa.bytes(), cudaMemcpyHostToDevice, stream));
            convert_vec_kernel<REAL><<<(unsigned)((E + 255) / 256), 256, 0, stream>>>(rawa.p, dAlpha.p, E);
            SB_CUDA(cudaGetLastError());
            count_launch();
            // q / qprim are uploaded and tabulated in slabs (the raw doubles never need 2*8*L*E bytes
            // on the device), alternating between two streams so the copy of one slab overlaps the
            // table build of the other
            const int64_t slab = std::max<int64_t>(1, std::min<int64_t>(E, (int64_t)(128u << 20) / (8 * (int64_t)L)));
            DevBuf<double> rq[2], rqp[2];
            struct SlabStreams {     // destroyed on every way out of this block, a throwing table build included
                cudaStream_t s[2] = {nullptr, nullptr};
                cudaStream_t &operator[](int b) { return s[b]; }
                ~SlabStreams()
                {
                    for (int b = 0; b < 2; b++)
                        if (s[b]) { cudaStreamSynchronize(s[b]); cudaStreamDestroy(s[b]); }
                }
            } ss;
            for (int b = 0; b < 2; b++) {
                rq[b].alloc((size_t)slab * L);
                rqp[b].alloc((size_t)slab * L);
                SB_CUDA(cudaStreamCreateWithFlags(&ss[b], cudaStreamNonBlocking));
            }
            SB_CUDA(cudaStreamSynchronize(stream));   // buffers and zero-fills above are ready
            int which = 0;
            for (int64_t e0 = 0; e0 < E; e0 += slab, which ^= 1) {
                const int64_t ne = std::min<int64_t>(slab, E - e0);
                SB_CUDA(cudaMemcpyAsync(rq[which].p, q + e0 * L, (size_t)ne * L * 8, cudaMemcpyHostToDevice, ss[which]));
                SB_CUDA(cudaMemcpyAsync(rqp[which].p, qprim + e0 * L, (size_t)ne * L * 8, cudaMemcpyHostToDevice, ss[which]));
                TablesLaunch tl;
                tl.precision = precision;
                tl.q = rq[which].p; tl.qp = rqp[which].p; tl.L = L; tl.E = ne;
                tl.posq = dPosQ.p + e0 * LP; tl.posqp = dPosQp.p + e0 * LP;
                tl.rank_q = dRankQ.p + e0 * LP; tl.rank_qp = dRankQp.p + e0 * LP;
                tl.cnt_q = dCntQ.p + e0 * LP; tl.cnt_qp = dCntQp.p + e0 * LP;
                tl.bad = dBad.p; tl.stream = ss[which];
                ops->tables(tl);
            }
            for (int b = 0; b < 2; b++) SB_CUDA(cudaStreamSynchronize(ss[b]));
            int bad = 0;
            SB_CUDA(cudaMemcpyAsync(&bad, dBad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
            SB_CUDA(cudaStreamSynchronize(stream));
            SB_REQUIRE(!(bad & 1), SB_EINVAL, "q contains NaN");        // trws.m:9-11
            SB_REQUIRE(!(bad & 2), SB_EINVAL, "qprim contains NaN");    // trws.m:13-15
        }

        std::memset(&P, 0, sizeof(P));
        P.H = H; P.W = W; P.L = L; P.LP = LP; P.N = N; P.E = E;
        P.nV = (long long)(H - 1) * W; P.nH = (long long)H * (W - 1);
        P.D = dD.p; P.msg = world > 1 ? ipc_msg : dMsg.p; P.posq = dPosQ.p; P.posqp = dPosQp.p;
        P.rank_q = dRankQ.p; P.rank_qp = dRankQp.p; P.cnt_q = dCntQ.p; P.cnt_qp = dCntQp.p;
        P.alpha = dAlpha.p; P.lambda = (REAL)tol;
        P.strip_len = plan->strip_len.p; P.strip_gid = plan->strip_gid.p; P.S = S;
        P.sol = dSol.p; P.selpos = dSelPos.p;
        P.mbox = world > 1 ? ipc_mbox : dMbox.p; P.selbox = world > 1 ? ipc_selbox : dSelBox.p;
        P.world = world;
        Ctrl *ctrl = reinterpret_cast<Ctrl *>(dCtrl.p);
        P.ticket = &ctrl->ticket; P.acc = ctrl->acc;
        P.progress = reinterpret_cast<int32_t *>(dCtrl.p + sizeof(Ctrl));

        if (const char *dbg = getenv("SB_TRWS_DEBUG")) P.debug = atoi(dbg);
        if (getenv("SB_TRWS_RECORD")) {
            rec_ctas = 1024;
            SB_CUDA(cudaHostAlloc((void **)&rec_host, (size_t)rec_ctas * 8 * 4 * sizeof(int), cudaHostAllocMapped));
            std::memset(rec_host, 0xff, (size_t)rec_ctas * 8 * 4 * sizeof(int));
            SB_CUDA(cudaHostGetDevicePointer((void **)&P.rec, rec_host, 0));
        }
        if (getenv("SB_TRWS_PROFILE") && !TDIAG)
            fprintf(stderr, "[stereo_b200] SB_TRWS_PROFILE needs the diagnostics build: make -C stereo_b200/csrc clean all EXTRA=-DSB_TRWS_DIAG=1\n");
        if (getenv("SB_TRWS_PROFILE")) {
            dProf.alloc(64);
            SB_CUDA(cudaMemsetAsync(dProf.p, 0, 64 * sizeof(long long), stream));
            P.prof = dProf.p;
        }

        wpb = ops->sweep_warps_per_block();
        nhw = 2;   // helper warps per CTA (four were measured slower: DESIGN.md)
        auto grid_for = [&](int pass) {
            const int bps = ops->sweep_blocks_per_sm(precision, kernel, pass, nhw);
            SB_REQUIRE(bps >= 1, SB_ECUDA, "sb_trws_solve: sweep kernel does not fit on an SM");
            long long g = (long long)bps * num_sms;
            if (g > S) g = S;
            // two strip walkers must be in flight: (H-3,1) waits for (H-2,1) (trws_order.cpp)
            SB_REQUIRE(S < 2 || g >= 2, SB_ECUDA, "sb_trws_solve: fewer than two resident CTAs");
            if (g < 1) g = 1;
            return (int)(g < 1 ? 1 : g);
        };
        grid_fwd = grid_for(PASS_FWD);
        grid_bwd = grid_for(PASS_BWD);
        reset();
        SB_CUDA(cudaStreamSynchronize(stream));
        setup_ms = now_ms() - t0;
    }

    void release()
    {
        if (hc) pinned_pool().put(hc);
        hc = nullptr;
        for (int d = 0; d < 2; d++)
            for (int a = 0; a < 3; a++) {
                if (peer_ptr[d][a]) cudaIpcCloseMemHandle(peer_ptr[d][a]);
                peer_ptr[d][a] = nullptr;
            }
        if (ipc_msg) cudaFree(ipc_msg);
        if (ipc_mbox) cudaFree(ipc_mbox);
        if (ipc_selbox) cudaFree(ipc_selbox);
        ipc_msg = nullptr; ipc_mbox = nullptr; ipc_selbox = nullptr;
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        ev0 = nullptr; ev1 = nullptr;
    }

    ~Solver() override { release(); }

    // ZeroMessages (MRFEnergy.cpp:115-131) + fresh epoch flags
    void reset() override
    {
        SB_CUDA(cudaMemsetAsync(dSol.p, 0, (size_t)N * 4, stream));
        SB_CUDA(cudaMemsetAsync(P.msg, 0, (size_t)E * LP * sizeof(REAL), stream));
        epoch = 0;
    }

    void run_pass(int pass, int mode)
    {
        SB_CUDA(cudaMemsetAsync(dCtrl.p, 0, dCtrl.bytes(), stream));
        ++epoch;
        P.epoch = ++launch_epoch;
        P.mode = mode;
        P.segs = plan->segs[pass == PASS_FWD ? 0 : 1].p;
        P.seg_ptr = plan->seg_ptr[pass == PASS_FWD ? 0 : 1].p;
        SweepLaunch sl;
        sl.precision = precision; sl.kern = kernel; sl.pass = pass; sl.problem = &P;
        sl.grid = pass == PASS_FWD ? grid_fwd : grid_bwd; sl.stream = stream; sl.nhw = nhw;
        if (rec_host) std::memset(rec_host, 0xff, (size_t)rec_ctas * 8 * 4 * sizeof(int));
        SB_CUDA(cudaEventRecord(ev0, stream));
        ops->sweep(sl);
        SB_CUDA(cudaEventRecord(ev1, stream));
        SB_CUDA(cudaMemcpyAsync(hc, dCtrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
        if (rec_host) {
            // flight recorder: wait with a deadline, dump where every warp was on a hang or a fault
            const double t_start = now_ms();
            cudaError_t q;
            while ((q = cudaStreamQuery(stream)) == cudaErrorNotReady && now_ms() - t_start < 4000.0) {}
            if (q != cudaSuccess) {
                fprintf(stderr, "[sb record] sweep pass=%d mode=%d epoch=%u grid=%d: %s\n", pass, mode, P.epoch, sl.grid,
                        q == cudaErrorNotReady ? "HANG" : cudaGetErrorString(q));
                fprintf(stderr, "[sb record] D=%p msg=%p posq=%p posqp=%p N=%lld E=%lld LP=%d\n", (void *)P.D, (void *)P.msg,
                        (void *)P.posq, (void *)P.posqp, (long long)N, (long long)E, LP);
                for (int c = 0; c < sl.grid && c < rec_ctas; c++) {
                    fprintf(stderr, "[sb record] cta %d:", c);
                    for (int w = 0; w < 8; w++) {
                        const int *r = rec_host + ((size_t)c * 8 + w) * 4;
                        fprintf(stderr, " w%d(s%d n%d st%d t%d)", w, r[0], r[1], r[2], r[3]);
                    }
                    fprintf(stderr, "\n");
                }
                fflush(stderr);
                _exit(3);
            }
        }
        SB_CUDA(cudaStreamSynchronize(stream));
        float ms = 0;
        SB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        kernel_ms += ms;
        kernel_count++;
    }

    void ipc_export(unsigned char *out) override
    {
        SB_REQUIRE(world > 1, SB_EINVAL, "sb_trws_ipc_export: the solver was not created for several ranks");
        cudaIpcMemHandle_t h[3];
        SB_CUDA(cudaIpcGetMemHandle(&h[0], ipc_msg));
        SB_CUDA(cudaIpcGetMemHandle(&h[1], ipc_mbox));
        SB_CUDA(cudaIpcGetMemHandle(&h[2], ipc_selbox));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        std::memcpy(out, h, sizeof(h));
    }

    void ipc_attach(const unsigned char *up, const unsigned char *down) override
    {
        SB_REQUIRE(world > 1, SB_EINVAL, "sb_trws_ipc_attach: the solver was not created for several ranks");
        const unsigned char *src[2] = {up, down};
        for (int d = 0; d < 2; d++) {
            if (!src[d]) continue;
            cudaIpcMemHandle_t h[3];
            std::memcpy(h, src[d], sizeof(h));
            for (int a = 0; a < 3; a++)
                SB_CUDA(cudaIpcOpenMemHandle(&peer_ptr[d][a], h[a], cudaIpcMemLazyEnablePeerAccess));
            P.peer_msg[d] = static_cast<REAL *>(peer_ptr[d][0]);
            P.peer_mbox[d] = static_cast<unsigned long long *>(peer_ptr[d][1]);
            P.peer_selbox[d] = static_cast<unsigned long long *>(peer_ptr[d][2]);
        }
    }

    void run_one_pass(int pass, int mode, double *acc) override
    {
        run_pass(pass == 0 ? PASS_FWD : PASS_BWD, mode);
        acc[0] = hc->acc[0];
        acc[1] = hc->acc[1];
    }

    // minimize.cpp:31-113.  With fused rounding the energy of iteration t becomes
    // known inside the forward sweep of t+1; the outputs (labels, energy, bound,
    // count) are those of iteration t either way.
    void minimize(double maxiter, double max_relgap, double *energy_out, double *lb_out, double *iters_out,
                  sb_trws_timing *timing) override
    {
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
            long long h[64];
            SB_CUDA(cudaMemcpy(h, dProf.p, sizeof(h), cudaMemcpyDeviceToHost));
            SB_CUDA(cudaMemsetAsync(dProf.p, 0, sizeof(h), stream));
            for (int g = 0; g < 2; g++) {
                const long long *q = h + 16 * g;
                const double nt = q[3] ? (double)q[3] : 1, nh = q[11] ? (double)q[11] : 1, np = q[13] ? (double)q[13] : 1;
                fprintf(stderr, "[sb profile] %s term0 detail: half1=%.0f prefetch-issue=%.0f update+stores=%.0f loop-tail=%.0f cyc/node\n",
                        g ? "rows" : "ring", h[32 + 8 * g] / nt, h[33 + 8 * g] / nt, h[34 + 8 * g] / nt, q[2] / nt);
                fprintf(stderr, "[sb profile] %s term0: waitFULL=%.0f read+round=%.0f update=%.0f cyc/node (%lld nodes) | helper0: desc=%.0f "
                        "static=%.0f flagspin=%.0f dynwait=%.0f reduce=%.0f waitEMPTY=%.0f other=%.0f cyc/node (%lld) | aux: fence=%.0f cyc/publish, "
                        "%.2f nodes/publish\n", g ? "rows" : "ring", q[0] / nt, q[1] / nt, q[2] / nt, q[3], q[4] / nh, q[5] / nh, q[6] / nh,
                        q[7] / nh, q[8] / nh, q[9] / nh, q[10] / nh, q[11], q[12] / np, q[14] / np);
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
        std::vector<int32_t> sol((size_t)N);
        SB_CUDA(cudaMemcpy(sol.data(), dSol.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
        for (int64_t u = 0; u < N; u++) out[u] = (double)(sol[u] + 1); // trws_mex.cpp:138
    }
};

} // namespace

double now_ms_public() { return now_ms(); }

} // namespace trws
} // namespace sb

extern "C" {

void sb_trws_default_options(sb_trws_options *opt)
{
    if (!opt) return;
    std::memset(opt, 0, sizeof(*opt));
    opt->maxiter = 1000;  // trws_mex.cpp:39
    opt->max_relgap = 0;  // trws_mex.cpp:40
    opt->precision = SB_F32;
    opt->fuse_rounding = 1;
}

struct sb_trws_solver {
    sb::trws::SolverBase *impl;
};

static int create_impl(int kernel, int L, int64_t N, int64_t E, const double *unary, const uint32_t *conn,
                       const double *q, const double *qprim, const double *alphas, double tol,
                       const sb_trws_options *opt_in, int rank, int world, sb_trws_solver **out)
{
    return sb::guarded([&] {
        SB_REQUIRE(world >= 1 && rank >= 0 && rank < world, SB_EINVAL, "sb_trws_create: bad rank / world");
        SB_REQUIRE(out, SB_EINVAL, "sb_trws_create: null output");
        *out = nullptr;
        // trws_mex.cpp:156-163
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unsupported kernel");
        // trws_mex.cpp:42-52
        SB_REQUIRE(L >= 1 && N >= 1 && E >= 0, SB_EINVAL, "sb_trws_solve: bad sizes L=%d N=%lld E=%lld", L,
                   (long long)N, (long long)E);
        SB_REQUIRE(unary, SB_EINVAL, "sb_trws_solve: null pointer");
        SB_REQUIRE(E == 0 || (conn && q && qprim && alphas), SB_EINVAL, "sb_trws_solve: null pointer");
        SB_REQUIRE(N < (1LL << 31), SB_EUNSUP, "sb_trws_solve: too many nodes");
        sb_trws_options opt;
        if (opt_in) opt = *opt_in; else sb_trws_default_options(&opt);
        SB_REQUIRE(opt.precision == SB_F32 || opt.precision == SB_F64, SB_EINVAL, "sb_trws_solve: bad precision");
        int H = 0, W = 0;
        SB_REQUIRE(sb::grid_from_connectivity(N, E, conn, H, W), SB_ENOTGRID,
                   "sb_trws_solve: connectivity (N=%lld, E=%lld) is not the 4-connected dispmap_super grid; "
                   "general graphs are not supported on the GPU path", (long long)N, (long long)E);
        sb::require_device();
        sb::trws::SolverBase *impl;
        if (opt.precision == SB_F64)
            impl = new sb::trws::Solver<double>(kernel, L, N, E, H, W, unary, q, qprim, alphas, tol, opt, rank, world);
        else
            impl = new sb::trws::Solver<float>(kernel, L, N, E, H, W, unary, q, qprim, alphas, tol, opt, rank, world);
        *out = new sb_trws_solver{impl};
    });
}

int sb_trws_create(int kernel, int L, int64_t N, int64_t E, const double *unary, const uint32_t *conn,
                   const double *q, const double *qprim, const double *alphas, double tol,
                   const sb_trws_options *opt_in, sb_trws_solver **out)
{
    return create_impl(kernel, L, N, E, unary, conn, q, qprim, alphas, tol, opt_in, 0, 1, out);
}

int sb_trws_create_banded(int kernel, int L, int64_t N, int64_t E, const double *unary, const uint32_t *conn,
                          const double *q, const double *qprim, const double *alphas, double tol,
                          const sb_trws_options *opt_in, int rank, int world, sb_trws_solver **out)
{
    return create_impl(kernel, L, N, E, unary, conn, q, qprim, alphas, tol, opt_in, rank, world, out);
}

int sb_trws_ipc_export(sb_trws_solver *s, unsigned char *handles)
{
    return sb::guarded([&] {
        SB_REQUIRE(s && s->impl && handles, SB_EINVAL, "sb_trws_ipc_export: null pointer");
        s->impl->ipc_export(handles);
    });
}

int sb_trws_ipc_attach(sb_trws_solver *s, const unsigned char *up, const unsigned char *down)
{
    return sb::guarded([&] {
        SB_REQUIRE(s && s->impl, SB_EINVAL, "sb_trws_ipc_attach: null solver");
        s->impl->ipc_attach(up, down);
    });
}

int sb_trws_pass(sb_trws_solver *s, int pass, int mode, double *acc)
{
    return sb::guarded([&] {
        SB_REQUIRE(s && s->impl && acc, SB_EINVAL, "sb_trws_pass: null pointer");
        SB_REQUIRE((pass == 0 || pass == 1) && mode >= 0 && mode <= 3, SB_EINVAL, "sb_trws_pass: bad pass / mode");
        s->impl->run_one_pass(pass, mode, acc);
    });
}

int sb_trws_reset(sb_trws_solver *s)
{
    return sb::guarded([&] {
        SB_REQUIRE(s && s->impl, SB_EINVAL, "sb_trws_reset: null solver");
        s->impl->reset();
    });
}

int sb_trws_minimize(sb_trws_solver *s, double maxiter, double max_relgap, double *energy, double *lower_bound,
                     double *iterations, sb_trws_timing *timing)
{
    return sb::guarded([&] {
        SB_REQUIRE(s && s->impl, SB_EINVAL, "sb_trws_minimize: null solver");
        SB_REQUIRE(energy && lower_bound && iterations, SB_EINVAL, "sb_trws_minimize: null pointer");
        s->impl->minimize(maxiter, max_relgap, energy, lower_bound, iterations, timing);
    });
}

int sb_trws_get_labels(sb_trws_solver *s, double *labels)
{
    return sb::guarded([&] {
        SB_REQUIRE(s && s->impl && labels, SB_EINVAL, "sb_trws_get_labels: null pointer");
        s->impl->labels(labels);
    });
}

void sb_trws_destroy(sb_trws_solver *s)
{
    if (!s) return;
    delete s->impl;
    delete s;
}

int sb_trws_solve(int kernel, int L, int64_t N, int64_t E, const double *unary, const uint32_t *conn,
                  const double *q, const double *qprim, const double *alphas, double tol,
                  const sb_trws_options *opt_in, double *labels, double *energy, double *lower_bound,
                  double *iterations, sb_trws_timing *timing)
{
    sb_trws_solver *s = nullptr;
    if (!labels || !energy || !lower_bound || !iterations) {
        sb::set_last_error("sb_trws_solve: null pointer");
        return SB_EINVAL;
    }
    int rc = sb_trws_create(kernel, L, N, E, unary, conn, q, qprim, alphas, tol, opt_in, &s);
    if (rc != SB_OK) return rc;
    sb_trws_options opt;
    if (opt_in) opt = *opt_in; else sb_trws_default_options(&opt);
    rc = sb_trws_minimize(s, opt.maxiter, opt.max_relgap, energy, lower_bound, iterations, timing);
    if (rc == SB_OK) {
        const double t0 = sb::trws::now_ms_public();
        rc = sb_trws_get_labels(s, labels);
        if (timing) timing->download_ms = sb::trws::now_ms_public() - t0;
    }
    sb_trws_destroy(s);
    return rc;
}

} // extern "C"
