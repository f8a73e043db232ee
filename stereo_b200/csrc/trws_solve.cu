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
#include <chrono>
#include <cstring>
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
};

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

template <typename REAL>
void solve_typed(int kernel, int L, int64_t N, int64_t E, int H, int W, const double *unary, const double *q,
                 const double *qprim, const double *alphas, double tol, const sb_trws_options &opt,
                 double *labels, double *energy_out, double *lb_out, double *iters_out, sb_trws_timing *timing)
{
    const int precision = sizeof(REAL) == 8 ? SB_F64 : SB_F32;
    const KOps *ops = kops_for_labels(L);
    SB_REQUIRE(ops, SB_EUNSUP, "sb_trws_solve: %d labels exceed SB_MAX_LABELS=%d", L, SB_MAX_LABELS);
    const int K = ops->K, LP = 32 * K;
    const int64_t launches0 = g_launches.load();
    const double t_setup0 = now_ms();

    cudaStream_t stream = 0;
    int dev = 0, num_sms = 0;
    SB_CUDA(cudaGetDevice(&dev));
    SB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));

    // ---- host graph logic: ordering + dispatch schedule
    std::vector<int32_t> order;
    SB_REQUIRE(grid_ordering(H, W, order), SB_EINVAL,
               "sb_trws_solve: %dx%d grid has no valid automatic ordering (the reference crashes on it)", H, W);
    std::vector<uint8_t> info;
    build_node_info(H, W, order, info);
    Schedule sched;
    build_schedule(H, W, order, sched);
    const int S = (int)sched.strip_ptr.size() - 1;
    std::vector<int32_t> strip_ptr32(sched.strip_ptr.begin(), sched.strip_ptr.end());

    // ---- device state
    DevBuf<REAL> dD((size_t)N * LP), dMsg((size_t)E * LP), dPosQ((size_t)E * LP), dPosQp((size_t)E * LP), dAlpha((size_t)E);
    DevBuf<uint8_t> dRankQ((size_t)E * LP), dRankQp((size_t)E * LP), dCntQ((size_t)E * LP), dCntQp((size_t)E * LP);
    DevBuf<int32_t> dNodes((size_t)N), dStripPtr((size_t)S + 1), dDone((size_t)N), dSol((size_t)N);
    DevBuf<uint8_t> dInfo((size_t)N);
    DevBuf<Ctrl> dCtrl(1);
    DevBuf<int> dBad(1);

    SB_CUDA(cudaMemcpyAsync(dNodes.p, sched.nodes.data(), (size_t)N * 4, cudaMemcpyHostToDevice, stream));
    SB_CUDA(cudaMemcpyAsync(dStripPtr.p, strip_ptr32.data(), ((size_t)S + 1) * 4, cudaMemcpyHostToDevice, stream));
    SB_CUDA(cudaMemcpyAsync(dInfo.p, info.data(), (size_t)N, cudaMemcpyHostToDevice, stream));
    SB_CUDA(cudaMemsetAsync(dDone.p, 0, (size_t)N * 4, stream));
    SB_CUDA(cudaMemsetAsync(dSol.p, 0, (size_t)N * 4, stream));
    SB_CUDA(cudaMemsetAsync(dMsg.p, 0, dMsg.bytes(), stream)); // ZeroMessages, MRFEnergy.cpp:115-131
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
        DevBuf<double> rq((size_t)E * L), rqp((size_t)E * L);
        SB_CUDA(cudaMemcpyAsync(rq.p, q, rq.bytes(), cudaMemcpyHostToDevice, stream));
        SB_CUDA(cudaMemcpyAsync(rqp.p, qprim, rqp.bytes(), cudaMemcpyHostToDevice, stream));
        TablesLaunch tl;
        tl.precision = precision;
        tl.q = rq.p; tl.qp = rqp.p; tl.L = L; tl.E = E;
        tl.posq = dPosQ.p; tl.posqp = dPosQp.p;
        tl.rank_q = dRankQ.p; tl.rank_qp = dRankQp.p; tl.cnt_q = dCntQ.p; tl.cnt_qp = dCntQp.p;
        tl.bad = dBad.p; tl.stream = stream;
        ops->tables(tl);
        int bad = 0;
        SB_CUDA(cudaMemcpyAsync(&bad, dBad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
        SB_REQUIRE(!bad, SB_EINVAL, "sb_trws_solve: q or qprim contains NaN (trws.m:9-15)");
    }

    Problem<REAL> P;
    std::memset(&P, 0, sizeof(P));
    P.H = H; P.W = W; P.L = L; P.LP = LP; P.N = N; P.E = E;
    P.nV = (long long)(H - 1) * W; P.nH = (long long)H * (W - 1);
    P.D = dD.p; P.msg = dMsg.p; P.posq = dPosQ.p; P.posqp = dPosQp.p;
    P.rank_q = dRankQ.p; P.rank_qp = dRankQp.p; P.cnt_q = dCntQ.p; P.cnt_qp = dCntQp.p;
    P.alpha = dAlpha.p; P.lambda = (REAL)tol;
    P.info = dInfo.p; P.nodes = dNodes.p; P.strip_ptr = dStripPtr.p; P.S = S;
    P.done = dDone.p; P.sol = dSol.p;
    P.ticket = &dCtrl.p->ticket; P.acc = dCtrl.p->acc;

    const int wpb = ops->sweep_warps_per_block();
    auto grid_for = [&](int pass) {
        const int bps = ops->sweep_blocks_per_sm(precision, kernel, pass);
        SB_REQUIRE(bps >= 1, SB_ECUDA, "sb_trws_solve: sweep kernel does not fit on an SM");
        long long g = (long long)bps * num_sms;
        if (g > S) g = S;
        return (int)(g < 1 ? 1 : g);
    };
    const int grid_fwd = grid_for(PASS_FWD), grid_bwd = grid_for(PASS_BWD);
    auto active_for = [&](int grid) {
        // at least two strip-walking warps in flight (trws_order.cpp: row H-3 waits for row H-2)
        long long a = ((long long)S + grid - 1) / grid;
        if ((long long)grid * a < 2) a = 2;
        return (int)std::min<long long>(a, wpb);
    };

    int epoch = 0;
    Ctrl hc;
    auto run_pass = [&](int pass, int mode) {
        SB_CUDA(cudaMemsetAsync(dCtrl.p, 0, sizeof(Ctrl), stream));
        P.epoch = ++epoch;
        P.mode = mode;
        SweepLaunch sl;
        sl.precision = precision; sl.kern = kernel; sl.pass = pass; sl.problem = &P;
        sl.grid = pass == PASS_FWD ? grid_fwd : grid_bwd; sl.stream = stream;
        P.active_warps = active_for(sl.grid);
        ops->sweep(sl);
        SB_CUDA(cudaMemcpyAsync(&hc, dCtrl.p, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
        SB_CUDA(cudaStreamSynchronize(stream));
    };

    const double t_setup1 = now_ms();
    EventTimer timer(stream);
    timer.start();

    // minimize.cpp:31-113.  With fused rounding the energy of iteration t becomes
    // known inside the forward sweep of t+1; the outputs (labels, energy, bound,
    // count) are those of iteration t either way.
    const int iter_max = (int)opt.maxiter;
    const double relgap_max = opt.max_relgap;
    const bool fuse = opt.fuse_rounding != 0;
    double energy = 0, lb = 0;
    int iterations = 0;
    for (int it = 1;; it++) {
        run_pass(PASS_FWD, MODE_SEND | ((fuse && it > 1) ? MODE_ROUND : 0));
        if (fuse && it > 1) {
            energy = hc.acc[0];
            if ((energy - lb) / energy < relgap_max) { iterations = it - 1; break; }
        }
        run_pass(PASS_BWD, 0);
        lb = hc.acc[1];
        if (!fuse || it >= iter_max) {
            run_pass(PASS_FWD, MODE_ROUND);
            energy = hc.acc[0];
            if (it >= iter_max || (energy - lb) / energy < relgap_max) { iterations = it; break; }
        }
    }
    const double solve_ms = timer.stop_ms();

    const double t_dl0 = now_ms();
    std::vector<int32_t> sol((size_t)N);
    SB_CUDA(cudaMemcpy(sol.data(), dSol.p, (size_t)N * 4, cudaMemcpyDeviceToHost));
    for (int64_t u = 0; u < N; u++) labels[u] = (double)(sol[u] + 1); // trws_mex.cpp:138
    *energy_out = energy;
    *lb_out = lb;
    *iters_out = (double)iterations;
    if (timing) {
        std::memset(timing, 0, sizeof(*timing));
        timing->setup_ms = t_setup1 - t_setup0;
        timing->solve_ms = solve_ms;
        timing->sweep_ms_avg = iterations ? solve_ms / iterations : 0;
        timing->download_ms = now_ms() - t_dl0;
        timing->kernel_launches = g_launches.load() - launches0;
    }
}

} // namespace
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

int sb_trws_solve(int kernel, int L, int64_t N, int64_t E, const double *unary, const uint32_t *conn,
                  const double *q, const double *qprim, const double *alphas, double tol,
                  const sb_trws_options *opt_in, double *labels, double *energy, double *lower_bound,
                  double *iterations, sb_trws_timing *timing)
{
    return sb::guarded([&] {
        // trws_mex.cpp:156-163
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unsupported kernel");
        // trws_mex.cpp:42-52
        SB_REQUIRE(L >= 1 && N >= 1 && E >= 0, SB_EINVAL, "sb_trws_solve: bad sizes L=%d N=%lld E=%lld", L,
                   (long long)N, (long long)E);
        SB_REQUIRE(unary && labels && energy && lower_bound && iterations, SB_EINVAL, "sb_trws_solve: null pointer");
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
        if (opt.precision == SB_F64)
            sb::trws::solve_typed<double>(kernel, L, N, E, H, W, unary, q, qprim, alphas, tol, opt, labels, energy,
                                          lower_bound, iterations, timing);
        else
            sb::trws::solve_typed<float>(kernel, L, N, E, H, W, unary, q, qprim, alphas, tol, opt, labels, energy,
                                         lower_bound, iterations, timing);
    });
}

} // extern "C"
