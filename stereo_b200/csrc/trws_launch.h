// trws_launch.h -- type-erased launch interface between the TRW-S host driver
// (trws_solve.cu) and the per-K kernel instantiation units (trws_inst_k*.cu).
// K = LP/32 labels per lane is a compile-time constant of the sweep kernels;
// each supported K lives in its own translation unit so they build in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {
namespace trws {

struct SweepLaunch {
    int precision;       // SB_F32 / SB_F64
    int kern;            // 1 linear, 2 quadratic
    int pass;            // PASS_FWD / PASS_BWD
    const void *problem; // Problem<float> or Problem<double>, host copy
    int grid;            // persistent CTAs (all co-resident)
    int nhw;             // helper warps per CTA: 2 or 4 (fp32 only)
    cudaStream_t stream;
};

struct TablesLaunch {
    int precision;
    const double *q, *qp; // raw L x E doubles on the device
    int L;
    long long E;
    void *posq, *posqp;
    uint8_t *rank_q, *rank_qp, *cnt_q, *cnt_qp;
    int *bad;
    cudaStream_t stream;
};

struct KOps {
    int K;
    // persistent CTAs per SM the sweep kernel can keep resident (occupancy query)
    int (*sweep_blocks_per_sm)(int precision, int kern, int pass, int nhw);
    int (*sweep_warps_per_block)();
    void (*sweep)(const SweepLaunch &);
    void (*tables)(const TablesLaunch &);
};

// Smallest supported K with 32*K >= L (nullptr if L > SB_MAX_LABELS).
const KOps *kops_for_labels(int L);

extern const KOps kops_k1, kops_k2, kops_k3, kops_k4, kops_k6, kops_k8;

} // namespace trws
} // namespace sb
