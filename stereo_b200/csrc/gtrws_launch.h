// gtrws_launch.h -- type-erased launch interface between the grid-native TRW-S host driver
// (gtrws_solve.cu) and the per-K kernel instantiation units (gtrws_inst_k*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {
namespace gtrws {

struct GSweepLaunch {
    int precision;       // SB_F32 / SB_F64
    int kern;            // 1 linear, 2 quadratic
    int pass;            // PASS_FWD / PASS_BWD
    const void *problem; // GProblem<float> or GProblem<double>, host copy
    int grid;            // persistent CTAs (all co-resident)
    int lat;             // 1: the latency build (at most two CTAs per SM, early operands)
    cudaStream_t stream;
};

struct GTablesLaunch {
    int precision;
    void *nodeF;
    uint8_t *nodeB, *pairB;
    int W, rows, L;
    cudaStream_t stream;
};

struct GUpdateLaunch {
    int precision, kern, L;
    const double *Di, *msg, *src, *dst;   // device
    double alpha, lambda, gamma;
    double *msg_out, *vmin_out;           // device
    cudaStream_t stream;
};

struct GOps {
    int K;
    int (*blocks_per_sm)(int precision, int kern, int pass, int lat);
    void (*sweep)(const GSweepLaunch &);
    void (*tables)(const GTablesLaunch &);   // pad rows, node ranks, pair tables
    size_t (*smem_bytes)(int precision);
    void (*update_message)(const GUpdateLaunch &);
};

const GOps *gops_for_labels(int L);
extern const GOps gops_k1, gops_k2, gops_k3, gops_k4, gops_k6, gops_k8;

} // namespace gtrws
} // namespace sb
