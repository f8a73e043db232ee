// fusion_schedule.cu -- dispmap_super.binary_fuse_until_convergence (dispmap_super.m:85-152) as ONE call over
// device-resident fields: the proposal plane fields and their unary costs are uploaded (or
// handed over as device pointers) once, every fusion move is sb_binary_fusion_grid in device-pointer mode, the accepted
// planes are adopted by a kernel, and only the energy (8 bytes) crosses the host per fusion -- the reference makes a MATLAB
// round trip with 14 N doubles per fusion.
//
// The visiting order is the caller's: `ids` is the array dispmap_super.m:98-101 builds (1:n followed by randi draws,
// repeats and out-of-range entries removed), so the MATLAB side keeps its own random stream.  The loop keeps the
// reference's bookkeeping, including the `iter = iter + 1` at :116 that makes it start at ids(2).
#include "sb_common.h"
#include "../../include/stereo_b200.h"
#include <vector>

namespace sb {
double device_energy(int H, int W, int kernel, const double *d_unary, const double *d_assignment, const double *d_weights,
                     double tol, double d_min, double d_step);

namespace {

// labelling == 1: take the proposal's plane and its unary cost (dispmap_super.m:78-80); *taken counts them
__global__ void adopt_kernel(long long N, const double *__restrict__ labels, const double *__restrict__ proposal,
                             const double *__restrict__ U1, double *__restrict__ assignment, double *__restrict__ unary,
                             unsigned long long *taken)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool take = false;
    if (u < N && labels[u] == 1.0) {
        take = true;
        const double2 a = reinterpret_cast<const double2 *>(proposal)[2 * u], b = reinterpret_cast<const double2 *>(proposal)[2 * u + 1];
        reinterpret_cast<double2 *>(assignment)[2 * u] = a;
        reinterpret_cast<double2 *>(assignment)[2 * u + 1] = b;
        unary[u] = U1[u];
    }
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(taken, (unsigned long long)__popc(m));
}

} // namespace
} // namespace sb

using namespace sb;

extern "C" {

int sb_binary_fuse_until_convergence_grid(int H, int W, int kernel, int n_proposals, const double *proposals,
                                          const double *unaries, double *assignment, double *unary, const double *weights,
                                          double tol, double d_min, double d_step, int improve, int maxiter,
                                          const int32_t *ids, int64_t n_ids, int on_device, double *energies,
                                          int *n_energies, double *stats)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1, SB_EINVAL, "sb_binary_fuse_until_convergence_grid: bad sizes H=%d W=%d", H, W);
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unkown kernel type");   // dispmap_super.m:232-233
        SB_REQUIRE(n_proposals >= 1 && proposals && unaries && assignment && unary && weights && ids && energies && n_energies,
                   SB_EINVAL, "sb_binary_fuse_until_convergence_grid: null pointer");
        SB_REQUIRE(maxiter >= 0 && n_ids >= 0, SB_EINVAL, "sb_binary_fuse_until_convergence_grid: bad iteration arguments");
        for (int64_t i = 0; i < n_ids; i++)
            SB_REQUIRE(ids[i] >= 1 && ids[i] <= n_proposals, SB_EINVAL,
                       "sb_binary_fuse_until_convergence_grid: ids[%lld] = %d is not a proposal number", (long long)i, (int)ids[i]);
        const int64_t N = (int64_t)H * W, E = 2 * ((int64_t)(H - 1) * W + (int64_t)H * (W - 1));
        require_device();
        DevBuf<double> bprop, bun, bcur, bu, bw, dlab((size_t)N);
        DevBuf<unsigned long long> dtaken(1);
        const double *props = proposals, *uns = unaries, *wt = weights;
        double *cur = assignment, *ucur = unary;
        if (!on_device) {
            bprop.alloc((size_t)n_proposals * N * 4); bun.alloc((size_t)n_proposals * N); bcur.alloc((size_t)N * 4); bu.alloc((size_t)N);
            bw.alloc((size_t)std::max<int64_t>(E, 1));
            SB_CUDA(cudaMemcpyAsync(bprop.p, proposals, bprop.bytes(), cudaMemcpyHostToDevice, 0));
            SB_CUDA(cudaMemcpyAsync(bun.p, unaries, bun.bytes(), cudaMemcpyHostToDevice, 0));
            SB_CUDA(cudaMemcpyAsync(bcur.p, assignment, bcur.bytes(), cudaMemcpyHostToDevice, 0));
            SB_CUDA(cudaMemcpyAsync(bu.p, unary, bu.bytes(), cudaMemcpyHostToDevice, 0));
            if (E) SB_CUDA(cudaMemcpyAsync(bw.p, weights, (size_t)E * 8, cudaMemcpyHostToDevice, 0));
            props = bprop.p; uns = bun.p; cur = bcur.p; ucur = bu.p; wt = bw.p;
        }
        // E = energy(self) (dispmap_super.m:104)
        std::vector<double> Es;
        Es.push_back(device_energy(H, W, kernel, ucur, cur, wt, tol, d_min, d_step));
        std::vector<char> visited((size_t)n_proposals, 0);
        int n_visited = 0;
        double fusions = 0, rounds = 0, solve_ms = 0, adopted = 0;
        for (int iter = 1; iter <= maxiter; iter++) {
            // (":110-113 if iter > number_of_random_ids, ids = [ids ids]": number_of_random_ids = 5 maxiter >= iter, never taken)
            const int64_t it1 = (int64_t)iter + 1;          // :116
            if (it1 > n_ids) break;                          // (the reference would index past ids here)
            const int id = ids[it1 - 1] - 1;
            if (visited[(size_t)id]) continue;               // :120-122
            double e = 0, lb = 0, nu = 0, st[4] = {0, 0, 0, 0};
            const int rc = sb_binary_fusion_grid(H, W, kernel, cur, props + (size_t)id * N * 4, ucur, uns + (size_t)id * N, wt, tol,
                                                 d_min, d_step, improve, 1, dlab.p, &e, &lb, &nu, st);
            SB_REQUIRE(rc == SB_OK, rc, "%s", sb_last_error());
            SB_CUDA(cudaMemsetAsync(dtaken.p, 0, sizeof(unsigned long long), 0));
            adopt_kernel<<<(unsigned)((N + 255) / 256), 256>>>(N, dlab.p, props + (size_t)id * N * 4, uns + (size_t)id * N, cur, ucur,
                                                                dtaken.p);
            SB_CUDA(cudaGetLastError());
            count_launch();
            unsigned long long taken = 0;
            SB_CUDA(cudaMemcpy(&taken, dtaken.p, sizeof(taken), cudaMemcpyDeviceToHost));
            // E(end+1) = energy(self) (:128): update_energy of the new assignment; an unchanged assignment has the same energy
            Es.push_back(taken ? device_energy(H, W, kernel, ucur, cur, wt, tol, d_min, d_step) : Es.back());
            fusions += 1; rounds += st[0]; solve_ms += st[3]; adopted += (double)taken;
            // :136-146 (iter > 1 always holds after :116)
            if (Es[Es.size() - 2] != Es.back()) {
                std::fill(visited.begin(), visited.end(), 0);
                n_visited = 0;
            } else if (!visited[(size_t)id]) {
                visited[(size_t)id] = 1;
                n_visited++;
            }
            if (n_visited == n_proposals) break;             // :148-150
        }
        if (!on_device) {
            SB_CUDA(cudaMemcpy(assignment, cur, (size_t)N * 4 * 8, cudaMemcpyDeviceToHost));
            SB_CUDA(cudaMemcpy(unary, ucur, (size_t)N * 8, cudaMemcpyDeviceToHost));
        } else {
            SB_CUDA(cudaDeviceSynchronize());
        }
        for (size_t i = 0; i < Es.size(); i++) energies[i] = Es[i];
        *n_energies = (int)Es.size();
        if (stats) { stats[0] = fusions; stats[1] = rounds; stats[2] = solve_ms; stats[3] = adopted; }
    });
}

} // extern "C"
