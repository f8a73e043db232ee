// dispmap_kernels.cu -- sm_100a kernels and C ABI of the array builders that feed the
// fusion solvers (include/stereo_b200.h, "cost volume / unary / pairwise"):
//   K1  ncc_volume        dispmap_ncc.compute_ncc            dispmap_ncc.m:116-198
//       ncc_best_disp     dispmap_ncc.best_disp_from_ncc     dispmap_ncc.m:208-221
//   K2  ncc_sample        dispmap_ncc.sample_ncc_from_disp   dispmap_ncc.m:222-276
//       photo_unary       dispmap_globalstereo.unary_cost    dispmap_globalstereo.m:355-375,405
//       interp2_linear    vgg_interp2 'linear'               imrender/vgg/vgg_interp2.cxx:246-322
//   K3  plane_disparity   disparitymap_from_assignment       dispmap_super.m:318-328,
//                                                             dispmap_globalstereo.m:336-345
//       pairwise_tables   all_pairwise_costs                 dispmap_super.m:226-262
//       fusion_positions  q / qprim of simultaneous_fusion   dispmap_super.m:170-183
//   K6  energy            update_energy                      dispmap_super.m:263-274
//
// All are HBM-streaming kernels: one pass over their inputs, coalesced along the image rows
// (MATLAB column-major: the row index is the fast one), grids sized in multiples of the SM
// count.  Window sums of the NCC volume are exact: 8-bit pixel values make every partial sum
// an integer below 2^24, so fp32 accumulation loses nothing; the cancelling mean / variance
// combination is done in fp64.
#include "sb_common.h"
#include <vector>
#include <cmath>

namespace sb {
namespace dm {

constexpr int NCC_TR = 64;   // tile rows (fast, coalesced dimension)
constexpr int NCC_TC = 16;   // tile columns
constexpr int NCC_PMAX = 8;  // largest half window

__device__ __forceinline__ double matlab_round(double x) { return x < 0 ? -floor(-x + 0.5) : floor(x + 0.5); }

// Box sums of the reference image: sR = sum over window and channels of R, sRR of R^2.
template <typename ACC>
__global__ void ncc_ref_sums_kernel(const ACC *__restrict__ im0, int H, int W, int p, ACC *__restrict__ sR,
                                    ACC *__restrict__ sRR)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)H * W) return;
    const int r = (int)(i % H), c = (int)(i / H);
    ACC a = 0, b = 0;
    for (int ch = 0; ch < 3; ch++)
        for (int dc = -p; dc <= p; dc++) {
            const int cc = c + dc;
            if (cc < 0 || cc >= W) continue;
            for (int dr = -p; dr <= p; dr++) {
                const int rr = r + dr;
                if (rr < 0 || rr >= H) continue;
                const ACC v = im0[(size_t)ch * H * W + (size_t)cc * H + rr];
                a += v;
                b += v * v;
            }
        }
    sR[i] = a;
    sRR[i] = b;
}

// One CTA = NCC_TR x NCC_TC outputs of one disparity level.  ACC = float when every sum is an
// integer below 2^24 (8-bit images, integer disparities: exact), double otherwise.
template <typename ACC>
__global__ void __launch_bounds__(256)
ncc_volume_kernel(const ACC *__restrict__ im0, const ACC *__restrict__ im1, int H, int W, int p,
                  const double *__restrict__ disps, const ACC *__restrict__ sR, const ACC *__restrict__ sRR,
                  float *__restrict__ ncc)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    ACC *sm = reinterpret_cast<ACC *>(sm_raw);
    const int HR = NCC_TR + 2 * p, HC = NCC_TC + 2 * p;
    ACC *pRT = sm, *pT = pRT + HR * HC, *pTT = pT + HR * HC;     // [HC][HR] per-pixel channel sums
    ACC *hRT = pTT + HR * HC, *hT = hRT + HR * NCC_TC, *hTT = hT + HR * NCC_TC; // [TC][HR] after the column box
    const int r0 = blockIdx.x * NCC_TR, c0 = blockIdx.y * NCC_TC, di = blockIdx.z;
    const double d = disps[di];
    const int cfirst = (int)ceil(d);              // 0-based first column of y_span (dispmap_ncc.m:148)
    const int nspan = W - cfirst;
    const double xstep = nspan > 1 ? ((double)W - d - 1.0) / (double)(nspan - 1) : 0.0;
    const size_t plane = (size_t)H * W;
    // ---- per-pixel channel sums over the halo tile
    for (int t = threadIdx.x; t < HR * HC; t += blockDim.x) {
        const int lr = t % HR, lc = t / HR;
        const int r = r0 + lr - p, c = c0 + lc - p;
        ACC vrt = 0, vt = 0, vtt = 0;
        if (r >= 0 && r < H && c >= cfirst && c < W) {
            // interp2 on linspace(1, W-d, nspan) (dispmap_ncc.m:149-153)
            const double X = 1.0 + (double)(c - cfirst) * xstep;
            int x0 = (int)floor(X);
            ACC f = (ACC)(X - (double)x0);
            if (x0 >= W) { x0 = W; f = 0; }
            const int x1 = min(x0 + 1, W);
            for (int ch = 0; ch < 3; ch++) {
                const ACC a = im1[ch * plane + (size_t)(x0 - 1) * H + r];
                ACC tv = a;
                if (f != 0) {
                    const ACC b = im1[ch * plane + (size_t)(x1 - 1) * H + r];
                    tv = a * (1 - f) + b * f;
                }
                const ACC rv = im0[ch * plane + (size_t)c * H + r];
                vrt += rv * tv;
                vt += tv;
                vtt += tv * tv;
            }
        }
        pRT[t] = vrt;
        pT[t] = vt;
        pTT[t] = vtt;
    }
    __syncthreads();
    // ---- box along the columns
    for (int t = threadIdx.x; t < HR * NCC_TC; t += blockDim.x) {
        const int lr = t % HR, lc = t / HR;
        ACC a = 0, b = 0, c = 0;
        for (int k = 0; k <= 2 * p; k++) {
            const int s = (lc + k) * HR + lr;
            a += pRT[s];
            b += pT[s];
            c += pTT[s];
        }
        hRT[t] = a;
        hT[t] = b;
        hTT[t] = c;
    }
    __syncthreads();
    // ---- box along the rows + the NCC combination (dispmap_ncc.m:174-192)
    const double n3 = 3.0 * (2 * p + 1) * (2 * p + 1);
    const int cmask = (int)matlab_round(d + 1.0) - 1;   // first kept column, 0-based (:146)
    for (int t = threadIdx.x; t < NCC_TR * NCC_TC; t += blockDim.x) {
        const int lr = t % NCC_TR, lc = t / NCC_TR;
        const int r = r0 + lr, c = c0 + lc;
        if (r >= H || c >= W) continue;
        ACC a = 0, b = 0, cc = 0;
        for (int k = 0; k <= 2 * p; k++) {
            const int s = lc * HR + lr + k;
            a += hRT[s];
            b += hT[s];
            cc += hTT[s];
        }
        const size_t u = (size_t)c * H + r;
        const double sRT = a, sT = b, sTT = cc, sr = sR[u], srr = sRR[u];
        const double mR = sr / n3, mT = sT / n3;
        const double nr2 = srr - 2.0 * mR * sr + n3 * mR * mR;
        const double nt2 = sTT - 2.0 * mT * sT + n3 * mT * mT;
        const double num = sRT - mR * sT - mT * sr + n3 * mT * mR;
        // complex square roots: real(num / sqrt(nr2) / sqrt(nt2))
        double v;
        if (nr2 >= 0 && nt2 >= 0) v = num / sqrt(nr2) / sqrt(nt2);
        else if (nr2 < 0 && nt2 < 0) v = -(num / sqrt(-nr2) / sqrt(-nt2));
        else v = 0.0;
        if (!isfinite(v)) v = 0.0;
        if (c < cmask) v = 0.0;
        ncc[(size_t)di * plane + u] = (float)v;
    }
}

// best_disp_from_ncc (dispmap_ncc.m:208-221) and sample_ncc_from_disp (:222-245) share the
// parabola through three neighbouring levels (interpolate_ncc, :246-275).
__device__ __forceinline__ void parabola(double y1, double y2, double y3, double d1, double d2, double d3,
                                         double &r, double &pp, double &q)
{
    const double a = y1 / (d1 - d2) / (d1 - d3);
    const double b = y2 / (d2 - d1) / (d2 - d3);
    const double c = y3 / (d3 - d1) / (d3 - d2);
    r = a + b + c;
    pp = -(a * (d2 + d3) + b * (d1 + d3) + c * (d1 + d2));
    q = a * d2 * d3 + b * d1 * d3 + c * d1 * d2;
}

template <typename VT>
__global__ void ncc_best_disp_kernel(const VT *__restrict__ ncc, long long N, int D, const double *__restrict__ disps,
                                     double *__restrict__ best)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= N) return;
    int t2 = 0;
    double y2 = (double)ncc[u];
    for (int i = 1; i < D; i++) {
        const double v = (double)ncc[(size_t)i * N + u];
        if (v > y2) { y2 = v; t2 = i; }     // MATLAB max: first occurrence
    }
    double out = disps[t2];
    if (t2 > 0 && t2 < D - 1) {
        double r, pp, q;
        parabola((double)ncc[(size_t)(t2 - 1) * N + u], y2, (double)ncc[(size_t)(t2 + 1) * N + u], disps[t2 - 1],
                 disps[t2], disps[t2 + 1], r, pp, q);
        out = -pp / r / 2;
    }
    best[u] = out;
}

// out = scale_a * (scale_b - nccs): (1, 0) gives the raw sample, (w, 1)... see callers
template <typename VT>
__global__ void ncc_sample_kernel(const VT *__restrict__ ncc, long long N, int D, const double *__restrict__ disps,
                                  double dmin, double dmax, const double *__restrict__ x, double unary_weight,
                                  int as_unary, double *__restrict__ out)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= N) return;
    const double xv = x[u];
    int t2 = 0;
    double smallest = fabs(xv - disps[0]);
    for (int i = 0; i < D; i++) {
        const double nd = fabs(xv - disps[i]);
        if (nd <= smallest) { smallest = nd; t2 = i; }   // ties -> later level (:232)
    }
    const double y2 = (double)ncc[(size_t)t2 * N + u];
    double v = y2;
    if (t2 > 0 && t2 < D - 1) {
        double r, pp, q;
        parabola((double)ncc[(size_t)(t2 - 1) * N + u], y2, (double)ncc[(size_t)(t2 + 1) * N + u], disps[t2 - 1],
                 disps[t2], disps[t2 + 1], r, pp, q);
        v = r * xv * xv + pp * xv + q;
    }
    if (!(xv <= dmax && xv >= dmin)) v = -1e6;            // :243
    out[u] = as_unary ? unary_weight * (1.0 - v) : v;     // dispmap_ncc.m:113
}

// disparitymap_from_assignment at the node's own point (dispmap_super.m:318-328)
__global__ void plane_disparity_kernel(const double *__restrict__ planes, const double *__restrict__ points,
                                       long long M, double d_min, double d_step, double *__restrict__ out, int *bad)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double a = planes[4 * i], b = planes[4 * i + 1], c = planes[4 * i + 2], d0 = planes[4 * i + 3];
    if (c == 0.0) atomicExch(bad, 1);
    const double d = -((a * points[2 * i] + b * points[2 * i + 1]) + d0) / c;
    out[i] = (d - d_min) / d_step;
}

// vgg_interp2 'linear' (vgg_interp2.cxx:246-322).  A: h x w x col (column-major), X, Y: n.
__device__ __forceinline__ void interp2_point(const double *__restrict__ A, int h, int w, int col, double X, double Y,
                                              double oobv, double *out /* col values */)
{
    const double dw = (double)w, dh = (double)h;
    const size_t step = (size_t)h * w;
    if (X >= 1 && Y >= 1) {
        if (X < dw) {
            if (Y < dh) {
                const int x = (int)X, y = (int)Y;
                const double u = X - x, v = Y - y;
                size_t k = (size_t)h * (x - 1) + y - 1;
                for (int j = 0; j < col; j++, k += step) {
                    // no FMA contraction: same roundings as the reference's x86 build
                    double o = __dadd_rn(A[k], __dmul_rn(A[k + h] - A[k], u));
                    o = __dadd_rn(o, __dmul_rn(__dadd_rn(A[k + 1] - o, __dmul_rn(A[k + h + 1] - A[k + 1], u)), v));
                    out[j] = o;
                }
                return;
            } else if (Y == dh) {
                const int x = (int)X;
                const double u = X - x;
                size_t k = (size_t)h * x - 1;
                for (int j = 0; j < col; j++, k += step) out[j] = __dadd_rn(A[k], __dmul_rn(A[k + h] - A[k], u));
                return;
            }
        } else if (X == dw) {
            if (Y < dh) {
                const int y = (int)Y;
                const double v = Y - y;
                size_t k = (size_t)h * (w - 1) + y - 1;
                for (int j = 0; j < col; j++, k += step) out[j] = __dadd_rn(A[k], __dmul_rn(A[k + 1] - A[k], v));
                return;
            } else if (Y == dh) {
                size_t k = (size_t)h * w - 1;
                for (int j = 0; j < col; j++, k += step) out[j] = A[k];
                return;
            }
        }
    }
    for (int j = 0; j < col; j++) out[j] = oobv;
}

__global__ void interp2_kernel(const double *__restrict__ A, int h, int w, int col, const double *__restrict__ X,
                               const double *__restrict__ Y, long long n, double oobv, double *__restrict__ B)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double o[4];
    for (int j0 = 0; j0 < col; j0 += 4) {
        const int cj = min(4, col - j0);
        interp2_point(A + (size_t)j0 * h * w, h, w, cj, X[i], Y[i], oobv, o);
        for (int j = 0; j < cj; j++) B[(size_t)(j0 + j) * n + i] = o[j];
    }
}

// dispmap_globalstereo.unary_cost (:355-375) + ephoto (:405); P2 = self.P(:,:,2) (4 x 3, column-major)
__global__ void photo_unary_kernel(const double *__restrict__ im0, const double *__restrict__ im1, int H, int W, int C,
                                   const double *__restrict__ P2, const double *__restrict__ ndisp, double d_min,
                                   double d_step, double col_thresh, double *__restrict__ U)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = (long long)H * W;
    if (u >= N) return;
    const int r = (int)(u % H), c = (int)(u / H);
    const double disp = d_step * (ndisp[u] + d_min);       // literal :356
    const double wc[4] = {(double)(c + 1), (double)(r + 1), 1.0, disp};
    double T[3];
    for (int k = 0; k < 3; k++) {
        double s = 0;
        for (int m = 0; m < 4; m++) s += wc[m] * P2[m + 4 * k];
        T[k] = s;
    }
    const double Nn = 1.0 / T[2];
    const double X = T[0] * Nn, Y = T[1] * Nn;
    double acc = 0;
    for (int j0 = 0; j0 < C; j0 += 4) {
        const int cj = min(4, C - j0);
        double m[4];
        interp2_point(im1 + (size_t)j0 * N, H, W, cj, X, Y, -1000.0, m);
        for (int j = 0; j < cj; j++) {
            const double dlt = m[j] - im0[(size_t)(j0 + j) * N + u];
            acc += dlt * dlt;
        }
    }
    U[u] = log(2.0) - log(exp(acc * (-1.0 / (col_thresh * C))) + 1.0);
}

// ---- dispmap_globalstereo.segpln, the window-matching volume (dispmap_globalstereo.m:83-115) ----------------------
// One disparity level at a time: photo cost of every pixel summed over the images (:86-103), separable box mean over
// the (2 w + 1)^2 window ('valid', :104), normalisation (:110-111) and the running first-maximum (:114).
// Pm = n_images x (3 x 4) camera matrices, column-major (the user's P(:, :, a)).
__global__ void wta_cost_kernel(const double *__restrict__ images, int H, int W, int C, int n_images,
                                const double *__restrict__ Pm, double disp, double col_thresh, double *__restrict__ cost)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long N = (long long)H * W;
    if (u >= N) return;
    const int r = (int)(u % H), c = (int)(u / H);
    const double x = (double)(c + 1), y = (double)(r + 1);
    const double *R = images;       // images{1} is the reference (dispmap_globalstereo.m:39-41)
    double total = 0;
    for (int a = 0; a < n_images; a++) {
        const double *P = Pm + 12 * a;
        double X[3];
        for (int k = 0; k < 3; k++) X[k] = (x * P[k] + y * P[k + 3]) + P[k + 6];          // WC * P(:, 1:3, a)' (:89)
        const double d1 = disp * P[9], d2 = disp * P[10], d3 = disp * P[11];               // :95
        const double Z = 1.0 / (X[2] + d3);                                                 // :96
        const double sx = (X[0] + d1) * Z, sy = (X[1] + d2) * Z;
        const double *im = images + (size_t)a * N * C;
        double acc = 0;
        for (int j0 = 0; j0 < C; j0 += 4) {
            const int cj = min(4, C - j0);
            double m[4];
            interp2_point(im + (size_t)j0 * N, H, W, cj, sx, sy, -1000.0, m);               // :99
            for (int j = 0; j < cj; j++) {
                const double dlt = m[j] - R[(size_t)(j0 + j) * N + u];
                acc += dlt * dlt;
            }
        }
        total += log(2.0) - log(exp(acc * (-1.0 / (col_thresh * C))) + 1.0);                // ephoto (:405)
    }
    cost[u] = total;
}
// horizontal pass of conv2(filt, filt', ., 'valid'): out is H x (W - 2 w)
__global__ void wta_hbox_kernel(const double *__restrict__ cost, int H, int W, int w, double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Wi = W - 2 * w;
    if (i >= (long long)H * Wi) return;
    const int r = (int)(i % H), c = (int)(i / H);
    const double f = 1.0 / (double)(2 * w + 1);
    double s = 0;
    for (int dc = 0; dc <= 2 * w; dc++) s += cost[(size_t)(c + dc) * H + r] * f;
    out[i] = s;
}
// vertical pass + normalisation + running maximum; best / best_idx are (H - 2 w) x (W - 2 w)
__global__ void wta_vbox_max_kernel(const double *__restrict__ hb, int H, int W, int w, double X1, int level,
                                    double *__restrict__ best, int *__restrict__ best_idx)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Hi = H - 2 * w, Wi = W - 2 * w;
    if (i >= (long long)Hi * Wi) return;
    const int r = (int)(i % Hi), c = (int)(i / Hi);
    const double f = 1.0 / (double)(2 * w + 1);
    double s = 0;
    for (int dr = 0; dr <= 2 * w; dr++) s += hb[(size_t)c * H + r + dr] * f;
    const double v = (X1 - s) / X1;                         // :111
    if (level == 0 || v > best[i]) {                        // max(., [], 3): the first maximum (:114)
        best[i] = v;
        best_idx[i] = level;
    }
}
// corr = disps(idx); corr(score < min_corr) = 0; padarray(corr, [w w], 'symmetric') (:115-117)
__global__ void wta_finish_kernel(const double *__restrict__ best, const int *__restrict__ best_idx,
                                  const double *__restrict__ disps, int H, int W, int w, double min_corr, double *__restrict__ out)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (long long)H * W) return;
    const int Hi = H - 2 * w, Wi = W - 2 * w;
    int r = (int)(u % H) - w, c = (int)(u / H) - w;
    if (r < 0) r = -r - 1;
    if (r >= Hi) r = 2 * Hi - 1 - r;
    if (c < 0) c = -c - 1;
    if (c >= Wi) c = 2 * Wi - 1 - c;
    const size_t i = (size_t)c * Hi + r;
    out[u] = best[i] < min_corr ? 0.0 : disps[best_idx[i]];
}

// ---- dispmap_ncc.generate_new_plane_RANSAC + fit_plane_to_points (dispmap_ncc.m:48-91) ---------------------------
// One CTA.  The points are the pixels within radius r of (x, y) with the WTA disparity as third coordinate; the plane
// normal is the right singular vector of the smallest singular value of the (IRLS-weighted) centred point matrix, i.e.
// the eigenvector of the smallest eigenvalue of its 3 x 3 scatter matrix (block reduction + Jacobi rotations).
__device__ void jacobi_min_eigenvector(double S[3][3], double v[3])
{
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; sweep++) {
        const double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
        if (off <= 1e-300 || off <= 1e-18 * (fabs(S[0][0]) + fabs(S[1][1]) + fabs(S[2][2]))) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (S[p][q] == 0.0) continue;
                const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; k++) {      // S <- S J
                    const double a = S[k][p], b = S[k][q];
                    S[k][p] = c * a - sn * b;
                    S[k][q] = sn * a + c * b;
                }
                for (int k = 0; k < 3; k++) {      // S <- J' S
                    const double a = S[p][k], b = S[q][k];
                    S[p][k] = c * a - sn * b;
                    S[q][k] = sn * a + c * b;
                }
                for (int k = 0; k < 3; k++) {
                    const double a = V[k][p], b = V[k][q];
                    V[k][p] = c * a - sn * b;
                    V[k][q] = sn * a + c * b;
                }
            }
    }
    int m = 0;
    if (S[1][1] < S[m][m]) m = 1;
    if (S[2][2] < S[m][m]) m = 2;
    for (int k = 0; k < 3; k++) v[k] = V[k][m];
}

constexpr int PF_THREADS = 256;
__device__ __forceinline__ void block_sum(double *vals, int n, double *scratch)
{
    // vals[0..n) per thread -> totals in vals on every thread (n <= 8)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = 0; i < n; i++) {
        double v = vals[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) scratch[warp * 8 + i] = v;
    }
    __syncthreads();
    for (int i = 0; i < n; i++) {
        double t = 0;
        for (int w = 0; w < PF_THREADS / 32; w++) t += scratch[w * 8 + i];
        vals[i] = t;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(PF_THREADS) plane_fit_kernel(const double *__restrict__ disp, int H, int W, double x, double y,
                                                                double r, int kernel, double *__restrict__ out /* 4 + count */)
{
    __shared__ double scratch[(PF_THREADS / 32) * 8];
    __shared__ double sv[3];
    const int c_lo = max(1, (int)floor(x - r)), c_hi = min(W, (int)ceil(x + r));
    const int r_lo = max(1, (int)floor(y - r)), r_hi = min(H, (int)ceil(y + r));
    const int bw = max(c_hi - c_lo + 1, 0), bh = max(r_hi - r_lo + 1, 0);
    const long long box = (long long)bw * bh;
    auto inside = [&](long long i, double &px, double &py, double &pd) {
        const int cc = c_lo + (int)(i / bh), rr = r_lo + (int)(i % bh);
        px = (double)cc; py = (double)rr;
        const double dx = px - x, dy = py - y;
        if (!(sqrt(dx * dx + dy * dy) < r)) return false;        // dispmap_ncc.m:57
        pd = disp[(size_t)(cc - 1) * H + (rr - 1)];
        return true;
    };
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long i = threadIdx.x; i < box; i += PF_THREADS) {
        double px, py, pd;
        if (inside(i, px, py, pd)) { acc[0] += 1.0; acc[1] += px; acc[2] += py; acc[3] += pd; }
    }
    block_sum(acc, 4, scratch);
    const double n = acc[0];
    if (n < 3.0) {
        if (threadIdx.x == 0) { out[0] = out[1] = out[2] = out[3] = 0.0; out[4] = n; }
        return;
    }
    const double mx = acc[1] / n, my = acc[2] / n, md = acc[3] / n;      // c = mean(points, 2) (:68)
    double v[3] = {0, 0, 0};
    const int rounds = kernel == 1 ? 20 : 1;                               // :76-87
    for (int it = 0; it < rounds; it++) {
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (long long i = threadIdx.x; i < box; i += PF_THREADS) {
            double px, py, pd;
            if (!inside(i, px, py, pd)) continue;
            const double a0 = -(px - mx), a1 = -(py - my), a2 = -(pd - md);   // cost_func = -(points - c)' (:71)
            // w .* cost_func with w = sqrt(|cost_func * V(:, end)|) of the previous round (ones in the first): the scatter
            // matrix carries w^2
            const double w2 = it == 0 ? 1.0 : fabs(a0 * v[0] + a1 * v[1] + a2 * v[2]);
            s[0] += w2 * a0 * a0; s[1] += w2 * a0 * a1; s[2] += w2 * a0 * a2;
            s[3] += w2 * a1 * a1; s[4] += w2 * a1 * a2; s[5] += w2 * a2 * a2;
        }
        block_sum(s, 6, scratch);
        if (threadIdx.x == 0) {
            double S[3][3] = {{s[0], s[1], s[2]}, {s[1], s[3], s[4]}, {s[2], s[4], s[5]}};
            double e[3];
            jacobi_min_eigenvector(S, e);
            sv[0] = e[0]; sv[1] = e[1]; sv[2] = e[2];
        }
        __syncthreads();
        v[0] = sv[0]; v[1] = sv[1]; v[2] = sv[2];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double p4 = -(v[0] * mx + v[1] * my + v[2] * md);              // :89
        out[0] = v[0] / v[2]; out[1] = v[1] / v[2]; out[2] = 1.0; out[3] = p4 / v[2];   // p = p / p(3) (:90)
        out[4] = n;
    }
}
// proposal = repmat(p, [1 N]) (:65)
__global__ void plane_repeat_kernel(const double *__restrict__ p4, long long N, double *__restrict__ out)
{
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= N) return;
    reinterpret_cast<double2 *>(out)[2 * u] = make_double2(p4[0], p4[1]);
    reinterpret_cast<double2 *>(out)[2 * u + 1] = make_double2(p4[2], p4[3]);
}

// dispmap_globalstereo.preprocess, the smoothness weights (dispmap_globalstereo.m:398-401): lambda_h on the terms whose
// two pixels share a segment, lambda_l on those that cross a segment boundary, both scaled by num_in / (connect == 8 + 1).
__global__ void smooth_weights_kernel(const unsigned *__restrict__ segment, int H, int W, long long E, double w_same,
                                      double w_cross, double *__restrict__ out)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    long long i1, i2;
    {
        const long long nV = (long long)(H - 1) * W, nH = (long long)H * (W - 1);
        if (p < 2 * nV) {
            const long long k = p < nV ? p : p - nV;
            const long long c = k / (H - 1), r = k % (H - 1);
            i1 = r + (long long)H * c; i2 = i1 + 1;
        } else {
            const long long k0 = p - 2 * nV;
            i1 = k0 < nH ? k0 : k0 - nH; i2 = i1 + H;
        }
    }
    out[p] = segment[i1] == segment[i2] ? w_same : w_cross;
}

// Geometry of pairwise term p of the dispmap_super grid: tail / head node (0-based) and the
// point of the head (dispmap_super.m:279-302 order).
__device__ __forceinline__ void term_nodes(long long p, int H, int W, long long &i1, long long &i2)
{
    const long long nV = (long long)(H - 1) * W, nH = (long long)H * (W - 1);
    if (p < 2 * nV) {
        const long long k = p < nV ? p : p - nV;
        const long long c = k / (H - 1), r = k % (H - 1);
        const long long s = r + (long long)H * c, f = s + 1;
        if (p < nV) { i1 = s; i2 = f; } else { i1 = f; i2 = s; }
    } else {
        const long long k0 = p - 2 * nV;
        const long long k = k0 < nH ? k0 : k0 - nH;
        const long long s = k, f = k + H;
        if (k0 < nH) { i1 = s; i2 = f; } else { i1 = f; i2 = s; }
    }
}

__device__ __forceinline__ double plane_at(const double *__restrict__ planes, long long node, double x, double y,
                                           double d_min, double d_step)
{
    const double *a = planes + 4 * node;
    return (-((a[0] * x + a[1] * y) + a[3]) / a[2] - d_min) / d_step;
}

__device__ __forceinline__ double pair_cost(int kernel, double w, double tol, double p, double q)
{
    const double d = p - q;
    return w * fmin(kernel == 1 ? fabs(d) : d * d, tol);
}

// all_pairwise_costs (dispmap_super.m:236-262).  prop may be null (E00 only).
__global__ void pairwise_tables_kernel(int H, int W, int kernel, const double *__restrict__ cur,
                                       const double *__restrict__ prop, const double *__restrict__ weights, double tol,
                                       double d_min, double d_step, long long E, double *__restrict__ E00,
                                       double *__restrict__ E01, double *__restrict__ E10, double *__restrict__ E11)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    long long i1, i2;
    term_nodes(p, H, W, i1, i2);
    const double x = (double)(i2 / H + 1), y = (double)(i2 % H + 1);
    const double w = weights[p];
    const double q = plane_at(cur, i2, x, y, d_min, d_step), qprim = plane_at(cur, i1, x, y, d_min, d_step);
    E00[p] = pair_cost(kernel, w, tol, q, qprim);
    if (prop) {
        const double nq = plane_at(prop, i2, x, y, d_min, d_step), nqprim = plane_at(prop, i1, x, y, d_min, d_step);
        E11[p] = pair_cost(kernel, w, tol, nq, nqprim);
        E10[p] = pair_cost(kernel, w, tol, q, nqprim);
        E01[p] = pair_cost(kernel, w, tol, nq, qprim);
    }
}

// q(l, p), qprim(l, p) of simultaneous_fusion (dispmap_super.m:170-183); proposals [L][4][N]
__global__ void fusion_positions_kernel(int H, int W, int L, const double *__restrict__ props, double d_min,
                                        double d_step, long long E, double *__restrict__ q, double *__restrict__ qprim)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * L) return;
    const long long p = t / L;
    const int l = (int)(t % L);
    long long i1, i2;
    term_nodes(p, H, W, i1, i2);
    const double x = (double)(i2 / H + 1), y = (double)(i2 % H + 1);
    const double *pl = props + (size_t)l * 4 * H * W;
    q[t] = plane_at(pl, i2, x, y, d_min, d_step);
    qprim[t] = plane_at(pl, i1, x, y, d_min, d_step);
}

// sum over a double array (deterministic two-stage tree)
__global__ void reduce_sum_kernel(const double *__restrict__ a, long long n, double *__restrict__ partial)
{
    __shared__ double sh[256];
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s += a[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

inline unsigned blocks_for(long long n, int t = 256) { return (unsigned)std::max<long long>(1, (n + t - 1) / t); }

template <typename T> void upload(DevBuf<T> &b, const T *h, size_t n)
{
    b.alloc(n);
    if (n) SB_CUDA(cudaMemcpy(b.p, h, n * sizeof(T), cudaMemcpyHostToDevice));
}

__global__ void to_float_kernel(const double *__restrict__ a, float *__restrict__ b, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i];
}
__global__ void to_double_kernel(const float *__restrict__ a, double *__restrict__ b, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}

double device_sum(const double *d, long long n)
{
    const int nb = 148 * 4;
    DevBuf<double> part(nb);
    reduce_sum_kernel<<<nb, 256>>>(d, n, part.p);
    SB_CUDA(cudaGetLastError());
    count_launch();
    std::vector<double> h(nb);
    SB_CUDA(cudaMemcpy(h.data(), part.p, nb * sizeof(double), cudaMemcpyDeviceToHost));
    double s = 0;
    for (double v : h) s += v;
    return s;
}

double compute_ncc_volume(int H, int W, const double *d_im0, const double *d_im1, const double *h_im0, const double *h_im1, int D,
                          const double *h_disps, const double *d_disps, int patchsize, float *vol, bool *fast);

// The general volume kernel (any disparities, any pixel values): one CTA per 64 x 16 tile of one level.
int ncc_volume_general(int H, int W, const double *d_im0, const double *d_im1, int D, const double *h_disps, const double *d_disps,
                       int patchsize, bool exact32, float *vol)
{
    (void)h_disps;
    const long long N = (long long)H * W;
    const int HR = NCC_TR + 2 * patchsize, HC = NCC_TC + 2 * patchsize;
    dim3 grid((H + NCC_TR - 1) / NCC_TR, (W + NCC_TC - 1) / NCC_TC, D);
    if (exact32) {
        // 8-bit integer images, integer disparities, every window sum < 2^24: exact in fp32
        DevBuf<float> f0((size_t)N * 3), f1((size_t)N * 3), sR((size_t)N), sRR((size_t)N);
        to_float_kernel<<<blocks_for(N * 3), 256>>>(d_im0, f0.p, N * 3);
        to_float_kernel<<<blocks_for(N * 3), 256>>>(d_im1, f1.p, N * 3);
        ncc_ref_sums_kernel<float><<<blocks_for(N), 256>>>(f0.p, H, W, patchsize, sR.p, sRR.p);
        const size_t smem = (size_t)(3 * HR * HC + 3 * HR * NCC_TC) * sizeof(float);
        SB_CUDA(cudaFuncSetAttribute(ncc_volume_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ncc_volume_kernel<float><<<grid, 256, smem>>>(f0.p, f1.p, H, W, patchsize, d_disps, sR.p, sRR.p, vol);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaDeviceSynchronize());
        count_launch(4);
    } else {
        DevBuf<double> sR((size_t)N), sRR((size_t)N);
        ncc_ref_sums_kernel<double><<<blocks_for(N), 256>>>(d_im0, H, W, patchsize, sR.p, sRR.p);
        const size_t smem = (size_t)(3 * HR * HC + 3 * HR * NCC_TC) * sizeof(double);
        SB_CUDA(cudaFuncSetAttribute(ncc_volume_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ncc_volume_kernel<double><<<grid, 256, smem>>>(d_im0, d_im1, H, W, patchsize, d_disps, sR.p, sRR.p, vol);
        SB_CUDA(cudaGetLastError());
        SB_CUDA(cudaDeviceSynchronize());
        count_launch(2);
    }
    return 0;
}

} // namespace dm
} // namespace sb

using namespace sb;
using namespace sb::dm;

extern "C" {

// ---------------------------------------------------------------- NCC volume: host-array entry and device-resident handle
struct sb_ncc_vol {
    int H, W, D;
    sb::DevBuf<float> vol;         // [D][W][H]
    sb::DevBuf<double> disps;
    std::vector<double> hdisps;
    double kernel_ms;
    int fast;
};

static void ncc_volume_build(int H, int W, int C, const double *im0, const double *im1, int D, const double *disparities,
                             int patchsize, sb_ncc_vol &v)
{
    SB_REQUIRE(H >= 1 && W >= 1 && D >= 1 && im0 && im1 && disparities, SB_EINVAL, "sb_ncc_volume: bad arguments");
    SB_REQUIRE(C == 3, SB_EINVAL, "sb_ncc_volume: images must have 3 channels (dispmap_ncc.m:125-131)");
    SB_REQUIRE(patchsize >= 0 && patchsize <= NCC_PMAX, SB_EUNSUP, "sb_ncc_volume: patchsize %d outside [0, %d]", patchsize, NCC_PMAX);
    require_device();
    const long long N = (long long)H * W;
    DevBuf<double> raw0, raw1;
    upload(raw0, im0, (size_t)N * 3);
    upload(raw1, im1, (size_t)N * 3);
    upload(v.disps, disparities, (size_t)D);
    v.hdisps.assign(disparities, disparities + D);
    v.H = H; v.W = W; v.D = D;
    v.vol.alloc((size_t)N * D);
    bool fast = false;
    v.kernel_ms = compute_ncc_volume(H, W, raw0.p, raw1.p, im0, im1, D, disparities, v.disps.p, patchsize, v.vol.p, &fast);
    v.fast = fast ? 1 : 0;
}

static void ncc_volume_to_host(const sb_ncc_vol &v, double *ncc_out)
{
    // back to the MATLAB layout in doubles, one level at a time (bounded staging)
    const long long N = (long long)v.H * v.W;
    DevBuf<double> stage((size_t)N);
    for (int i = 0; i < v.D; i++) {
        to_double_kernel<<<blocks_for(N), 256>>>(v.vol.p + (size_t)i * N, stage.p, N);
        SB_CUDA(cudaMemcpy(ncc_out + (size_t)i * N, stage.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    }
    count_launch(v.D);
}

int sb_ncc_volume(int H, int W, int C, const double *im0, const double *im1, int D, const double *disparities,
                  int patchsize, double *ncc_out)
{
    return guarded([&] {
        SB_REQUIRE(ncc_out, SB_EINVAL, "sb_ncc_volume: bad arguments");
        sb_ncc_vol v;
        ncc_volume_build(H, W, C, im0, im1, D, disparities, patchsize, v);
        ncc_volume_to_host(v, ncc_out);
    });
}

int sb_ncc_vol_create(int H, int W, int C, const double *im0, const double *im1, int D, const double *disparities,
                      int patchsize, sb_ncc_vol **out)
{
    return guarded([&] {
        SB_REQUIRE(out, SB_EINVAL, "sb_ncc_vol_create: null output");
        *out = nullptr;
        sb_ncc_vol *v = new sb_ncc_vol();
        try {
            ncc_volume_build(H, W, C, im0, im1, D, disparities, patchsize, *v);
        } catch (...) {
            delete v;
            throw;
        }
        *out = v;
    });
}

int sb_ncc_vol_get(sb_ncc_vol *v, double *ncc_out)
{
    return guarded([&] {
        SB_REQUIRE(v && ncc_out, SB_EINVAL, "sb_ncc_vol_get: null pointer");
        ncc_volume_to_host(*v, ncc_out);
    });
}

int sb_ncc_vol_best_disp(sb_ncc_vol *v, double *best_disp)
{
    return guarded([&] {
        SB_REQUIRE(v && best_disp, SB_EINVAL, "sb_ncc_vol_best_disp: null pointer");
        const long long N = (long long)v->H * v->W;
        DevBuf<double> out((size_t)N);
        ncc_best_disp_kernel<float><<<blocks_for(N), 256>>>(v->vol.p, N, v->D, v->disps.p, out.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(best_disp, out.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_ncc_vol_sample(sb_ncc_vol *v, const double *disps, double unary_weight, int as_unary, double *out)
{
    return guarded([&] {
        SB_REQUIRE(v && disps && out, SB_EINVAL, "sb_ncc_vol_sample: null pointer");
        const long long N = (long long)v->H * v->W;
        double dmin = v->hdisps[0], dmax = v->hdisps[0];
        for (int i = 1; i < v->D; i++) { dmin = std::min(dmin, v->hdisps[i]); dmax = std::max(dmax, v->hdisps[i]); }
        DevBuf<double> x, o((size_t)N);
        upload(x, disps, (size_t)N);
        ncc_sample_kernel<float><<<blocks_for(N), 256>>>(v->vol.p, N, v->D, v->disps.p, dmin, dmax, x.p, unary_weight, as_unary, o.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(out, o.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_ncc_vol_info(sb_ncc_vol *v, double *info)
{
    return guarded([&] {
        SB_REQUIRE(v && info, SB_EINVAL, "sb_ncc_vol_info: null pointer");
        info[0] = v->kernel_ms;
        info[1] = (double)v->fast;
        info[2] = (double)v->vol.bytes();
    });
}

void sb_ncc_vol_destroy(sb_ncc_vol *v) { delete v; }

int sb_ncc_best_disp(int H, int W, int D, const double *ncc, const double *disparities, double *best_disp)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && D >= 1 && ncc && disparities && best_disp, SB_EINVAL, "sb_ncc_best_disp: bad arguments");
        require_device();
        const long long N = (long long)H * W;
        DevBuf<double> v, dd, out((size_t)N);
        upload(v, ncc, (size_t)N * D);
        upload(dd, disparities, (size_t)D);
        ncc_best_disp_kernel<double><<<blocks_for(N), 256>>>(v.p, N, D, dd.p, out.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(best_disp, out.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_ncc_sample(int H, int W, int D, const double *ncc, const double *disparities, const double *disps,
                  double unary_weight, int as_unary, double *out)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && D >= 1 && ncc && disparities && disps && out, SB_EINVAL, "sb_ncc_sample: bad arguments");
        require_device();
        const long long N = (long long)H * W;
        double dmin = disparities[0], dmax = disparities[0];
        for (int i = 1; i < D; i++) { dmin = std::min(dmin, disparities[i]); dmax = std::max(dmax, disparities[i]); }
        DevBuf<double> v, dd, x, o((size_t)N);
        upload(v, ncc, (size_t)N * D);
        upload(dd, disparities, (size_t)D);
        upload(x, disps, (size_t)N);
        ncc_sample_kernel<double><<<blocks_for(N), 256>>>(v.p, N, D, dd.p, dmin, dmax, x.p, unary_weight, as_unary, o.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(out, o.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_plane_disparity(int64_t M, const double *planes, const double *points, double d_min, double d_step, double *out)
{
    return guarded([&] {
        SB_REQUIRE(M >= 0 && (M == 0 || (planes && points && out)), SB_EINVAL, "sb_plane_disparity: bad arguments");
        if (M == 0) return;
        require_device();
        DevBuf<double> pl, pt, o((size_t)M);
        DevBuf<int> bad(1);
        upload(pl, planes, (size_t)M * 4);
        upload(pt, points, (size_t)M * 2);
        SB_CUDA(cudaMemset(bad.p, 0, sizeof(int)));
        plane_disparity_kernel<<<blocks_for(M), 256>>>(pl.p, pt.p, M, d_min, d_step, o.p, bad.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        int hb = 0;
        SB_CUDA(cudaMemcpy(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
        SB_REQUIRE(!hb, SB_EINVAL, "Infinite disparity");   // dispmap_super.m:323-325
        SB_CUDA(cudaMemcpy(out, o.p, (size_t)M * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_interp2_linear(const double *A, int h, int w, int col, const double *X, const double *Y, int64_t n, double oobv,
                      double *B)
{
    return guarded([&] {
        SB_REQUIRE(A && h >= 1 && w >= 1 && col >= 1 && n >= 0 && (n == 0 || (X && Y && B)), SB_EINVAL, "sb_interp2_linear: bad arguments");
        if (n == 0) return;
        require_device();
        DevBuf<double> dA, dX, dY, dB((size_t)n * col);
        upload(dA, A, (size_t)h * w * col);
        upload(dX, X, (size_t)n);
        upload(dY, Y, (size_t)n);
        interp2_kernel<<<blocks_for(n), 256>>>(dA.p, h, w, col, dX.p, dY.p, n, oobv, dB.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(B, dB.p, (size_t)n * col * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_segpln_wta(int H, int W, int C, int n_images, const double *images, const double *P, int D, const double *disps,
                  int window, double col_thresh, double min_corr, double *corr, double *score)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && C >= 1 && n_images >= 1 && D >= 1 && images && P && disps && corr, SB_EINVAL,
                   "sb_segpln_wta: bad arguments");
        SB_REQUIRE(window >= 0 && 3 * window < H && 3 * window < W, SB_EINVAL,
                   "sb_segpln_wta: window %d does not fit a %d x %d image (symmetric padding needs H, W > 3 window)", window, H, W);
        require_device();
        const long long N = (long long)H * W;
        const int Hi = H - 2 * window, Wi = W - 2 * window;
        DevBuf<double> im, dP, dd, cost((size_t)N), hb((size_t)H * Wi), best((size_t)Hi * Wi), out((size_t)N);
        DevBuf<int> bidx((size_t)Hi * Wi);
        upload(im, images, (size_t)n_images * N * C);
        upload(dP, P, (size_t)12 * n_images);
        upload(dd, disps, (size_t)D);
        // X = ephoto(-1000 - Rvec) * numel(images); only X(1), the value at pixel (1, 1), is used (:110-111)
        double acc = 0;
        for (int j = 0; j < C; j++) { const double v = -1000.0 - images[(size_t)j * N]; acc += v * v; }
        const double X1 = (std::log(2.0) - std::log(std::exp(acc * (-1.0 / (col_thresh * C))) + 1.0)) * n_images;
        for (int b = 0; b < D; b++) {
            wta_cost_kernel<<<blocks_for(N), 256>>>(im.p, H, W, C, n_images, dP.p, disps[b], col_thresh, cost.p);
            wta_hbox_kernel<<<blocks_for((long long)H * Wi), 256>>>(cost.p, H, W, window, hb.p);
            wta_vbox_max_kernel<<<blocks_for((long long)Hi * Wi), 256>>>(hb.p, H, W, window, X1, b, best.p, bidx.p);
        }
        wta_finish_kernel<<<blocks_for(N), 256>>>(best.p, bidx.p, dd.p, H, W, window, min_corr, out.p);
        SB_CUDA(cudaGetLastError());
        count_launch(3 * D + 1);
        SB_CUDA(cudaMemcpy(corr, out.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
        if (score) SB_CUDA(cudaMemcpy(score, best.p, (size_t)Hi * Wi * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_plane_from_disparity(int H, int W, const double *disp, double x, double y, double r, int kernel, int on_device,
                            double *plane, double *proposal, double *n_points)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && disp && plane, SB_EINVAL, "sb_plane_from_disparity: bad arguments");
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unkown kernel type");
        SB_REQUIRE(r > 0, SB_EINVAL, "sb_plane_from_disparity: radius must be positive");
        require_device();
        const long long N = (long long)H * W;
        DevBuf<double> dd, out(5), dprop;
        const double *d = disp;
        if (!on_device) { upload(dd, disp, (size_t)N); d = dd.p; }
        plane_fit_kernel<<<1, PF_THREADS>>>(d, H, W, x, y, r, kernel, out.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        double h[5];
        SB_CUDA(cudaMemcpy(h, out.p, sizeof(h), cudaMemcpyDeviceToHost));
        if (n_points) *n_points = h[4];
        SB_REQUIRE(h[4] >= 3.0, SB_EINVAL, "sb_plane_from_disparity: %d points within the radius (at least 3 needed)", (int)h[4]);
        for (int i = 0; i < 4; i++) plane[i] = h[i];
        if (proposal) {
            double *dst = proposal;
            if (!on_device) { dprop.alloc((size_t)N * 4); dst = dprop.p; }
            plane_repeat_kernel<<<blocks_for(N), 256>>>(out.p, N, dst);
            SB_CUDA(cudaGetLastError());
            count_launch();
            if (!on_device) SB_CUDA(cudaMemcpy(proposal, dst, (size_t)N * 4 * 8, cudaMemcpyDeviceToHost));
            else SB_CUDA(cudaDeviceSynchronize());
        }
    });
}

int sb_smooth_weights(int H, int W, const uint32_t *segment, double lambda_h, double lambda_l, double scale, double *weights)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && segment && weights, SB_EINVAL, "sb_smooth_weights: bad arguments");
        require_device();
        const long long N = (long long)H * W, E = 2 * ((long long)(H - 1) * W + (long long)H * (W - 1));
        if (E == 0) return;
        DevBuf<unsigned> seg((size_t)N);
        DevBuf<double> o((size_t)E);
        SB_CUDA(cudaMemcpyAsync(seg.p, segment, (size_t)N * 4, cudaMemcpyHostToDevice, 0));
        // EW = EW * lambda_h + ~EW * lambda_l; EW = EW * (num_in / ((connect == 8) + 1))  (:399-400)
        smooth_weights_kernel<<<blocks_for(E), 256>>>(seg.p, H, W, E, lambda_h * scale, lambda_l * scale, o.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(weights, o.p, (size_t)E * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_photo_unary(int H, int W, int C, const double *im0, const double *im1, const double *P2, const double *planes,
                   double d_min, double d_step, double col_thresh, double *U)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && C >= 1 && im0 && im1 && P2 && planes && U, SB_EINVAL, "sb_photo_unary: bad arguments");
        require_device();
        const long long N = (long long)H * W;
        std::vector<double> pts((size_t)N * 2);
        for (long long u = 0; u < N; u++) { pts[2 * u] = (double)(u / H + 1); pts[2 * u + 1] = (double)(u % H + 1); }
        DevBuf<double> a0, a1, dP, pl, pt, nd((size_t)N), o((size_t)N);
        DevBuf<int> bad(1);
        upload(a0, im0, (size_t)N * C);
        upload(a1, im1, (size_t)N * C);
        upload(dP, P2, 12);
        upload(pl, planes, (size_t)N * 4);
        upload(pt, pts.data(), (size_t)N * 2);
        SB_CUDA(cudaMemset(bad.p, 0, sizeof(int)));
        plane_disparity_kernel<<<blocks_for(N), 256>>>(pl.p, pt.p, N, d_min, d_step, nd.p, bad.p);
        photo_unary_kernel<<<blocks_for(N), 256>>>(a0.p, a1.p, H, W, C, dP.p, nd.p, d_min, d_step, col_thresh, o.p);
        SB_CUDA(cudaGetLastError());
        count_launch(2);
        int hb = 0;
        SB_CUDA(cudaMemcpy(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost));
        SB_REQUIRE(!hb, SB_EINVAL, "Infinite disparity");
        SB_CUDA(cudaMemcpy(U, o.p, (size_t)N * 8, cudaMemcpyDeviceToHost));
    });
}

} // extern "C"

namespace sb {
// device pointers in, device pointers out (sb_binary_fusion_grid, qpbo.cu)
void launch_pairwise_tables(int H, int W, int kernel, const double *cur, const double *prop, const double *weights, double tol,
                            double d_min, double d_step, long long E, double *E00, double *E01, double *E10, double *E11)
{
    pairwise_tables_kernel<<<blocks_for(E), 256>>>(H, W, kernel, cur, prop, weights, tol, d_min, d_step, E, E00, E01, E10, E11);
    SB_CUDA(cudaGetLastError());
    count_launch();
}
// dispmap_super.update_energy (dispmap_super.m:263-274) on device-resident fields: sum of the unary costs plus the sum of the
// (current, current) pairwise table.  Fixed partial sums added on the host: the same fields give the same bits.
double device_energy(int H, int W, int kernel, const double *d_unary, const double *d_assignment, const double *d_weights,
                     double tol, double d_min, double d_step)
{
    const long long N = (long long)H * W, E = 2 * ((long long)(H - 1) * W + (long long)H * (W - 1));
    double e = dm::device_sum(d_unary, N);
    if (E > 0) {
        DevBuf<double> o((size_t)E);
        pairwise_tables_kernel<<<blocks_for(E), 256>>>(H, W, kernel, d_assignment, nullptr, d_weights, tol, d_min, d_step, E, o.p,
                                                       nullptr, nullptr, nullptr);
        SB_CUDA(cudaGetLastError());
        count_launch();
        e += dm::device_sum(o.p, E);
    }
    return e;
}
} // namespace sb

extern "C" {

int sb_pairwise_tables(int H, int W, int kernel, const double *assignment, const double *proposal,
                       const double *weights, double tol, double d_min, double d_step, double *E00, double *E01,
                       double *E10, double *E11)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && assignment && weights && E00, SB_EINVAL, "sb_pairwise_tables: bad arguments");
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unkown kernel type");   // dispmap_super.m:232-233
        SB_REQUIRE(!proposal || (E01 && E10 && E11), SB_EINVAL, "sb_pairwise_tables: null output");
        require_device();
        const long long N = (long long)H * W, E = 2 * ((long long)(H - 1) * W + (long long)H * (W - 1));
        if (E == 0) return;
        DevBuf<double> cur, prop, wt, o((size_t)E * (proposal ? 4 : 1));
        upload(cur, assignment, (size_t)N * 4);
        if (proposal) upload(prop, proposal, (size_t)N * 4);
        upload(wt, weights, (size_t)E);
        pairwise_tables_kernel<<<blocks_for(E), 256>>>(H, W, kernel, cur.p, proposal ? prop.p : nullptr, wt.p, tol, d_min,
                                                       d_step, E, o.p, o.p + (proposal ? E : 0), o.p + (proposal ? 2 * E : 0),
                                                       o.p + (proposal ? 3 * E : 0));
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(E00, o.p, (size_t)E * 8, cudaMemcpyDeviceToHost));
        if (proposal) {
            SB_CUDA(cudaMemcpy(E01, o.p + E, (size_t)E * 8, cudaMemcpyDeviceToHost));
            SB_CUDA(cudaMemcpy(E10, o.p + 2 * E, (size_t)E * 8, cudaMemcpyDeviceToHost));
            SB_CUDA(cudaMemcpy(E11, o.p + 3 * E, (size_t)E * 8, cudaMemcpyDeviceToHost));
        }
    });
}

int sb_fusion_positions(int H, int W, int L, const double *proposals, double d_min, double d_step, double *q,
                        double *qprim)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && L >= 1 && proposals && q && qprim, SB_EINVAL, "sb_fusion_positions: bad arguments");
        require_device();
        const long long N = (long long)H * W, E = 2 * ((long long)(H - 1) * W + (long long)H * (W - 1));
        if (E == 0) return;
        DevBuf<double> pr, dq((size_t)E * L), dqp((size_t)E * L);
        upload(pr, proposals, (size_t)N * 4 * L);
        fusion_positions_kernel<<<blocks_for(E * L), 256>>>(H, W, L, pr.p, d_min, d_step, E, dq.p, dqp.p);
        SB_CUDA(cudaGetLastError());
        count_launch();
        SB_CUDA(cudaMemcpy(q, dq.p, (size_t)E * L * 8, cudaMemcpyDeviceToHost));
        SB_CUDA(cudaMemcpy(qprim, dqp.p, (size_t)E * L * 8, cudaMemcpyDeviceToHost));
    });
}

int sb_energy(int H, int W, int kernel, const double *unary, const double *assignment, const double *weights, double tol,
              double d_min, double d_step, double *energy)
{
    return guarded([&] {
        SB_REQUIRE(H >= 1 && W >= 1 && unary && assignment && weights && energy, SB_EINVAL, "sb_energy: bad arguments");
        SB_REQUIRE(kernel == 1 || kernel == 2, SB_EINVAL, "Unkown kernel type");
        require_device();
        const long long N = (long long)H * W, E = 2 * ((long long)(H - 1) * W + (long long)H * (W - 1));
        DevBuf<double> un, cur, wt;
        upload(un, unary, (size_t)N);
        upload(cur, assignment, (size_t)N * 4);
        if (E > 0) upload(wt, weights, (size_t)E);
        *energy = device_energy(H, W, kernel, un.p, cur.p, wt.p, tol, d_min, d_step);
    });
}

} // extern "C"
