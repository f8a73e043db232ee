// sb_builders_mex.cpp -- one gateway for the array builders the dispmap_* classes call
// (cost volume, unary sampling, plane -> disparity, pairwise tables, q / qprim, energy).
//   out = sb_builders_mex('ncc_volume', im0, im1, disparities, patchsize)
//   out = sb_builders_mex('ncc_best_disp', ncc, disparities)
//   out = sb_builders_mex('ncc_sample', ncc, disparities, disps, unary_weight, as_unary)
//   out = sb_builders_mex('plane_disparity', planes, points, d_min, d_step)
//   out = sb_builders_mex('interp2_linear', A, X, Y, oobv)
//   out = sb_builders_mex('photo_unary', im0, im1, P2, planes, d_min, d_step, col_thresh)
//   [E00,E01,E10,E11] = sb_builders_mex('pairwise_tables', sz, kernel, assignment, proposal, weights, tol, d_min, d_step)
//   [q,qprim] = sb_builders_mex('fusion_positions', sz, proposals(4 x N x L), d_min, d_step)
//   e = sb_builders_mex('energy', sz, kernel, unary, assignment, weights, tol, d_min, d_step)
//   [corr, score] = sb_builders_mex('segpln_wta', images(H x W x C x n), P(3 x 4 x n), disps, window, col_thresh, min_corr)
//   w = sb_builders_mex('smooth_weights', segment(H x W uint32), lambda_h, lambda_l, scale)
//   [p, proposal] = sb_builders_mex('plane_from_disparity', best_disp(H x W), x, y, r, kernel)
//   [assignment, E, unary] = sb_builders_mex('fuse_until_convergence', sz, kernel, proposals(4 x N x n), unaries(N x n),
//                                assignment, unary, weights, tol, d_min, d_step, improve, maxiter, ids(int32))
// Each forwards to the entry point of the same name in include/stereo_b200.h.
#include "sb_mex_common.h"

static double scalar(const mxArray *a) { return *mxGetPr(a); }

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    SB_MEX_ASSERT(nrhs >= 1 && mxGetClassID(prhs[0]) == mxCHAR_CLASS);
    char op[64];
    SB_MEX_ASSERT(mxGetString(prhs[0], op, sizeof(op)) == 0);
    const mxArray **a = prhs + 1;
    const int n = nrhs - 1;
    if (!strcmp(op, "ncc_volume")) {
        SB_MEX_ASSERT(n == 4 && mxGetNumberOfDimensions(a[0]) == 3);
        const mwSize *d = mxGetDimensions(a[0]);
        const int D = (int)mxGetNumberOfElements(a[2]);
        mwSize od[3] = {d[0], d[1], (mwSize)D};
        plhs[0] = mxCreateNumericArray(3, od, mxDOUBLE_CLASS, mxREAL);
        sb_mex_check(sb_ncc_volume((int)d[0], (int)d[1], (int)d[2], mxGetPr(a[0]), mxGetPr(a[1]), D, mxGetPr(a[2]),
                                   (int)scalar(a[3]), mxGetPr(plhs[0])));
    } else if (!strcmp(op, "ncc_best_disp") || !strcmp(op, "ncc_sample")) {
        const mwSize *d = mxGetDimensions(a[0]);
        const int D = mxGetNumberOfDimensions(a[0]) == 3 ? (int)d[2] : 1;
        plhs[0] = sb_mex_matrix(d[0], d[1]);
        if (op[4] == 'b') {
            SB_MEX_ASSERT(n == 2);
            sb_mex_check(sb_ncc_best_disp((int)d[0], (int)d[1], D, mxGetPr(a[0]), mxGetPr(a[1]), mxGetPr(plhs[0])));
        } else {
            SB_MEX_ASSERT(n == 5);
            sb_mex_check(sb_ncc_sample((int)d[0], (int)d[1], D, mxGetPr(a[0]), mxGetPr(a[1]), mxGetPr(a[2]), scalar(a[3]),
                                       (int)scalar(a[4]), mxGetPr(plhs[0])));
        }
    } else if (!strcmp(op, "plane_disparity")) {
        SB_MEX_ASSERT(n == 4 && mxGetM(a[0]) == 4 && mxGetM(a[1]) == 2 && mxGetN(a[0]) == mxGetN(a[1]));
        plhs[0] = sb_mex_matrix(1, mxGetN(a[0]));
        sb_mex_check(sb_plane_disparity((int64_t)mxGetN(a[0]), mxGetPr(a[0]), mxGetPr(a[1]), scalar(a[2]), scalar(a[3]),
                                        mxGetPr(plhs[0])));
    } else if (!strcmp(op, "interp2_linear")) {
        SB_MEX_ASSERT(n == 4);
        const mwSize *d = mxGetDimensions(a[0]);
        const int col = mxGetNumberOfDimensions(a[0]) == 3 ? (int)d[2] : 1;
        const mwSize np = (mwSize)mxGetNumberOfElements(a[1]);
        plhs[0] = sb_mex_matrix(np, col);
        sb_mex_check(sb_interp2_linear(mxGetPr(a[0]), (int)d[0], (int)d[1], col, mxGetPr(a[1]), mxGetPr(a[2]), (int64_t)np,
                                       scalar(a[3]), mxGetPr(plhs[0])));
    } else if (!strcmp(op, "photo_unary")) {
        SB_MEX_ASSERT(n == 7);
        const mwSize *d = mxGetDimensions(a[0]);
        const int C = mxGetNumberOfDimensions(a[0]) == 3 ? (int)d[2] : 1;
        plhs[0] = sb_mex_matrix(d[0] * d[1], 1);
        sb_mex_check(sb_photo_unary((int)d[0], (int)d[1], C, mxGetPr(a[0]), mxGetPr(a[1]), mxGetPr(a[2]), mxGetPr(a[3]),
                                    scalar(a[4]), scalar(a[5]), scalar(a[6]), mxGetPr(plhs[0])));
    } else if (!strcmp(op, "pairwise_tables")) {
        SB_MEX_ASSERT(n == 8);
        const int H = (int)mxGetPr(a[0])[0], W = (int)mxGetPr(a[0])[1];
        const mwSize E = 2 * ((H - 1) * W + H * (W - 1));
        const bool has_prop = mxGetNumberOfElements(a[3]) > 0;
        // MATLAB hands over max(nlhs, 1) output slots: never write past them
        SB_MEX_ASSERT(has_prop ? nlhs == 4 : nlhs <= 1);
        for (int i = 0; i < (has_prop ? 4 : 1); i++) plhs[i] = sb_mex_matrix(1, E);
        sb_mex_check(sb_pairwise_tables(H, W, (int)scalar(a[1]), mxGetPr(a[2]), has_prop ? mxGetPr(a[3]) : NULL, mxGetPr(a[4]),
                                        scalar(a[5]), scalar(a[6]), scalar(a[7]), mxGetPr(plhs[0]),
                                        has_prop ? mxGetPr(plhs[1]) : NULL, has_prop ? mxGetPr(plhs[2]) : NULL,
                                        has_prop ? mxGetPr(plhs[3]) : NULL));
    } else if (!strcmp(op, "fusion_positions")) {
        SB_MEX_ASSERT(n == 4);
        const int H = (int)mxGetPr(a[0])[0], W = (int)mxGetPr(a[0])[1];
        const mwSize E = 2 * ((H - 1) * W + H * (W - 1));
        const int L = (int)(mxGetNumberOfElements(a[1]) / ((size_t)4 * H * W));
        SB_MEX_ASSERT(nlhs == 2);
        plhs[0] = sb_mex_matrix(L, E);
        plhs[1] = sb_mex_matrix(L, E);
        sb_mex_check(sb_fusion_positions(H, W, L, mxGetPr(a[1]), scalar(a[2]), scalar(a[3]), mxGetPr(plhs[0]), mxGetPr(plhs[1])));
    } else if (!strcmp(op, "energy")) {
        SB_MEX_ASSERT(n == 8);
        const int H = (int)mxGetPr(a[0])[0], W = (int)mxGetPr(a[0])[1];
        plhs[0] = mxCreateDoubleScalar(0);
        sb_mex_check(sb_energy(H, W, (int)scalar(a[1]), mxGetPr(a[2]), mxGetPr(a[3]), mxGetPr(a[4]), scalar(a[5]), scalar(a[6]),
                               scalar(a[7]), mxGetPr(plhs[0])));
    } else if (!strcmp(op, "segpln_wta")) {
        SB_MEX_ASSERT(n == 6 && nlhs <= 2);
        const mwSize nd = mxGetNumberOfDimensions(a[0]);
        const mwSize *d = mxGetDimensions(a[0]);
        const int H = (int)d[0], W = (int)d[1], C = nd >= 3 ? (int)d[2] : 1, ni = nd >= 4 ? (int)d[3] : 1;
        SB_MEX_ASSERT(mxGetNumberOfElements(a[1]) == (size_t)12 * ni);
        const int D = (int)mxGetNumberOfElements(a[2]), w = (int)scalar(a[3]);
        plhs[0] = sb_mex_matrix(H, W);
        mxArray *score = sb_mex_matrix(H - 2 * w, W - 2 * w);
        sb_mex_check(sb_segpln_wta(H, W, C, ni, mxGetPr(a[0]), mxGetPr(a[1]), D, mxGetPr(a[2]), w, scalar(a[4]), scalar(a[5]),
                                   mxGetPr(plhs[0]), mxGetPr(score)));
        if (nlhs == 2) plhs[1] = score; else mxDestroyArray(score);
    } else if (!strcmp(op, "plane_from_disparity")) {
        SB_MEX_ASSERT(n == 5 && nlhs <= 2);
        const int H = (int)mxGetM(a[0]), W = (int)mxGetN(a[0]);
        plhs[0] = sb_mex_matrix(4, 1);
        mxArray *prop = nlhs == 2 ? sb_mex_matrix(4, (mwSize)H * W) : NULL;
        sb_mex_check(sb_plane_from_disparity(H, W, mxGetPr(a[0]), scalar(a[1]), scalar(a[2]), scalar(a[3]), (int)scalar(a[4]), 0,
                                             mxGetPr(plhs[0]), prop ? mxGetPr(prop) : NULL, NULL));
        if (prop) plhs[1] = prop;
    } else if (!strcmp(op, "smooth_weights")) {
        SB_MEX_ASSERT(n == 4 && mxGetClassID(a[0]) == mxUINT32_CLASS);
        const int H = (int)mxGetM(a[0]), W = (int)mxGetN(a[0]);
        plhs[0] = sb_mex_matrix(1, 2 * ((mwSize)(H - 1) * W + (mwSize)H * (W - 1)));
        sb_mex_check(sb_smooth_weights(H, W, (const uint32_t *)mxGetData(a[0]), scalar(a[1]), scalar(a[2]), scalar(a[3]),
                                       mxGetPr(plhs[0])));
    } else if (!strcmp(op, "fuse_until_convergence")) {
        SB_MEX_ASSERT(n == 13 && nlhs <= 3 && mxGetClassID(a[12]) == mxINT32_CLASS);
        const int H = (int)mxGetPr(a[0])[0], W = (int)mxGetPr(a[0])[1];
        const size_t N = (size_t)H * W;
        const int np = (int)(mxGetNumberOfElements(a[2]) / (4 * N));
        SB_MEX_ASSERT(np >= 1 && mxGetNumberOfElements(a[3]) == N * np && mxGetNumberOfElements(a[4]) == 4 * N &&
                      mxGetNumberOfElements(a[5]) == N);
        const int maxiter = (int)scalar(a[11]);
        plhs[0] = mxDuplicateArray(a[4]);              // the library fuses in place: work on copies of the inputs
        mxArray *un = mxDuplicateArray(a[5]);
        mxArray *E = sb_mex_matrix(1, (mwSize)maxiter + 1);
        int ne = 0;
        sb_mex_check(sb_binary_fuse_until_convergence_grid(H, W, (int)scalar(a[1]), np, mxGetPr(a[2]), mxGetPr(a[3]), mxGetPr(plhs[0]),
                                                           mxGetPr(un), mxGetPr(a[6]), scalar(a[7]), scalar(a[8]), scalar(a[9]),
                                                           (int)scalar(a[10]), maxiter, (const int32_t *)mxGetData(a[12]),
                                                           (int64_t)mxGetNumberOfElements(a[12]), 0, mxGetPr(E), &ne, NULL));
        mxSetN(E, (mwSize)ne);
        if (nlhs >= 2) plhs[1] = E; else mxDestroyArray(E);
        if (nlhs >= 3) plhs[2] = un; else mxDestroyArray(un);
    } else {
        mexErrMsgTxt("sb_builders_mex: unknown operation");
    }
}
