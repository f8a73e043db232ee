/* interp2_port.c -- plain-C restatement of vgg_interp2's linear path
 * (imrender/vgg/vgg_interp2.cxx:246-322), double in / double out.
 * TEST INFRASTRUCTURE ONLY (see trws_port.c header).  A is h x w x c column-major,
 * X (column) and Y (row) are 1-based; output n x c column-major. */
#include <stdint.h>

int port_interp2_linear(const double *A, int h, int w, int c, const double *X, const double *Y, int64_t n,
                        double oobv, double *B)
{
    const int64_t end = n * c, step = (int64_t)h * w;
    const double dw = (double)w, dh = (double)h;
    for (int64_t i = 0; i < n; i++) {
        int64_t j, k;
        if (X[i] >= 1 && Y[i] >= 1) {
            if (X[i] < dw) {
                if (Y[i] < dh) {
                    int x = (int)X[i], y = (int)Y[i];
                    double u = X[i] - x, v = Y[i] - y;
                    k = (int64_t)h * (x - 1) + y - 1;
                    for (j = i; j < end; j += n, k += step) {
                        double out = A[k] + (A[k + h] - A[k]) * u;
                        out += ((A[k + 1] - out) + (A[k + h + 1] - A[k + 1]) * u) * v;
                        B[j] = out;
                    }
                } else if (Y[i] == dh) {
                    int x = (int)X[i];
                    double u = X[i] - x;
                    k = (int64_t)h * x - 1;
                    for (j = i; j < end; j += n, k += step) B[j] = A[k] + (A[k + h] - A[k]) * u;
                } else {
                    for (j = i; j < end; j += n) B[j] = oobv;
                }
            } else if (X[i] == dw) {
                if (Y[i] < dh) {
                    int y = (int)Y[i];
                    double v = Y[i] - y;
                    k = (int64_t)h * (w - 1) + y - 1;
                    for (j = i; j < end; j += n, k += step) B[j] = A[k] + (A[k + 1] - A[k]) * v;
                } else if (Y[i] == dh) {
                    k = (int64_t)h * w - 1;
                    for (j = i; j < end; j += n, k += step) B[j] = A[k];
                } else {
                    for (j = i; j < end; j += n) B[j] = oobv;
                }
            } else {
                for (j = i; j < end; j += n) B[j] = oobv;
            }
        } else {
            for (j = i; j < end; j += n) B[j] = oobv;
        }
    }
    return 0;
}
