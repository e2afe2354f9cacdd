// Dependent-chain latency of FP64 operations on sm_100a (one warp, clock64 around a chain).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void chain(double *out, double x0, double y, int n, long long *cyc) {
    double x = x0 + threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (OP == 0) x = pow(x, y) + 1.0;
        if (OP == 1) x = exp(-x * 1e-3) + 1.5;
        if (OP == 2) x = cbrt(x) + 2.0;
        if (OP == 3) x = 3.0 / x + 1.0;
        if (OP == 4) x = fma(x, 0.999, 0.01);
        if (OP == 5) x = sqrt(x) + 2.0;
        if (OP == 6) x = log(x) + 3.0;
        if (OP == 7) x = x * 0.999 + 0.01;   // -fmad=false: DMUL + DADD
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
    const char *names[] = {"pow(x,0.341)+1", "exp+add", "cbrt+add", "div+add", "dfma", "sqrt+add", "log+add", "dmul+dadd"};
    const int n = 2000;
    for (int w = 1; w <= 8; w *= 8)
    for (int op = 0; op < 8; op++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (op) {
                case 0: chain<0><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
                case 1: chain<1><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
                case 2: chain<2><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
                case 3: chain<3><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
                case 4: chain<4><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
                case 5: chain<5><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
                case 6: chain<6><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
                case 7: chain<7><<<1, 32 * w>>>(out, 2.5, 0.341, n, cyc); break;
            }
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("warps=%d %-16s %8.1f cycles per dependent op\n", w, names[op], (double)h / n);
    }
    return 0;
}
