// Accuracy of the fp32 softplus(-|x|) = log1p(exp(-|x|)) formulations of polar_scl_fast.cuh against fp64,
// over a dense grid of |x|.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o _variants/softplus_accuracy scripts/softplus_accuracy.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
__device__ float ex2f_(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float rcpf_(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float lg2f_(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ float sp_series(float ax)
{
    const float u = ex2f_(ax * -1.4426950408889634f);
    const float s = u * rcpf_(2.0f + u);
    const float s2 = s * s;
    float pl = fmaf(s2, 0.07692308f, 0.09090909f);
    pl = fmaf(s2, pl, 0.11111111f);
    pl = fmaf(s2, pl, 0.14285715f);
    pl = fmaf(s2, pl, 0.2f);
    pl = fmaf(s2, pl, 0.33333334f);
    pl = fmaf(s2, pl, 1.0f);
    return 2.0f * s * pl;
}
__device__ float sp_lg2(float ax)
{
    const float u = ex2f_(ax * -1.4426950408889634f);
    return lg2f_(1.0f + u) * 0.6931471805599453f;
}
__global__ void sweep(int n, float hi, double* out)
{
    double e1 = 0, e2 = 0, s1 = 0, s2 = 0, q1 = 0, q2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float ax = hi * (float)i / (float)n;
        const double ref = log1p(exp(-(double)ax));
        const double d1 = (double)sp_series(ax) - ref, d2 = (double)sp_lg2(ax) - ref;
        e1 = fmax(e1, fabs(d1)); e2 = fmax(e2, fabs(d2));
        s1 += d1; s2 += d2; q1 += d1 * d1; q2 += d2 * d2;
    }
    double* o = out + 6 * (blockIdx.x * blockDim.x + threadIdx.x);
    o[0] = e1; o[1] = e2; o[2] = s1; o[3] = s2; o[4] = q1; o[5] = q2;
}
int main()
{
    const int T = 148 * 256, n = 1 << 26;
    double* d; cudaMalloc(&d, T * 6 * sizeof(double));
    for (float hi : {2.0f, 8.0f, 32.0f, 128.0f}) {
        sweep<<<148, 256>>>(n, hi, d);
        double* h = new double[T * 6];
        cudaMemcpy(h, d, T * 6 * sizeof(double), cudaMemcpyDeviceToHost);
        double e1 = 0, e2 = 0, s1 = 0, s2 = 0, q1 = 0, q2 = 0;
        for (int t = 0; t < T; t++) { e1 = fmax(e1, h[6*t]); e2 = fmax(e2, h[6*t+1]); s1 += h[6*t+2]; s2 += h[6*t+3]; q1 += h[6*t+4]; q2 += h[6*t+5]; }
        printf("{\"range\": [0, %g], \"series\": {\"max_abs\": %.3e, \"mean\": %.3e, \"rms\": %.3e}, \"lg2\": {\"max_abs\": %.3e, \"mean\": %.3e, \"rms\": %.3e}}\n",
               hi, e1, s1 / n, sqrt(q1 / n), e2, s2 / n, sqrt(q2 / n));
        delete[] h;
    }
    return 0;
}
