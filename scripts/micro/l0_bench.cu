// Micro-benchmark + self-check of the structured level-0 kernels (csrc/lm_l0.cu) on synthetic frame blocks, and two
// latency probes (dependent DFMA chain, IEEE fp64 division vs MUFU seed + Newton) that explain their critical path.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr [-DACINO_L0_TIMING] \
//        -o scripts/micro/l0_bench scripts/micro/l0_bench.cu && scripts/micro/l0_bench [frames]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../acinoset_b200/csrc/lm_l0.cu"

using namespace acino;

__global__ void probe_dfma(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, 0.999999, 1e-12);
    const long long t1 = clock64();
    out[0] = a;
    out[1] = (double)(t1 - t0) / iters;
}
__global__ void probe_div(double* out, int iters, int mode) {
    double a = 3.0 + threadIdx.x * 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (mode == 0) a = 1.0 / a + 2.5;
        else {              // MUFU seed + two Newton steps
            double r = (double)fast_rcp((float)a);
            r = fma(r, fma(-a, r, 1.0), r);
            r = fma(r, fma(-a, r, 1.0), r);
            a = r + 2.5;
        }
    }
    const long long t1 = clock64();
    out[0] = a;
    out[1] = (double)(t1 - t0) / iters;
}

#define CHECK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(_e), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 10000;
    const int M = (N + 2) / 3;
    // synthetic data blocks: H_n = G G^T + I (fp32, packed upper), gradient, no frozen variables
    std::vector<float> H((size_t)N * NU);
    std::vector<double> gt((size_t)N * NA), sw(NA), ctl(CTL_SIZE, 0.0);
    srand(1);
    for (int n = 0; n < N; ++n) {
        double G[NA][4];
        for (int i = 0; i < NA; ++i)
            for (int k = 0; k < 4; ++k) G[i][k] = 300.0 * (rand() / (double)RAND_MAX - 0.5);
        for (int i = 0; i < NA; ++i)
            for (int j = i; j < NA; ++j) {
                double s = i == j ? 50.0 : 0.0;
                for (int k = 0; k < 4; ++k) s += G[i][k] * G[j][k];
                H[(size_t)n * NU + upper_index(i, j)] = (float)s;
            }
        for (int i = 0; i < NA; ++i) gt[(size_t)n * NA + i] = 100.0 * (rand() / (double)RAND_MAX - 0.5);
    }
    for (int p = 0; p < NA; ++p) sw[p] = 2.0 / ((4.0 + 13.0 * p) * (4.0 + 13.0 * p)) * pow(120.0, 4);
    ctl[CTL_LAM] = 1e-3;
    std::vector<int> elim, surv;
    for (int e = 1; e < M; e += 2) { elim.push_back(e); elim.push_back(e - 1); elim.push_back(e + 1 < M ? e + 1 : -1); }
    for (int j = 0; j < M; j += 2) { surv.push_back(j); surv.push_back(j - 1 >= 0 ? j - 1 : -1); surv.push_back(j + 1 < M ? j + 1 : -1); }
    const int ne = (int)elim.size() / 3, ns = (int)surv.size() / 3;
    float* dH; double *dg, *dsw, *dctl, *dW, *dD, *dLc, *drhs, *dx; unsigned char* dfix; int *delim, *dsurv, *dinfo;
    CHECK(cudaMalloc(&dH, H.size() * 4)); CHECK(cudaMalloc(&dg, gt.size() * 8)); CHECK(cudaMalloc(&dsw, NA * 8));
    CHECK(cudaMalloc(&dctl, CTL_SIZE * 8)); CHECK(cudaMalloc(&dfix, (size_t)N * NA)); CHECK(cudaMemset(dfix, 0, (size_t)N * NA));
    const size_t blk = (size_t)SBN * SBN * 8;
    CHECK(cudaMalloc(&dW, M * blk)); CHECK(cudaMalloc(&dD, M * blk)); CHECK(cudaMalloc(&dLc, M * blk));
    CHECK(cudaMalloc(&drhs, (size_t)M * SBN * 8)); CHECK(cudaMalloc(&dx, (size_t)M * SBN * 8)); CHECK(cudaMemset(dx, 0, (size_t)M * SBN * 8));
    CHECK(cudaMalloc(&delim, elim.size() * 4 + 4)); CHECK(cudaMalloc(&dsurv, surv.size() * 4 + 4)); CHECK(cudaMalloc(&dinfo, 4)); CHECK(cudaMemset(dinfo, 0, 4));
    CHECK(cudaMemcpy(dH, H.data(), H.size() * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dg, gt.data(), gt.size() * 8, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(dsw, sw.data(), NA * 8, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dctl, ctl.data(), CTL_SIZE * 8, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(delim, elim.data(), elim.size() * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dsurv, surv.data(), surv.size() * 4, cudaMemcpyHostToDevice));
    const LmShard sh{N, 0, N};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[3] = {0, 0, 0};
    const int reps = 20;
    for (int which = 0; which < 3; ++which) {
        for (int rep = -3; rep < reps; ++rep) {
            if (rep == 0) CHECK(cudaEventRecord(e0));
            if (which == 0) CHECK(launch_l0_invert(sh, ne, delim, dH, dg, dfix, dsw, dctl, dW, drhs, dinfo, 0));
            if (which == 1) CHECK(launch_l0_update(sh, ns, dsurv, dH, dg, dfix, dsw, dctl, dW, dD, dLc, drhs, 0));
            if (which == 2) CHECK(launch_l0_backsub(sh, ne, delim, dfix, dsw, dW, drhs, dx, 0));
        }
        CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms[which], e0, e1);
        ms[which] /= reps;
    }
    printf("frames %d  blocks %d  (eliminated %d, surviving %d)\n", N, M, ne, ns);
    printf("l0_invert  %8.1f us   l0_update %8.1f us   l0_backsub %8.1f us\n", 1e3 * ms[0], 1e3 * ms[1], 1e3 * ms[2]);
    // ---- self-check of block e = 1: W D_e = I with D_e rebuilt on the host
    std::vector<double> W((size_t)SBN * SBN);
    CHECK(cudaMemcpy(W.data(), dW + (size_t)SBN * SBN, blk, cudaMemcpyDeviceToHost));
    int info = 0; CHECK(cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost));
    if (M > 1) {
        std::vector<double> De((size_t)SBN * SBN, 0.0);
        for (int r = 0; r < SBN; ++r)
            for (int c = 0; c < SBN; ++c) {
                const int a = r / NA, p = r % NA, b = c / NA, q = c % NA, na = 3 + a, nb = 3 + b;
                double v = 0.0;
                if (na >= N || nb >= N) v = r == c ? 1.0 : 0.0;
                else if (a == b) {
                    v = H[(size_t)na * NU + upper_index(p < q ? p : q, p < q ? q : p)];
                    if (p == q) v = (v + band_coef(na, 0, N) * sw[p]) * (1.0 + ctl[CTL_LAM]);
                } else if (p == q) v = band_coef(a < b ? na : nb, abs(a - b), N) * sw[p];
                De[(size_t)r * SBN + c] = v;
            }
        double worst = 0.0;
        for (int r = 0; r < SBN; ++r)
            for (int c = 0; c < SBN; ++c) {
                double s = 0.0;
                for (int k = 0; k < SBN; ++k) s += W[(size_t)r * SBN + k] * De[(size_t)k * SBN + c];
                worst = fmax(worst, fabs(s - (r == c ? 1.0 : 0.0)));
            }
        printf("self-check block 1: max |W D - I| = %.3e, info = %d\n", worst, info);
    }
#ifdef ACINO_L0_TIMING
    {
        acino_debug_l0_reset();
        CHECK(launch_l0_invert(sh, 1, delim, dH, dg, dfix, dsw, dctl, dW, drhs, dinfo, 0));
        CHECK(cudaDeviceSynchronize());
        long long cy[8]; acino_debug_l0_cycles(cy);
        printf("l0_invert single block cycles: stage+build+inv0 %lld | panel products %lld | update+publish+inv %lld | tail %lld | total %lld\n",
               cy[0], cy[1], cy[2], cy[3], cy[0] + cy[1] + cy[2] + cy[3]);
    }
#endif
    double* dp; CHECK(cudaMalloc(&dp, 16));
    double hp[2];
    probe_dfma<<<1, 32>>>(dp, 100000); CHECK(cudaMemcpy(hp, dp, 16, cudaMemcpyDeviceToHost));
    printf("dependent DFMA: %.1f cycles per instruction\n", hp[1]);
    for (int mode = 0; mode < 2; ++mode) {
        probe_div<<<1, 32>>>(dp, 100000, mode); CHECK(cudaMemcpy(hp, dp, 16, cudaMemcpyDeviceToHost));
        printf("dependent %s + DADD: %.1f cycles\n", mode == 0 ? "1.0/x (IEEE)" : "MUFU seed + 2 Newton", hp[1]);
    }
    return 0;
}
