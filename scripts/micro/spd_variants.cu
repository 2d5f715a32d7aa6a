// Tuning harness for the SPD Gram kernel: rebuilds gabotorch_b200/csrc/spd_gram.cu with other knob values
// (-DGABO_TILE_OVERHEAD_ROWS=..) and times SPD(3) Gram builds through the C ABI.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gabotorch_b200/csrc/spd_gram.cu"
#include "../../gabotorch_b200/csrc/host_common.cu"

static double urand() { return rand() / (double)RAND_MAX; }

int main(int argc, char** argv) {
    const int d = 3;
    for (int64_t N : {512, 1024, 1536, 2048, 3072, 4096, 8192}) {
        std::vector<double> h(N * d * d);
        srand(1);
        for (int64_t i = 0; i < N; ++i) {   // X = R R^T + 0.05 I, R uniform in [-1, 1]
            double R[3][3];
            for (auto& row : R) for (double& v : row) v = 2 * urand() - 1;
            for (int r = 0; r < d; ++r)
                for (int c = 0; c < d; ++c) {
                    double s = (r == c) ? 0.05 : 0.0;
                    for (int k = 0; k < d; ++k) s += R[r][k] * R[c][k];
                    h[(i * d + r) * d + c] = s;
                }
        }
        double *x, *f1, *f2, *out;
        int32_t* flags;
        const int64_t fs = gabo_spd_factor_stride(d);
        cudaMalloc(&x, sizeof(double) * N * d * d);
        cudaMalloc(&f1, sizeof(double) * N * fs);
        cudaMalloc(&f2, sizeof(double) * N * fs);
        cudaMalloc(&out, sizeof(double) * N * N);
        cudaMalloc(&flags, 4);
        cudaMemset(flags, 0, 4);
        cudaMemcpy(x, h.data(), sizeof(double) * N * d * d, cudaMemcpyHostToDevice);
        gabo_spd_factor2(x, N, x, N, d, 0, f1, f2, flags, nullptr);
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        for (int it = 0; it < 3; ++it) gabo_spd_ai_gram(f1, N, f2, N, d, 1.19, 0, GABO_F32, 0, out, GABO_F64, N, nullptr);
        float best = 1e9f;
        for (int rep = 0; rep < 10; ++rep) {
            cudaEventRecord(a);
            gabo_spd_ai_gram(f1, N, f2, N, d, 1.19, 0, GABO_F32, 0, out, GABO_F64, N, nullptr);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            best = fminf(best, ms);
        }
        printf("SPD(3) N=%lld: best %.4f ms, %.3e pairs/s (%s)\n", (long long)N, best, N * N / best * 1e3,
               cudaGetErrorString(cudaGetLastError()));
        cudaFree(x); cudaFree(f1); cudaFree(f2); cudaFree(out); cudaFree(flags);
    }
    return 0;
}
