// Tuning / ablation harness for the nested projection kernel: rebuilds gabotorch_b200/csrc/nested_project.cu with
//   -DGABO_NP_ROWGROUPS=2|4   (32-row tiles x 2 CTAs per SM, or 64-row tiles x 1 CTA)
//   -DGABO_NP_ABLATE=1        (skip the mma.sync instructions: loads, 3xTF32 split, epilogue and stores only)
//   -DGABO_NP_ABLATE=2        (skip the split as well: one mma per k-step, i.e. plain TF32)
// and times SPD(20) -> SPD(5) on N = 2^20 Mandel vectors.  Answers "would a faster MMA instruction move this kernel?".
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gabotorch_b200/csrc/nested_project.cu"
#include "../../gabotorch_b200/csrc/host_common.cu"

int main() {
    const int D = 20, d = 5;
    const int64_t N = 1 << 20;
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    std::vector<double> w(D * d);
    srand(3);
    for (auto& v : w) v = rand() / (double)RAND_MAX - 0.5;
    double* wd;
    float *x, *y, *pack;
    cudaMalloc(&wd, sizeof(double) * D * d);
    cudaMemcpy(wd, w.data(), sizeof(double) * D * d, cudaMemcpyHostToDevice);
    cudaMalloc(&x, sizeof(float) * N * dvh);
    cudaMalloc(&y, sizeof(float) * N * dvl);
    cudaMemset(x, 0x3c, sizeof(float) * N * dvh);      // 0x3c3c3c3c = 0.0115 as float: finite data
    cudaMalloc(&pack, sizeof(float) * gabo_nested_projection_pack_size(D, d));
    gabo_nested_projection_matrix(wd, D, d, pack, nullptr);
    for (int it = 0; it < 3; ++it) gabo_nested_spd_project(x, N, D, d, pack, y, nullptr);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e9f;
    for (int rep = 0; rep < 10; ++rep) {
        cudaEventRecord(a);
        gabo_nested_spd_project(x, N, D, d, pack, y, nullptr);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        best = fminf(best, ms);
    }
    printf("N=%lld: best %.4f ms, %.0f GB/s (%s; %s)\n", (long long)N, best, N * 4.0 * (dvh + dvl) / best / 1e6,
           cudaGetErrorString(cudaGetLastError()), gabo_last_error());
    return 0;
}
