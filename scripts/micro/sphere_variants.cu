// Tuning harness for the sphere Gram kernel: rebuilds gabotorch_b200/csrc/sphere_gram.cu with other knob values
// (-DGABO_SG_UNROLL=.., -DGABO_SG_TILEM=.., -DGABO_SG_MINBLOCKS=..) and times the N = 32768 launches through the C ABI.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DGABO_SG_UNROLL=4 \
//        scripts/micro/sphere_variants.cu -o scripts/micro/sphere_u4
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gabotorch_b200/csrc/sphere_gram.cu"
#include "../../gabotorch_b200/csrc/host_common.cu"

int main(int argc, char** argv) {
    const int64_t N = argc > 1 ? atoll(argv[1]) : 32768;
    for (int D : {3, 9}) {
        std::vector<double> h(N * D);
        srand(1);
        for (int64_t i = 0; i < N; ++i) {
            double s = 0;
            for (int k = 0; k < D; ++k) {
                double u = 0;
                for (int t = 0; t < 12; ++t) u += rand() / (double)RAND_MAX;
                h[i * D + k] = u - 6.0;
                s += h[i * D + k] * h[i * D + k];
            }
            for (int k = 0; k < D; ++k) h[i * D + k] /= sqrt(s);
        }
        double* x;
        cudaMalloc(&x, sizeof(double) * N * D);
        cudaMemcpy(x, h.data(), sizeof(double) * N * D, cudaMemcpyHostToDevice);
        for (int dt : {GABO_F32, GABO_F64}) {
            const size_t es = dt == GABO_F32 ? 4 : 8;
            void* out;
            cudaMalloc(&out, es * N * N);
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            for (int it = 0; it < 2; ++it) gabo_sphere_gram(x, N, x, N, D, 6.5 + log(2.0), 0, out, dt, N, nullptr);
            cudaEventRecord(a);
            for (int it = 0; it < 5; ++it) gabo_sphere_gram(x, N, x, N, D, 6.5 + log(2.0), 0, out, dt, N, nullptr);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            ms /= 5;
            printf("D=%d out=%s: %.3f ms, %.0f GB/s (%s)\n", D, dt == GABO_F32 ? "f32" : "f64", ms, N * N * es / ms / 1e6,
                   cudaGetErrorString(cudaGetLastError()));
            cudaFree(out);
        }
        cudaFree(x);
    }
    return 0;
}
