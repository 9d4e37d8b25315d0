// tools/peak_fp64.cu -- measures the roofline denominators MEASURED_PEAKS.json does not carry:
// FP64 FMA (CUDA-core) rate, FP64 DMMA (mma.sync m8n8k4) rate, cuBLAS DGEMM and cuBLAS int8 GEMM.
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_fp64 tools/peak_fp64.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void dfma_kernel(double* out, int iters, double a, double b)
{
    double r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters)
{
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
    // DFMA
    double best_dfma = 0, best_dmma = 0;
    for (int rep = 0; rep < 5; ++rep) {
        const int iters = 20000, blocks = sms * 8, threads = 512;
        cudaEventRecord(e0); dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        double fl = 2.0 * 16 * iters * (double)blocks * threads;
        double tf = fl / (time_ms(e0, e1) * 1e-3) / 1e12; if (tf > best_dfma) best_dfma = tf;
    }
    for (int rep = 0; rep < 5; ++rep) {
        const int iters = 20000, blocks = sms * 8, threads = 256;
        cudaEventRecord(e0); dmma_kernel<<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        double fl = 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (threads / 32);
        double tf = fl / (time_ms(e0, e1) * 1e-3) / 1e12; if (tf > best_dmma) best_dmma = tf;
    }
    // cuBLAS DGEMM 8192^3
    cublasHandle_t h; cublasCreate(&h);
    const int N = 8192;
    double *A, *B, *C; cudaMalloc(&A, sizeof(double) * N * N); cudaMalloc(&B, sizeof(double) * N * N); cudaMalloc(&C, sizeof(double) * N * N);
    cudaMemset(A, 0, sizeof(double) * N * N); cudaMemset(B, 0, sizeof(double) * N * N);
    const double one = 1, zero = 0;
    double best_dgemm = 0, sust_dgemm = 0;
    cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, N, N, &one, A, N, B, N, &zero, C, N); cudaDeviceSynchronize();
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, N, N, &one, A, N, B, N, &zero, C, N); cudaEventRecord(e1); cudaEventSynchronize(e1);
        double tf = 2.0 * N * (double)N * N / (time_ms(e0, e1) * 1e-3) / 1e12; if (tf > best_dgemm) best_dgemm = tf;
    }
    {
        cudaEventRecord(e0);
        for (int rep = 0; rep < 40; ++rep) cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, N, N, &one, A, N, B, N, &zero, C, N);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        sust_dgemm = 40 * 2.0 * N * (double)N * N / (time_ms(e0, e1) * 1e-3) / 1e12;
    }
    // cuBLAS int8 GEMM (int32 accumulate) 16384 x 8192 x 8192, TN layout
    int8_t *A8, *B8; int32_t* C32; const int M8 = 16384, N8 = 8192, K8 = 8192;
    cudaMalloc(&A8, (size_t)M8 * K8); cudaMalloc(&B8, (size_t)N8 * K8); cudaMalloc(&C32, sizeof(int32_t) * (size_t)M8 * N8);
    cudaMemset(A8, 1, (size_t)M8 * K8); cudaMemset(B8, 1, (size_t)N8 * K8);
    const int32_t ione = 1, izero = 0; double best_i8 = 0, sust_i8 = 0; int i8_status = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        cublasStatus_t st = cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, M8, N8, K8, &ione, A8, CUDA_R_8I, K8, B8, CUDA_R_8I, K8, &izero, C32, CUDA_R_32I, M8, CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        i8_status = (int)st;
        double tf = 2.0 * M8 * (double)N8 * K8 / (time_ms(e0, e1) * 1e-3) / 1e12; if (rep > 0 && tf > best_i8) best_i8 = tf;
    }
    {
        cudaEventRecord(e0);
        for (int rep = 0; rep < 100; ++rep) cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, M8, N8, K8, &ione, A8, CUDA_R_8I, K8, B8, CUDA_R_8I, K8, &izero, C32, CUDA_R_32I, M8, CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        sust_i8 = 100 * 2.0 * M8 * (double)N8 * K8 / (time_ms(e0, e1) * 1e-3) / 1e12;
    }
    // same int8 GEMM with B handed over transposed (MN-major), the layout the rotation uses since r01-v4
    double best_i8tt = 0, sust_i8tt = 0; int i8tt_status = 0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        cublasStatus_t st = cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_T, M8, N8, K8, &ione, A8, CUDA_R_8I, K8, B8, CUDA_R_8I, N8, &izero, C32, CUDA_R_32I, M8, CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        i8tt_status = (int)st;
        double tf = 2.0 * M8 * (double)N8 * K8 / (time_ms(e0, e1) * 1e-3) / 1e12; if (rep > 0 && tf > best_i8tt) best_i8tt = tf;
    }
    {
        cudaEventRecord(e0);
        for (int rep = 0; rep < 100; ++rep) cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_T, M8, N8, K8, &ione, A8, CUDA_R_8I, K8, B8, CUDA_R_8I, N8, &izero, C32, CUDA_R_32I, M8, CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        sust_i8tt = 100 * 2.0 * M8 * (double)N8 * K8 / (time_ms(e0, e1) * 1e-3) / 1e12;
    }
    printf("{\"int8_gemm_tt_tops\": %.1f, \"int8_gemm_tt_tops_sustained\": %.1f, \"int8_tt_status\": %d}\n", best_i8tt, sust_i8tt, i8tt_status);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fp64_dmma_tflops\": %.2f, \"dgemm_tflops\": %.2f, \"dgemm_tflops_sustained\": %.2f, \"int8_gemm_tops\": %.1f, \"int8_gemm_tops_sustained\": %.1f, \"int8_status\": %d}\n",
           prop.name, sms, best_dfma, best_dmma, best_dgemm, sust_dgemm, best_i8, sust_i8, i8_status);
    return 0;
}
