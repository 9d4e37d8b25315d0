// l2_probe.cu -- how much data stays resident in the B200 L2 when every SM re-reads the same working set?
//   ./l2_probe            prints, per working-set size, the time of 8 passes in two access patterns
//   ncu --metrics dram__bytes_read.sum ./l2_probe    gives the DRAM bytes per launch: ~S when the set stays in L2,
//                                                    ~8 S when every pass misses
// pattern 0 ("broadcast"): every CTA reads the whole buffer (like the digit planes of an eigen-tile group, read by all clusters)
// pattern 1 ("partitioned"): the buffer is split over the CTAs (each byte is read by one SM per pass)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_probe tools/l2_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

__global__ void __launch_bounds__(512) probe(const int4* __restrict__ buf, size_t n16, int passes, int pattern, int* sink)
{
    int acc = 0;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        if (pattern == 0) {
            // every CTA walks the whole buffer, CTAs staggered so that they do not all hit the same line at once
            const size_t off = ((size_t)blockIdx.x * 9973u * 64u) % n16;
            for (size_t i = threadIdx.x; i < n16; i += blockDim.x) {
                size_t j = i + off;
                if (j >= n16) j -= n16;
                const int4 v = __ldg(buf + j);
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        } else {
            for (size_t i = tid; i < n16; i += nthr) {
                const int4 v = __ldg(buf + i);
                acc ^= v.x ^ v.y ^ v.z ^ v.w;
            }
        }
    }
    if (acc == 0x7fffffff) *sink = acc;
}

int main(int argc, char** argv)
{
    const int passes = 8;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t maxb = (size_t)512 << 20;
    int4* buf;
    int* sink;
    cudaMalloc(&buf, maxb);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, maxb);
    const int sizes_mb[] = {8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 96, 112, 128, 160, 256};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int pattern = 0; pattern < 2; ++pattern)
        for (int s : sizes_mb) {
            const size_t n16 = ((size_t)s << 20) / 16;
            const int p = pattern == 0 ? 2 : passes;   // broadcast: 148 x S bytes per pass through the L2 already
            probe<<<sms, 512>>>(buf, n16, 1, pattern, sink);   // warm
            cudaEventRecord(e0);
            probe<<<sms, 512>>>(buf, n16, p, pattern, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double sm_bytes = (pattern == 0 ? (double)sms : 1.0) * (double)s * 1048576.0 * p;
            printf("pattern %d  S = %3d MB  passes %d  %8.3f ms  L2->SM %8.1f GB/s  (unique bytes per pass / time: %7.1f GB/s)\n", pattern, s, p,
                   ms, sm_bytes / ms / 1e6, (double)s * 1048576.0 * p / ms / 1e6);
        }
    return 0;
}
