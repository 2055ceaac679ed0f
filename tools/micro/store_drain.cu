// Microbenchmark (round 2): how long does it take an SM to push N bytes of results to L2 and make them visible (release), for
// the store shapes the persistent decode kernel could use?  48 CTAs x 512 threads (as the kernel), each CTA writes `bytes`
// to its own region, then releases.  Build + run on the GPU box:  nvcc -arch=sm_100a -O3 -o /tmp/store_drain tools/micro/store_drain.cu && /tmp/store_drain
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__global__ void k_stg32(float* out, int bytes_per_cta, long long* clk) {      // 4 warps store 128-byte rows with 4-byte lanes (v3/v4 staging)
    float* base = out + (size_t)blockIdx.x * (bytes_per_cta / 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < 4)
        for (int i = warp; i < bytes_per_cta / 128; i += 4) base[i * 32 + lane] = (float)i;
    const long long t1 = clock64();
    __threadfence();
    __syncthreads();
    const long long t2 = clock64();
    if (threadIdx.x == 0) { clk[blockIdx.x * 2] = t1 - t0; clk[blockIdx.x * 2 + 1] = t2 - t0; }
}
__global__ void k_stg128(float4* out, int bytes_per_cta, long long* clk, int nwarps) {   // 16-byte lanes: 512 B per warp instruction
    float4* base = out + (size_t)blockIdx.x * (bytes_per_cta / 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps)
        for (int i = warp; i < bytes_per_cta / 512; i += nwarps) base[i * 32 + lane] = make_float4(i, i, i, i);
    const long long t1 = clock64();
    __threadfence();
    __syncthreads();
    const long long t2 = clock64();
    if (threadIdx.x == 0) { clk[blockIdx.x * 2] = t1 - t0; clk[blockIdx.x * 2 + 1] = t2 - t0; }
}
__global__ void k_bulk(float* out, int bytes_per_cta, long long* clk) {       // cp.async.bulk shared -> global, 16 KB pieces
    extern __shared__ __align__(128) unsigned char sm[];
    float* base = out + (size_t)blockIdx.x * (bytes_per_cta / 4);
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = (float)i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
        for (int off = 0; off < bytes_per_cta; off += 16384)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 16384;" ::"l"(reinterpret_cast<char*>(base) + off),
                         "r"((unsigned)__cvta_generic_to_shared(sm))
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    const long long t1 = clock64();
    __threadfence();
    __syncthreads();
    const long long t2 = clock64();
    if (threadIdx.x == 0) { clk[blockIdx.x * 2] = t1 - t0; clk[blockIdx.x * 2 + 1] = t2 - t0; }
}

int main() {
    const int ctas = 48;
    long long* clk;
    cudaMallocManaged(&clk, ctas * 2 * sizeof(long long));
    float* buf;
    cudaMalloc(&buf, (size_t)ctas * 262144 * 2);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int bytes : {65536, 131072, 196608}) {
        for (int rep = 0; rep < 3; ++rep) {
            auto report = [&](const char* name) {
                cudaDeviceSynchronize();
                long long a = 0, b = 0;
                for (int i = 0; i < ctas; ++i) { a = a > clk[2 * i] ? a : clk[2 * i]; b = b > clk[2 * i + 1] ? b : clk[2 * i + 1]; }
                if (rep == 2) printf("%-34s %7d B/CTA: issue %7lld cycles, visible %7lld cycles (%.1f B/clk/SM)\n", name, bytes, a, b, (double)bytes / b);
            };
            k_stg32<<<ctas, 512, 200 * 1024>>>(buf, bytes, clk); report("STG.32 x32 lanes, 4 warps");
            k_stg128<<<ctas, 512, 200 * 1024>>>((float4*)buf, bytes, clk, 4); report("STG.128 x32 lanes, 4 warps");
            k_stg128<<<ctas, 512, 200 * 1024>>>((float4*)buf, bytes, clk, 16); report("STG.128 x32 lanes, 16 warps");
            k_bulk<<<ctas, 512, 200 * 1024>>>(buf, bytes, clk); report("cp.async.bulk smem->global 16 KB");
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
