// Microbenchmark (dev tool, not product): what does a pure WRITE stream reach on this B200, and how do
// plain vector stores compare with bulk asynchronous (TMA) shared->global stores?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o write_bw write_bw.cu && ./write_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void fill_stg(uint4 *dst, size_t n_vec) {
    const uint4 z = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) dst[i] = z;
}

__global__ void copy_ldst(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n_vec) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void read_only(const uint4 *__restrict__ src, uint4 *out, size_t n_vec) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = src[i]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if (acc.x == 0x12345678 && acc.y == 77) out[0] = acc;
}

template <int STAGES>
__global__ void fill_tma(uint8_t *dst, size_t n_tiles, int tile_bytes, int rezero) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    for (int i = tid; i < STAGES * tile_bytes / 16; i += blockDim.x) ((uint4 *)smem)[i] = make_uint4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    int it = 0;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        uint8_t *buf = smem + (size_t)(it % STAGES) * tile_bytes;
        if (rezero) {
            if (it >= STAGES) { if (tid == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory"); }
            __syncthreads();
            for (int i = tid; i < tile_bytes / 16; i += blockDim.x) ((uint4 *)buf)[i] = make_uint4(0, 0, 0, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
        if (tid == 0) {
            if (!rezero && it >= STAGES) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
            uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + tile * (size_t)tile_bytes), "r"(s), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
float time_ms(F f, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
    const size_t bytes = (size_t)1 << 20; const size_t total = bytes * 400;  // 419 MB, like one obs batch
    uint8_t *a, *b; CK(cudaMalloc(&a, total)); CK(cudaMalloc(&b, total)); CK(cudaMemset(a, 1, total)); CK(cudaMemset(b, 2, total));
    const size_t n_vec = total / 16;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, buffer %.1f MB\n", sms, total / 1e6);
    for (int mult : {4, 8, 16}) {
        float ms = time_ms([&] { fill_stg<<<sms * mult, 256>>>((uint4 *)a, n_vec); });
        printf("fill_stg   grid=%4d x256        : %7.1f us  %7.1f GB/s\n", sms * mult, ms * 1e3, total / ms / 1e6);
    }
    { float ms = time_ms([&] { cudaMemsetAsync(a, 0, total); }); printf("cudaMemset                      : %7.1f us  %7.1f GB/s\n", ms * 1e3, total / ms / 1e6); }
    for (int mult : {8, 16}) {
        float ms = time_ms([&] { copy_ldst<<<sms * mult, 256>>>((const uint4 *)a, (uint4 *)b, n_vec); });
        printf("copy_ldst  grid=%4d x256        : %7.1f us  %7.1f GB/s (r+w)\n", sms * mult, ms * 1e3, 2.0 * total / ms / 1e6);
    }
    { float ms = time_ms([&] { cudaMemcpyAsync(b, a, total, cudaMemcpyDeviceToDevice); }); printf("cudaMemcpy D2D                  : %7.1f us  %7.1f GB/s (r+w)\n", ms * 1e3, 2.0 * total / ms / 1e6); }
    for (int mult : {8, 16}) {
        float ms = time_ms([&] { read_only<<<sms * mult, 256>>>((const uint4 *)a, (uint4 *)b, n_vec); });
        printf("read_only  grid=%4d x256        : %7.1f us  %7.1f GB/s\n", sms * mult, ms * 1e3, total / ms / 1e6);
    }
    struct Cfg { int tile_bytes, ctas, threads; };
    for (int rezero : {0, 1})
        for (Cfg c : std::vector<Cfg>{{51200, 2, 128}, {51200, 1, 128}, {25600, 4, 64}, {25600, 2, 128}, {12800, 8, 64}, {12800, 4, 128}, {102400, 1, 256}, {6400, 8, 64}}) {
            const size_t n_tiles = total / c.tile_bytes;
            cudaFuncSetAttribute(fill_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * c.tile_bytes);
            float ms = time_ms([&] { fill_tma<2><<<sms * c.ctas, c.threads, 2 * c.tile_bytes>>>(a, n_tiles, c.tile_bytes, rezero); });
            CK(cudaGetLastError());
            printf("fill_tma<2> tile=%6d ctas/sm=%d thr=%3d rezero=%d : %7.1f us  %7.1f GB/s\n", c.tile_bytes, c.ctas, c.threads, rezero, ms * 1e3, total / ms / 1e6);
        }
    for (Cfg c : std::vector<Cfg>{{51200, 1, 128}, {25600, 2, 128}, {25600, 1, 128}}) {
        const size_t n_tiles = total / c.tile_bytes;
        cudaFuncSetAttribute(fill_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * c.tile_bytes);
        float ms = time_ms([&] { fill_tma<4><<<sms * c.ctas, c.threads, 4 * c.tile_bytes>>>(a, n_tiles, c.tile_bytes, 0); });
        CK(cudaGetLastError());
        printf("fill_tma<4> tile=%6d ctas/sm=%d thr=%3d rezero=0 : %7.1f us  %7.1f GB/s\n", c.tile_bytes, c.ctas, c.threads, ms * 1e3, total / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
