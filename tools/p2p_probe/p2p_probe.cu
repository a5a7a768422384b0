// Ground truth for the multi-GPU exchange design: SM-driven peer reads / writes between two GPUs of
// the box (coalesced 16-byte loads per lane, coalesced stores, and scattered 16-byte stores), and the
// copy-engine bandwidth, for the transfer sizes of this path (1-64 MB).  nvcc -O3 -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_read(const float4* __restrict__ src, float4* __restrict__ dst, size_t n)
{
    float4 a = make_float4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const float4 v = src[i];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    dst[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = a;
}
// one element per thread, UNROLL independent loads in flight per thread
template<int U>
__global__ void k_read_once(const float4* __restrict__ src, float4* __restrict__ dst, size_t n)
{
    const size_t base = (blockIdx.x * (size_t)blockDim.x) * U + threadIdx.x;
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) { const size_t i = base + (size_t)u * blockDim.x; v[u] = i < n ? src[i] : make_float4(0, 0, 0, 0); }
    float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; u++) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
    dst[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = a;
}
__global__ void k_write(float4* __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = make_float4(1.f, 2.f, 3.f, (float)i);
}
__global__ void k_write_scatter(float4* __restrict__ dst, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const size_t j = (i * 2654435761ull) % n; // every lane of a warp hits a different line
        dst[j] = make_float4(1.f, 2.f, 3.f, 4.f);
    }
}

int main()
{
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    int can01 = 0, can10 = 0;
    cudaDeviceCanAccessPeer(&can01, 0, 1);
    cudaDeviceCanAccessPeer(&can10, 1, 0);
    int perf = -1, atom = -1;
    cudaDeviceGetP2PAttribute(&perf, cudaDevP2PAttrPerformanceRank, 0, 1);
    cudaDeviceGetP2PAttribute(&atom, cudaDevP2PAttrNativeAtomicSupported, 0, 1);
    printf("peer access 0->1 %d 1->0 %d perf rank %d native atomics %d\n", can01, can10, perf, atom);
    CK(cudaSetDevice(1));
    cudaDeviceEnablePeerAccess(0, 0);
    const size_t maxb = 256u << 20;
    float4 *remote, *local, *sink;
    CK(cudaMalloc(&remote, maxb));
    CK(cudaMemset(remote, 0, maxb));
    CK(cudaSetDevice(0));
    cudaDeviceEnablePeerAccess(1, 0);
    CK(cudaMalloc(&local, maxb));
    CK(cudaMalloc(&sink, 64u << 20));
    CK(cudaMemset(local, 0, maxb));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t sizes[] = { 1u << 20, 4u << 20, 16u << 20, 64u << 20, 256u << 20 };
    for (size_t bytes : sizes)
    {
        const size_t n = bytes / 16;
        const int reps = bytes <= (16u << 20) ? 20 : 5;
        float ms;
        auto report = [&](const char* what) {
            cudaEventElapsedTime(&ms, e0, e1);
            printf("%-34s %6.0f MB  %8.1f us  %7.1f GB/s\n", what, bytes / 1048576.0, ms * 1e3 / reps, bytes * reps / (ms * 1e-3) / 1e9);
        };
        const int grid = 148 * 8;
        for (int target = 0; target < 2; target++)
        {
            float4* p = target ? remote : local;
            const char* where = target ? "peer " : "local";
            char name[64];
            k_read<<<grid, 256>>>(p, sink, n);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) k_read<<<grid, 256>>>(p, sink, n);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            snprintf(name, 64, "%s read, grid-stride 1184x256", where); report(name);
            const unsigned g4 = (unsigned)((n + 256 * 4 - 1) / (256 * 4));
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) k_read_once<4><<<g4, 256>>>(p, sink, n);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            snprintf(name, 64, "%s read, 4 loads/thread one pass", where); report(name);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) k_write<<<grid, 256>>>(p, n);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            snprintf(name, 64, "%s write, coalesced", where); report(name);
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) k_write_scatter<<<grid, 256>>>(p, n);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            snprintf(name, 64, "%s write, scattered 16 B", where); report(name);
        }
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) cudaMemcpyPeerAsync(local, 0, remote, 1, bytes, 0);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        report("copy engine peer -> local");
    }
    return 0;
}
