// Second set of micro-benchmarks (B200): what bounds the scalar-flux tally and the q-bar gather.
//   red_cluster : red.global.add.f64 where the 32 lanes of a warp hit NS distinct 32-byte sectors
//                 (32/NS distinct doubles per sector) of an L2-resident array of n_reg doubles
//   ld_cluster  : same pattern with 8-byte loads (q-bar gather)
//   atoms_f64   : atomicAdd(double) on shared memory, random addresses in a tile
//   smem_rmw    : non-atomic LDS + DADD + STS on shared memory, random addresses (ownership model)
//   stream      : coalesced 16-byte loads over a buffer that fits L2 (90 MB) / does not (1 GB)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench2 ubench2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE> // 0 red, 1 load
__global__ void k_cluster(double* arr, int n_sector, int NS, int K, double* out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int per = 32 / NS;           // lanes per sector (1, 2, 4)
    const int grp = lane / per, sub = lane % per;
    uint32_t s = hash32(warp * 2654435761U + grp * 40503U + 17U);
    double acc = 0.0;
    for (int k = 0; k < K; k++) {
        s = hash32(s + k);
        const size_t idx = (size_t)(s % (uint32_t)n_sector) * 4 + sub;
        if (MODE == 0) atomicAdd(&arr[idx], 1.0);
        else acc += arr[idx];
    }
    if (MODE == 1) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> // 0 atomicAdd shared, 1 plain rmw
__global__ void k_smem(double* out, int T, int K) {
    extern __shared__ double tile[];
    for (int i = threadIdx.x; i < T; i += blockDim.x) tile[i] = 0.0;
    __syncthreads();
    uint32_t s = hash32(blockIdx.x * blockDim.x + threadIdx.x);
    for (int k = 0; k < K; k++) {
        s = hash32(s + k);
        const int idx = s % (uint32_t)T;
        if (MODE == 0) atomicAdd(&tile[idx], 1.0);
        else tile[idx] += 1.0;
    }
    __syncthreads();
    double a = 0.0;
    for (int i = threadIdx.x; i < T; i += blockDim.x) a += tile[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

__global__ void k_stream(const double2* __restrict__ src, size_t n, double* out) {
    double a = 0.0;
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double2 v0 = src[i], v1 = src[i + stride], v2 = src[i + 2 * stride], v3 = src[i + 3 * stride];
        a += v0.x + v0.y + v1.x + v1.y + v2.x + v2.y + v3.x + v3.y;
    }
    for (; i < n; i += stride) { double2 v = src[i]; a += v.x + v.y; }
    if (a == 123.456) out[0] = a;
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    const int n_reg = 86992, n_sector = n_reg / 4;
    double *arr, *out; CK(cudaMalloc(&arr, n_reg * 8)); CK(cudaMalloc(&out, (size_t)sms * 1024 * 8 * 4));
    CK(cudaMemset(arr, 0, n_reg * 8));
    const int K = 64;
    for (int block : {512, 1024}) {
        const int grid = sms * (1024 / block) * 2;
        const double lanes = (double)grid * block * K;
        for (int NS : {32, 16, 8}) {
            float ms = timeit([&] { k_cluster<0><<<grid, block>>>(arr, n_sector, NS, K, out); });
            printf("{\"bench\": \"red_cluster\", \"block\": %d, \"sectors_per_instr\": %d, \"lanes_per_sector\": %d, \"ms\": %.4f, \"Gred_per_s\": %.1f, \"cyc_per_warp_instr_per_sm\": %.1f}\n",
                   block, NS, 32 / NS, ms, lanes / ms * 1e-6, ms * 1e-3 * 1.965e9 * sms / (lanes / 32));
            ms = timeit([&] { k_cluster<1><<<grid, block>>>(arr, n_sector, NS, K, out); });
            printf("{\"bench\": \"ld_cluster\", \"block\": %d, \"sectors_per_instr\": %d, \"ms\": %.4f, \"Gld_per_s\": %.1f, \"cyc_per_warp_instr_per_sm\": %.1f}\n",
                   block, NS, ms, lanes / ms * 1e-6, ms * 1e-3 * 1.965e9 * sms / (lanes / 32));
        }
    }
    for (int T : {4096, 9728}) {
        CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 8));
        CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, T * 8));
        const int Ks = 256, block = 512, grid = sms;
        const double lanes = (double)grid * block * Ks;
        float ms = timeit([&] { k_smem<0><<<grid, block, T * 8>>>(out, T, Ks); });
        printf("{\"bench\": \"atoms_f64\", \"T\": %d, \"ms\": %.4f, \"Gatom_per_s\": %.1f, \"cyc_per_warp_instr_per_sm\": %.1f}\n", T, ms, lanes / ms * 1e-6, ms * 1e-3 * 1.965e9 * sms / (lanes / 32));
        ms = timeit([&] { k_smem<1><<<grid, block, T * 8>>>(out, T, Ks); });
        printf("{\"bench\": \"smem_rmw\", \"T\": %d, \"ms\": %.4f, \"Grmw_per_s\": %.1f, \"cyc_per_warp_instr_per_sm\": %.1f}\n", T, ms, lanes / ms * 1e-6, ms * 1e-3 * 1.965e9 * sms / (lanes / 32));
    }
    for (size_t mb : {64, 96, 1024}) {
        const size_t n = mb * 1024 * 1024 / 16;
        double2* src; CK(cudaMalloc(&src, n * 16)); CK(cudaMemset(src, 0, n * 16));
        for (int block : {512, 1024}) {
            const int grid = sms * (2048 / block);
            float ms = timeit([&] { k_stream<<<grid, block>>>(src, n, out); }, 8);
            printf("{\"bench\": \"stream\", \"MB\": %zu, \"block\": %d, \"ms\": %.4f, \"GB_per_s\": %.1f}\n", mb, block, ms, n * 16.0 / ms * 1e-6);
        }
        CK(cudaFree(src));
    }
    return 0;
}
