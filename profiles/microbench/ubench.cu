// Micro-benchmarks that size the design of the MoC sweep kernel on B200:
//   (1) FP64 red.global.add throughput for the tally patterns the sweep produces
//   (2) FP64 FMA issue rate
//   (3) shared-memory exponential-table lookup rate (random index, 8B vs 16B entries)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// pattern: each "item" (ray) visits K random regions; GL lanes per item hit GL consecutive
// doubles of flux[reg][GS] (GS = group stride). share = number of adjacent items that visit
// the same region at the same step (models adjacent parallel rays crossing the same FSR).
template <bool RED>
__global__ void k_atomic(double* flux, int n_reg, int GL, int GS, int K, int share, int n_items) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int item = t / GL, g = t % GL;
    if (item >= n_items) return;
    uint32_t s = hash32((uint32_t)(item / share) * 2654435761U + 12345U);
    double v = 1.0 + g;
    for (int k = 0; k < K; k++) {
        s = hash32(s + k);
        int reg = s % (uint32_t)n_reg;
        if (RED) atomicAdd(&flux[(size_t)reg * GS + g], v);
        else flux[(size_t)reg * GS + g] = v;   // plain scattered store for comparison
    }
}

// warp-aggregated variant: lanes compare region ids with match_any and one lane per distinct id adds
__global__ void k_atomic_match(double* flux, int n_reg, int K, int share, int n_items) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int item = t;
    if (item >= n_items) return;
    uint32_t s = hash32((uint32_t)(item / share) * 2654435761U + 12345U);
    double v = 1.0;
    int lane = threadIdx.x & 31;
    for (int k = 0; k < K; k++) {
        s = hash32(s + k);
        int reg = s % (uint32_t)n_reg;
        unsigned m = __match_any_sync(__activemask(), reg);
        // reduce v over the lanes in m (generic loop over set bits)
        double acc = 0.0;
        unsigned mm = m;
        while (mm) { int l = __ffs(mm) - 1; mm &= mm - 1; acc += __shfl_sync(m, v, l); }
        if (lane == __ffs(m) - 1) atomicAdd(&flux[reg], acc);
    }
}

__global__ void k_fma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// dependent-chain latency of DFMA/DADD/DMUL
__global__ void k_lat(double* out, int iters, long long* cyc) {
    double a = threadIdx.x * 1e-3; const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a = fma(a, b, c); }
    long long t1 = clock64();
    double d = a;
    for (int i = 0; i < iters; i++) { d = d - (d - c) * b; }
    long long t2 = clock64();
    out[threadIdx.x] = a + d; if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}

// exp table lookups. MODE 0: d[i], d[i+1] as two 8B LDS; MODE 1: double2 {d[i], d[i+1]-d[i]} one 16B LDS
template <int MODE>
__global__ void k_exptab(const double* tab_g, double* out, int iters, int n_tab) {
    extern __shared__ double tab[];
    if (MODE == 0) { for (int i = threadIdx.x; i < n_tab + 2; i += blockDim.x) tab[i] = tab_g[i]; }
    else { for (int i = threadIdx.x; i < n_tab + 1; i += blockDim.x) { tab[2*i] = tab_g[i]; tab[2*i+1] = tab_g[i+1] - tab_g[i]; } }
    __syncthreads();
    uint32_t s = hash32(blockIdx.x * blockDim.x + threadIdx.x);
    double acc = 0.0;
    const double vmin = -10.0, space = 1e-3, rspace = 1000.0;
    for (int k = 0; k < iters; k++) {
        s = s * 1664525U + 1013904223U;
        double v = -(double)(s >> 8) * (3.0 / 16777216.0);   // tau in [0,3)
        int i = (int)((v - vmin) * rspace);
        double r = v - (space * i + vmin);
        if (MODE == 0) acc += tab[i] + (tab[i + 1] - tab[i]) * r * rspace;
        else { double2 e = reinterpret_cast<const double2*>(tab)[i]; acc += e.x + e.y * r * rspace; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
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
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
    const int n_reg = 86989;
    double* flux; CK(cudaMalloc(&flux, (size_t)n_reg * 8 * sizeof(double) * 2)); CK(cudaMemset(flux, 0, (size_t)n_reg * 8 * 8 * 2));
    // (1) atomics
    struct Cfg { int GL, GS, share; } cfgs[] = { {1,1,1}, {1,1,4}, {1,1,8}, {1,1,32}, {7,7,1}, {7,8,1}, {8,8,1}, {7,7,4}, {1,7,1}, {1,8,1} };
    for (auto c : cfgs) {
        int n_items = 26448 * 2, K = 344;
        long long nthreads = (long long)n_items * c.GL;
        int bs = 256; int gs = (int)((nthreads + bs - 1) / bs);
        float ms = timeit([&] { k_atomic<true><<<gs, bs>>>(flux, n_reg, c.GL, c.GS, K, c.share, n_items); });
        float ms_st = timeit([&] { k_atomic<false><<<gs, bs>>>(flux, n_reg, c.GL, c.GS, K, c.share, n_items); });
        double n_at = (double)nthreads * K;
        printf("{\"bench\": \"red_f64\", \"GL\": %d, \"GS\": %d, \"share\": %d, \"atomics\": %.0f, \"ms\": %.4f, \"Gatom_per_s\": %.2f, \"ms_plain_store\": %.4f}\n",
               c.GL, c.GS, c.share, n_at, ms, n_at / ms * 1e-6, ms_st);
    }
    for (int big = 0; big < 2; big++) {  // more threads: 8x items
        int n_items = 26448 * 2 * (big ? 8 : 1), K = 344;
        for (int share : {1, 4, 8}) {
            int bs = 256; int gs = (n_items + bs - 1) / bs;
            float ms = timeit([&] { k_atomic_match<<<gs, bs>>>(flux, n_reg, K, share, n_items); });
            float ms2 = timeit([&] { k_atomic<true><<<gs, bs>>>(flux, n_reg, 1, 1, K, share, n_items); });
            printf("{\"bench\": \"red_f64_match\", \"items\": %d, \"share\": %d, \"ms_match\": %.4f, \"ms_plain_red\": %.4f, \"Gupd_per_s_match\": %.2f, \"Gupd_per_s_red\": %.2f}\n",
                   n_items, share, ms, ms2, (double)n_items * K / ms * 1e-6, (double)n_items * K / ms2 * 1e-6);
        }
    }
    // (2) FMA rate
    {
        double* out; CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
        int iters = 4096;
        for (int bs : {256, 512, 1024}) {
            int gs = 148 * (2048 / bs);
            float ms = timeit([&] { k_fma<<<gs, bs>>>(out, iters); });
            double fmas = (double)gs * bs * iters * 8;
            printf("{\"bench\": \"dfma\", \"block\": %d, \"grid\": %d, \"ms\": %.4f, \"Tfma_per_s\": %.3f, \"fma_per_clk_per_sm_at_max_clock\": %.2f}\n",
                   bs, gs, ms, fmas / ms * 1e-9, fmas / (ms * 1e-3) / 148 / (p.clockRate * 1e3));
        }
        long long* cyc; CK(cudaMalloc(&cyc, 16)); long long h[2];
        k_lat<<<1, 32>>>(out, 4096, cyc); CK(cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost));
        printf("{\"bench\": \"dp_latency\", \"dfma_chain_cyc\": %.2f, \"sub_mul_sub_chain_cyc\": %.2f}\n", h[0] / 4096.0, h[1] / 4096.0);
    }
    // (3) exp table
    {
        const int n_tab = 10000; std::vector<double> h(n_tab + 2);
        for (int i = 0; i <= n_tab; i++) h[i] = exp(-10.0 + i * 1e-3); h[n_tab + 1] = h[n_tab];
        double* tab; CK(cudaMalloc(&tab, (n_tab + 2) * 8)); CK(cudaMemcpy(tab, h.data(), (n_tab + 2) * 8, cudaMemcpyHostToDevice));
        double* out; CK(cudaMalloc(&out, 148 * 2048 * sizeof(double)));
        int iters = 2048;
        CK(cudaFuncSetAttribute(k_exptab<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CK(cudaFuncSetAttribute(k_exptab<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int bs : {512, 1024}) {
            int gs0 = 148 * 2;   // 80 KB table -> 2 CTAs/SM
            float ms0 = timeit([&] { k_exptab<0><<<gs0, bs, (n_tab + 2) * 8>>>(tab, out, iters, n_tab); });
            int gs1 = 148;       // 160 KB table -> 1 CTA/SM
            float ms1 = timeit([&] { k_exptab<1><<<gs1, bs, (n_tab + 1) * 16>>>(tab, out, iters, n_tab); });
            printf("{\"bench\": \"exptab\", \"block\": %d, \"Glookup_per_s_2x8B\": %.2f, \"Glookup_per_s_16B\": %.2f}\n",
                   bs, (double)gs0 * bs * iters / ms0 * 1e-6, (double)gs1 * bs * iters / ms1 * 1e-6);
        }
    }
    return 0;
}
