// Micro-benchmark behind the persistent-loop design decisions: what does a grid-wide exchange cost on a B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_latency barrier_latency.cu && ./barrier_latency [nctas]
// N CTAs (one per SM, 256 threads, cooperative launch).  Per iteration: every CTA writes a 16 KB slab, the grid synchronises, every
// CTA reads data of another CTA.  Variants of the barrier (single atomic counter vs per-CTA flags) and of the read (fresh remote data vs
// never-written data, 1 vs 8 loads in flight).  Prints median cycles per phase over the iterations (CTA 0, thread 0's clock).
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int PT = 256;
constexpr int SLAB = 4096;      // floats per CTA slab (16 KB)
constexpr int ITERS = 400;

__device__ __forceinline__ void barrier_counter(unsigned* counter, unsigned& target, unsigned n) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += n;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        for (;;) {
            unsigned v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
        }
        asm volatile("fence.acquire.gpu;" ::: "memory");
    }
    __syncthreads();
}
// every CTA publishes its epoch in its own word; thread t waits for CTA t's word: no atomic, all polls in parallel
__device__ __forceinline__ void barrier_flags(unsigned* flags, unsigned epoch, unsigned n) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    for (unsigned t = threadIdx.x; t < n; t += PT) {
        for (;;) {
            unsigned v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + t) : "memory");
            if (v >= epoch) break;
        }
    }
    asm volatile("fence.acquire.gpu;" ::: "memory");
    __syncthreads();
}
// flags spread over separate 128-byte lines
__device__ __forceinline__ void barrier_flags_lines(unsigned* flags, unsigned epoch, unsigned n) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 32 * blockIdx.x), "r"(epoch) : "memory");
    for (unsigned t = threadIdx.x; t < n; t += PT) {
        for (;;) {
            unsigned v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32 * t) : "memory");
            if (v >= epoch) break;
        }
    }
    asm volatile("fence.acquire.gpu;" ::: "memory");
    __syncthreads();
}

struct Args { float* slabs; const float* constant; unsigned* counter; unsigned* flags; long long* out; int mode; };

__global__ void __launch_bounds__(PT, 1) bench_kernel(Args a) {
    const int cta = blockIdx.x, n = gridDim.x, tid = threadIdx.x;
    unsigned target = 0, epoch = 0;
    float sink = 0.f;
    __shared__ float sh[PT];
    for (int it = 0; it < ITERS; ++it) {
        // write the slab (16 floats per thread, coalesced)
        float* mine = a.slabs + (size_t)cta * SLAB;
#pragma unroll
        for (int j = 0; j < SLAB / PT; ++j) mine[j * PT + tid] = (float)(it + j) + sink * 1e-30f;
        const long long t0 = clock64();
        ++epoch;
        if (a.mode & 1) barrier_flags(a.flags, epoch, n);
        else if (a.mode & 2) barrier_flags_lines(a.flags, epoch, n);
        else barrier_counter(a.counter, target, n);
        const long long t1 = clock64();
        // read: 1 load per thread of the neighbour's fresh slab
        const float* theirs = a.slabs + (size_t)((cta + 1 + it % (n - 1)) % n) * SLAB;
        float v = __ldcg(theirs + tid);
        sh[tid] = v; __syncthreads();
        const long long t2 = clock64();
        // read: 8 independent loads per thread of another fresh slab
        const float* theirs2 = a.slabs + (size_t)((cta + 2 + it % (n - 2)) % n) * SLAB;
        float w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = __ldcg(theirs2 + j * PT + tid);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += w[j];
        sh[tid] = s; __syncthreads();
        const long long t3 = clock64();
        // read: 1 load per thread of data nobody writes (L2-resident after the first iterations)
        float c = __ldcg(a.constant + (size_t)cta * SLAB + tid);
        sh[tid] = c; __syncthreads();
        const long long t4 = clock64();
        // dependent chain of 4 loads of constant data (pointer-free: index depends on the previous value being 0)
        int idx = tid;
        float d = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { d = __ldcg(a.constant + (size_t)cta * SLAB + idx); idx = tid + (int)(d * 0.f) + (j + 1) * PT; }
        sh[tid] = d; __syncthreads();
        const long long t5 = clock64();
        sink += sh[(tid + 1) % PT];
        if (cta == 0 && tid == 0) {
            long long* o = a.out + (size_t)it * 8;
            o[0] = t1 - t0; o[1] = t2 - t1; o[2] = t3 - t2; o[3] = t4 - t3; o[4] = t5 - t4;
        }
        // second barrier so that nobody overwrites a slab that is still being read
        ++epoch;
        if (a.mode & 1) barrier_flags(a.flags, epoch, n);
        else if (a.mode & 2) barrier_flags_lines(a.flags, epoch, n);
        else barrier_counter(a.counter, target, n);
    }
    if (sink == 12345.f) a.out[0] = 0;
}

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 144;
    float *slabs, *constant; unsigned *counter, *flags; long long* out;
    cudaMalloc(&slabs, (size_t)148 * SLAB * 4); cudaMalloc(&constant, (size_t)148 * SLAB * 4);
    cudaMemset(slabs, 0, (size_t)148 * SLAB * 4); cudaMemset(constant, 0, (size_t)148 * SLAB * 4);
    cudaMalloc(&counter, 256); cudaMalloc(&flags, 148 * 128); cudaMalloc(&out, (size_t)ITERS * 8 * 8);
    const char* names[3] = {"single atomic counter (red.release + poll)", "per-CTA flags, packed words", "per-CTA flags, one 128-B line each"};
    for (int mode = 0; mode < 3; ++mode) {
        cudaMemset(counter, 0, 256); cudaMemset(flags, 0, 148 * 128);
        Args a{slabs, constant, counter, flags, out, mode};
        void* params[] = {&a};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)bench_kernel, dim3(n), dim3(PT), params, 0, 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        std::vector<long long> h((size_t)ITERS * 8);
        cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
        printf("%d CTAs, barrier = %s\n", n, names[mode]);
        const char* ph[5] = {"grid barrier", "1 load/thread, fresh remote slab", "8 loads/thread, fresh remote slab", "1 load/thread, constant data",
                             "4 DEPENDENT loads, constant data"};
        for (int k = 0; k < 5; ++k) {
            std::vector<long long> v;
            for (int it = 20; it < ITERS; ++it) v.push_back(h[(size_t)it * 8 + k]);
            std::sort(v.begin(), v.end());
            printf("   %-36s median %6lld   p10 %6lld   p90 %6lld cycles\n", ph[k], v[v.size() / 2], v[v.size() / 10], v[v.size() * 9 / 10]);
        }
    }
    return 0;
}
