// Micro-benchmark: variants of the 2D upwind FV sweep on a uniform level-L layout (rows of N+2 doubles, ghost ring),
// to find what bounds the production kernel.  nvcc -O3 -arch=sm_100a -fmad=false tools/fv_micro.cu -o /tmp/fv_micro
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct P { double ha, haa, hb, hbb, dt, inv; };

__device__ __forceinline__ double fl(double ha, double haa, double l, double r) { return ha * (l + r) + haa * (l - r); }

__device__ __forceinline__ double upd(const P& p, double c, double xm, double xp, double ym, double yp)
{
    double acc = -fl(p.ha, p.haa, xm, c) + fl(p.ha, p.haa, c, xp);
    acc        = (acc + -fl(p.hb, p.hbb, ym, c)) + fl(p.hb, p.hbb, c, yp);
    return c - p.dt * (acc * p.inv);
}

// A: thread per cell, CPT cells per thread strided by blockDim (the production mapping), 8-byte accesses
template <int CPT>
__global__ void __launch_bounds__(256) fvA(const double* __restrict__ u, double* __restrict__ out, int N, long stride, P p)
{
    const long cells = (long)N * N;
    const long base  = (long)blockIdx.x * 256 * CPT;
    double c[CPT], xm[CPT], xp[CPT], ym[CPT], yp[CPT];
    long off[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k)
    {
        long g = base + threadIdx.x + k * 256;
        if (g >= cells) g = cells - 1;
        const long row = g / N, col = g - row * N;
        off[k]         = (row + 1) * stride + col + 1;
        c[k]           = u[off[k]];
        xm[k]          = u[off[k] - 1];
        xp[k]          = u[off[k] + 1];
        ym[k]          = u[off[k] - stride];
        yp[k]          = u[off[k] + stride];
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k)
    {
        if (base + threadIdx.x + k * 256 < cells) out[off[k]] = upd(p, c[k], xm[k], xp[k], ym[k], yp[k]);
    }
}

// B: neighbours i-1 / i+1 through warp shuffles (edge lanes load), 8-byte accesses
template <int CPT>
__global__ void __launch_bounds__(256) fvB(const double* __restrict__ u, double* __restrict__ out, int N, long stride, P p)
{
    const long base = (long)blockIdx.x * 256 * CPT;
    const int lane  = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < CPT; ++k)
    {
        const long g   = base + threadIdx.x + k * 256;
        const long row = g / N, col = g - row * N;
        const long o   = (row + 1) * stride + col + 1;
        const double c = u[o];
        double xm      = __shfl_up_sync(0xffffffffu, c, 1);
        double xp      = __shfl_down_sync(0xffffffffu, c, 1);
        if (lane == 0) xm = u[o - 1];
        if (lane == 31) xp = u[o + 1];
        out[o] = upd(p, c, xm, xp, u[o - stride], u[o + stride]);
    }
}

// C: shared-memory tile: CTA = 256 columns x R rows; (R+2) rows staged once, 8-byte coalesced loads
template <int R>
__global__ void __launch_bounds__(256) fvC(const double* __restrict__ u, double* __restrict__ out, int N, long stride, P p)
{
    __shared__ double t[R + 2][256 + 2];
    const int tilesx = N / 256;
    const int tx = blockIdx.x % tilesx, ty = blockIdx.x / tilesx;
    const long col0 = (long)tx * 256, row0 = (long)ty * R;
    for (int r = 0; r < R + 2; ++r)
    {
        const long o = (row0 + r) * stride + col0 + threadIdx.x;
        t[r][threadIdx.x] = u[o];
        if (threadIdx.x < 2) t[r][256 + threadIdx.x] = u[o + 256];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r)
    {
        const int x = threadIdx.x + 1;
        out[(row0 + r + 1) * stride + col0 + x] = upd(p, t[r + 1][x], t[r + 1][x - 1], t[r + 1][x + 1], t[r][x], t[r + 2][x]);
    }
}

// D: register rolling window down a column strip: each thread owns one column of an R-row strip, reads each value once
template <int R>
__global__ void __launch_bounds__(256) fvD(const double* __restrict__ u, double* __restrict__ out, int N, long stride, P p)
{
    const int tilesx = N / 256;
    const int tx = blockIdx.x % tilesx, ty = blockIdx.x / tilesx;
    const long col = (long)tx * 256 + threadIdx.x + 1;
    const long row0 = (long)ty * R;
    const int lane  = threadIdx.x & 31;
    double v[R + 2];
#pragma unroll
    for (int r = 0; r < R + 2; ++r) v[r] = u[(row0 + r) * stride + col];
#pragma unroll
    for (int r = 1; r <= R; ++r)
    {
        double xm = __shfl_up_sync(0xffffffffu, v[r], 1);
        double xp = __shfl_down_sync(0xffffffffu, v[r], 1);
        const long o = (row0 + r) * stride + col;
        if (lane == 0) xm = u[o - 1];
        if (lane == 31) xp = u[o + 1];
        out[o] = upd(p, v[r], xm, xp, v[r - 1], v[r + 1]);
    }
}

// E: plain copies for reference
__global__ void copy8(const double* __restrict__ a, double* __restrict__ b, long n)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void copy16(const double2* __restrict__ a, double2* __restrict__ b, long n)
{
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F>
float timeit(F&& f, int iters = 10)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main(int argc, char** argv)
{
    const int L = argc > 1 ? atoi(argv[1]) : 13;
    const int N = 1 << L;
    const long stride = N + 2, total = stride * (N + 2);
    double *u, *o;
    CK(cudaMalloc(&u, total * 8));
    CK(cudaMalloc(&o, total * 8));
    CK(cudaMemset(u, 0, total * 8));
    CK(cudaMemset(o, 0, total * 8));
    P p{0.5, 0.5, 0.5, 0.5, 0.5 / N, (double)N};
    const long cells = (long)N * N;
    const double gb  = 16.0 * cells / 1e9;
    auto rep = [&](const char* name, float ms) { printf("%-28s %8.3f ms  %8.1f GB/s (algorithmic 16 B/cell)\n", name, ms, gb / (ms * 1e-3)); };
    rep("copy8  (grid-stride)", timeit([&] { copy8<<<148 * 16, 256>>>(u, o, cells); }));
    rep("copy16 (grid-stride)", timeit([&] { copy16<<<148 * 16, 256>>>((double2*)u, (double2*)o, cells / 2); }));
    rep("A cpt=1 strided", timeit([&] { fvA<1><<<(cells + 255) / 256, 256>>>(u, o, N, stride, p); }));
    rep("A cpt=2 strided", timeit([&] { fvA<2><<<(cells + 511) / 512, 256>>>(u, o, N, stride, p); }));
    rep("A cpt=4 strided", timeit([&] { fvA<4><<<(cells + 1023) / 1024, 256>>>(u, o, N, stride, p); }));
    rep("A cpt=8 strided", timeit([&] { fvA<8><<<(cells + 2047) / 2048, 256>>>(u, o, N, stride, p); }));
    rep("B cpt=4 shuffle", timeit([&] { fvB<4><<<cells / 1024, 256>>>(u, o, N, stride, p); }));
    rep("C smem tile R=4", timeit([&] { fvC<4><<<(N / 256) * (N / 4), 256>>>(u, o, N, stride, p); }));
    rep("C smem tile R=8", timeit([&] { fvC<8><<<(N / 256) * (N / 8), 256>>>(u, o, N, stride, p); }));
    rep("C smem tile R=16", timeit([&] { fvC<16><<<(N / 256) * (N / 16), 256>>>(u, o, N, stride, p); }));
    rep("D column strip R=4", timeit([&] { fvD<4><<<(N / 256) * (N / 4), 256>>>(u, o, N, stride, p); }));
    rep("D column strip R=8", timeit([&] { fvD<8><<<(N / 256) * (N / 8), 256>>>(u, o, N, stride, p); }));
    rep("D column strip R=16", timeit([&] { fvD<16><<<(N / 256) * (N / 16), 256>>>(u, o, N, stride, p); }));
    rep("D column strip R=32", timeit([&] { fvD<32><<<(N / 256) * (N / 32), 256>>>(u, o, N, stride, p); }));
    CK(cudaDeviceSynchronize());
    return 0;
}
