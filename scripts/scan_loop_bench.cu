// Microbenchmark of the KNN scan's inner loop shapes (development aid, not product code).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o scan_loop_bench scan_loop_bench.cu
// Each variant runs the filter loop (d=3: 3 FMA per query-candidate pair, min over a candidate
// group, one compare per (query, group)) over a shared-memory tile, with no slow path, and
// reports pairs/s and the FMA rate (3 FMA = 6 flop per pair).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ float min3f(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2f(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2f(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2f(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

constexpr int kTile = 1536;

// KQ queries per lane, G candidates per compare group, PACKED: FFMA2, MASK: per-candidate FMUL+select
// (the product kernel's tail masking + slack scaling), PF: software prefetch of the next group.
template <int KQ, int G, bool PACKED, bool MASK, bool PF, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) loop_kernel(const float4 *__restrict__ q, const float4 *__restrict__ cand,
                                                             int reps, int cnt, unsigned *__restrict__ hits) {
    __shared__ __align__(16) float4 tile[kTile];
    for (int i = threadIdx.x; i < kTile; i += THREADS) tile[i] = cand[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = THREADS / 32;
    float a0[KQ], a1[KQ], a2[KQ], th[KQ];
    unsigned long long p0[KQ / 2], p1[KQ / 2], p2[KQ / 2];
#pragma unroll
    for (int i = 0; i < KQ; ++i) {
        const float4 v = q[i * 32 + lane];
        a0[i] = v.x; a1[i] = v.y; a2[i] = v.z; th[i] = v.w;
    }
#pragma unroll
    for (int m = 0; m < KQ / 2; ++m) {
        p0[m] = pack2f(a0[2 * m], a0[2 * m + 1]);
        p1[m] = pack2f(a1[2 * m], a1[2 * m + 1]);
        p2[m] = pack2f(a2[2 * m], a2[2 * m + 1]);
    }
    unsigned nh = 0;
    for (int r = 0; r < reps; ++r) {
        float4 nx[G];
        if (PF) {
#pragma unroll
            for (int j = 0; j < G; ++j) nx[j] = tile[warp + j * W];
        }
        for (int c0 = warp; c0 < kTile; c0 += W * G) {
            float x[G], y[G], z[G], np[G];
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const int c = c0 + j * W;
                float4 t;
                if (PF) t = nx[j]; else t = tile[c];
                x[j] = t.x; y[j] = t.y; z[j] = t.z;
                if (MASK) np[j] = (c < cnt) ? __fmul_rn(t.w, 0.99999618530273437f) : __builtin_huge_valf();
                else np[j] = t.w;
            }
            if (PF) {
                const int cn = (c0 + W * G < kTile) ? c0 + W * G : warp;
#pragma unroll
                for (int j = 0; j < G; ++j) nx[j] = tile[cn + j * W];
            }
            bool any = false;
            if (PACKED) {
#pragma unroll
                for (int m = 0; m < KQ / 2; ++m) {
                    float lo[G], hi[G];
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        unsigned long long acc = pack2f(np[j], np[j]);
                        acc = fma2f(p2[m], pack2f(z[j], z[j]), acc);
                        acc = fma2f(p1[m], pack2f(y[j], y[j]), acc);
                        acc = fma2f(p0[m], pack2f(x[j], x[j]), acc);
                        unpack2f(acc, lo[j], hi[j]);
                    }
                    float ml = lo[0], mh = hi[0];
                    if (G == 3) { ml = min3f(lo[0], lo[1], lo[2]); mh = min3f(hi[0], hi[1], hi[2]); }
                    else if (G == 6) { ml = fminf(min3f(lo[0], lo[1], lo[2]), min3f(lo[3], lo[4], lo[5]));
                                       mh = fminf(min3f(hi[0], hi[1], hi[2]), min3f(hi[3], hi[4], hi[5])); }
                    else {
#pragma unroll
                        for (int j = 1; j < G; ++j) { ml = fminf(ml, lo[j]); mh = fminf(mh, hi[j]); }
                    }
                    any |= (ml <= th[2 * m]);
                    any |= (mh <= th[2 * m + 1]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < KQ; ++i) {
                    float f[G];
#pragma unroll
                    for (int j = 0; j < G; ++j) f[j] = fmaf(a0[i], x[j], fmaf(a1[i], y[j], fmaf(a2[i], z[j], np[j])));
                    float mn = f[0];
                    if (G == 3) mn = min3f(f[0], f[1], f[2]);
                    else {
#pragma unroll
                        for (int j = 1; j < G; ++j) mn = fminf(mn, f[j]);
                    }
                    any |= (mn <= th[i]);
                }
            }
            if (any) ++nh;
        }
    }
    if (nh) atomicAdd(hits, nh);
}


// ---- flipped roles: lanes hold candidates, the 256 queries are warp-uniform and come from the constant bank
// (FFMA2 takes the packed coefficient pair as a UNIFORM-register operand: no register-file read for it)
__constant__ float2 cq0[128], cq1[128], cq2[128];
template <int C, bool TH_SMEM, int THREADS, int MINB, int U>
__global__ void __launch_bounds__(THREADS, MINB) flip_kernel(const float4 *__restrict__ cand, const float *__restrict__ theta,
                                                             int reps, unsigned *__restrict__ hits) {
    __shared__ __align__(16) float4 tile[kTile];
    __shared__ __align__(16) float th[256];
    for (int i = threadIdx.x; i < kTile; i += THREADS) tile[i] = cand[i];
    for (int i = threadIdx.x; i < 256; i += THREADS) th[i] = theta[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = THREADS / 32;
    unsigned nh = 0;
    for (int r = 0; r < reps; ++r) {
        // warp w takes candidate blocks of 32*C: block b -> candidates b*32*C + j*32 + lane
        for (int b = warp; b * 32 * C < kTile; b += W) {
            float x[C], y[C], z[C], np[C];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const float4 t = tile[b * 32 * C + j * 32 + lane];
                x[j] = t.x; y[j] = t.y; z[j] = t.z; np[j] = __fmul_rn(t.w, 0.99999618530273437f);
            }
            bool any = false;
#pragma unroll U
            for (int m = 0; m < 128; ++m) {
                const unsigned long long a0 = *reinterpret_cast<const unsigned long long *>(&cq0[m]);
                const unsigned long long a1 = *reinterpret_cast<const unsigned long long *>(&cq1[m]);
                const unsigned long long a2 = *reinterpret_cast<const unsigned long long *>(&cq2[m]);
                float lo[C], hi[C];
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    unsigned long long f = fma2f(a2, pack2f(z[j], z[j]), pack2f(np[j], np[j]));
                    f = fma2f(a1, pack2f(y[j], y[j]), f);
                    f = fma2f(a0, pack2f(x[j], x[j]), f);
                    unpack2f(f, lo[j], hi[j]);
                }
                float ml, mh;
                if (C == 3) { ml = min3f(lo[0], lo[1], lo[2]); mh = min3f(hi[0], hi[1], hi[2]); }
                else if (C == 6) { ml = fminf(min3f(lo[0], lo[1], lo[2]), min3f(lo[3 % C], lo[4 % C], lo[5 % C]));
                                   mh = fminf(min3f(hi[0], hi[1], hi[2]), min3f(hi[3 % C], hi[4 % C], hi[5 % C])); }
                else { ml = lo[0]; mh = hi[0];
#pragma unroll
                    for (int j = 1; j < C; ++j) { ml = fminf(ml, lo[j]); mh = fminf(mh, hi[j]); } }
                if (TH_SMEM) {
                    const float2 t2 = *reinterpret_cast<const float2 *>(&th[2 * m]);
                    any |= (ml <= t2.x); any |= (mh <= t2.y);
                } else {
                    any |= (ml <= -100.f); any |= (mh <= -100.f);
                }
            }
            if (any) ++nh;
        }
    }
    if (nh) atomicAdd(hits, nh);
}

template <int C, bool TH_SMEM, int THREADS, int MINB, int U>
void run_flip(const char *name, const float4 *cand, const float *theta, unsigned *hits, int sms) {
    const int blocks = sms * MINB;
    const int reps = 64;
    auto k = flip_kernel<C, TH_SMEM, THREADS, MINB, U>;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k);
    k<<<blocks, THREADS>>>(cand, theta, 2, hits);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(a);
        k<<<blocks, THREADS>>>(cand, theta, reps, hits);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double pairs = (double)blocks * reps * kTile * 256.0;
    printf("%-44s regs %3d occ %d/SM(%2d warps)  %.3f ms  %.3e pairs/s  %.1f TFLOP/s(FMA)  c3-scan-equiv %.1f us  err=%s\n", name,
           fa.numRegs, occ, occ * THREADS / 32, best, pairs / (best * 1e-3), pairs * 6.0 / (best * 1e-3) / 1e12,
           1.024e9 * 0.99996 / (pairs / (best * 1e-3)) * 1e6, cudaGetErrorString(cudaGetLastError()));
}

template <int KQ, int G, bool PACKED, bool MASK, bool PF, int THREADS, int MINB>
void run(const char *name, const float4 *q, const float4 *cand, unsigned *hits, int sms) {
    const int blocks = sms * MINB;
    const int reps = 64;
    auto k = loop_kernel<KQ, G, PACKED, MASK, PF, THREADS, MINB>;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k);
    k<<<blocks, THREADS>>>(q, cand, 2, kTile, hits);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int it = 0; it < 5; ++it) {
        cudaEventRecord(a);
        k<<<blocks, THREADS>>>(q, cand, reps, kTile, hits);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    // candidates visited per warp per rep: kTile / W (each a full 32*KQ query block)
    const double W = THREADS / 32;
    const double pairs = (double)blocks * W * reps * (kTile / W) * 32.0 * KQ;
    const double tfl = pairs * 6.0 / (best * 1e-3) / 1e12;
    printf("%-44s regs %3d occ %d/SM(%2d warps)  %.3f ms  %.3e pairs/s  %.1f TFLOP/s(FMA)  c3-scan-equiv %.1f us  err=%s\n", name,
           fa.numRegs, occ, occ * THREADS / 32, best, pairs / (best * 1e-3), tfl, 1.024e9 * 0.99996 / (pairs / (best * 1e-3)) * 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float4 *q, *cand; unsigned *hits;
    cudaMalloc(&q, 256 * sizeof(float4));
    cudaMalloc(&cand, kTile * sizeof(float4));
    cudaMalloc(&hits, 4);
    cudaMemset(hits, 0, 4);
    float4 hq[256], *hc = (float4 *)malloc(kTile * sizeof(float4));
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    for (int i = 0; i < 256; ++i) hq[i] = make_float4(-2 * rnd(), -2 * rnd(), -2 * rnd(), -100.f);
    for (int i = 0; i < kTile; ++i) { float x = rnd(), y = rnd(), z = rnd(); hc[i] = make_float4(x, y, z, x * x + y * y + z * z); }
    cudaMemcpy(q, hq, sizeof(hq), cudaMemcpyHostToDevice);
    cudaMemcpy(cand, hc, kTile * sizeof(float4), cudaMemcpyHostToDevice);

    {
        float2 h0[128], h1[128], h2[128]; float hth[256];
        for (int m = 0; m < 128; ++m) {
            h0[m] = make_float2(hq[2 * m].x, hq[2 * m + 1].x); h1[m] = make_float2(hq[2 * m].y, hq[2 * m + 1].y);
            h2[m] = make_float2(hq[2 * m].z, hq[2 * m + 1].z);
        }
        for (int i = 0; i < 256; ++i) hth[i] = -100.f;
        cudaMemcpyToSymbol(cq0, h0, sizeof(h0)); cudaMemcpyToSymbol(cq1, h1, sizeof(h1)); cudaMemcpyToSymbol(cq2, h2, sizeof(h2));
        float *theta; cudaMalloc(&theta, 1024); cudaMemcpy(theta, hth, 1024, cudaMemcpyHostToDevice);
        run_flip<3, true, 256, 2, 128>("flip C3 th-smem 256x2 U128", cand, theta, hits, sms);
        run_flip<3, true, 256, 2, 16>("flip C3 th-smem 256x2 U16", cand, theta, hits, sms);
        run_flip<3, true, 256, 4, 16>("flip C3 th-smem 256x4 U16", cand, theta, hits, sms);
        run_flip<3, true, 256, 4, 8>("flip C3 th-smem 256x4 U8", cand, theta, hits, sms);
        run_flip<3, false, 256, 4, 8>("flip C3 th-imm 256x4 U8", cand, theta, hits, sms);
        run_flip<6, true, 256, 2, 128>("flip C6 th-smem 256x2 U128", cand, theta, hits, sms);
        run_flip<6, true, 256, 2, 16>("flip C6 th-smem 256x2 U16", cand, theta, hits, sms);
        run_flip<6, true, 256, 2, 8>("flip C6 th-smem 256x2 U8", cand, theta, hits, sms);
        run_flip<6, true, 256, 3, 8>("flip C6 th-smem 256x3 U8", cand, theta, hits, sms);
        run_flip<6, true, 256, 4, 8>("flip C6 th-smem 256x4 U8", cand, theta, hits, sms);
        run_flip<6, true, 256, 4, 4>("flip C6 th-smem 256x4 U4", cand, theta, hits, sms);
        run_flip<6, false, 256, 4, 8>("flip C6 th-imm 256x4 U8", cand, theta, hits, sms);
        run_flip<4, true, 256, 4, 8>("flip C4 th-smem 256x4 U8", cand, theta, hits, sms);
        run_flip<2, true, 256, 4, 8>("flip C2 th-smem 256x4 U8", cand, theta, hits, sms);
        run_flip<12, true, 128, 4, 4>("flip C12 th-smem 128x4 U4", cand, theta, hits, sms);
        run_flip<12, true, 128, 4, 8>("flip C12 th-smem 128x4 U8", cand, theta, hits, sms);
    }
    //   KQ G  PACKED MASK  PF    THREADS MINB
    run<8, 3, true, true, false, 256, 2>("kq8 g3 ffma2 mask (current) 256x2", q, cand, hits, sms);
    run<8, 3, false, true, false, 256, 2>("kq8 g3 ffma  mask 256x2", q, cand, hits, sms);
    run<8, 3, true, false, false, 256, 2>("kq8 g3 ffma2 nomask 256x2", q, cand, hits, sms);
    run<8, 3, true, false, true, 256, 2>("kq8 g3 ffma2 nomask prefetch 256x2", q, cand, hits, sms);
    run<8, 3, false, false, false, 256, 2>("kq8 g3 ffma  nomask 256x2", q, cand, hits, sms);
    run<8, 3, false, false, true, 256, 2>("kq8 g3 ffma  nomask prefetch 256x2", q, cand, hits, sms);
    run<8, 6, true, false, false, 256, 2>("kq8 g6 ffma2 nomask 256x2", q, cand, hits, sms);
    run<8, 2, true, false, false, 256, 2>("kq8 g2 ffma2 nomask 256x2", q, cand, hits, sms);
    run<8, 3, true, false, false, 256, 3>("kq8 g3 ffma2 nomask 256x3", q, cand, hits, sms);
    run<8, 3, true, false, false, 512, 1>("kq8 g3 ffma2 nomask 512x1", q, cand, hits, sms);
    run<4, 3, true, false, false, 256, 4>("kq4 g3 ffma2 nomask 256x4", q, cand, hits, sms);
    run<4, 3, true, false, true, 256, 4>("kq4 g3 ffma2 nomask prefetch 256x4", q, cand, hits, sms);
    run<4, 6, true, false, false, 256, 4>("kq4 g6 ffma2 nomask 256x4", q, cand, hits, sms);
    run<4, 6, true, false, true, 256, 4>("kq4 g6 ffma2 nomask prefetch 256x4", q, cand, hits, sms);
    run<4, 6, false, false, false, 256, 4>("kq4 g6 ffma  nomask 256x4", q, cand, hits, sms);
    run<4, 3, false, false, false, 256, 4>("kq4 g3 ffma  nomask 256x4", q, cand, hits, sms);
    run<4, 6, true, false, false, 256, 6>("kq4 g6 ffma2 nomask 256x6", q, cand, hits, sms);
    run<2, 6, true, false, false, 256, 8>("kq2 g6 ffma2 nomask 256x8", q, cand, hits, sms);
    run<2, 12, true, false, false, 256, 8>("kq2 g12 ffma2 nomask 256x8", q, cand, hits, sms);
    run<8, 3, true, false, false, 128, 4>("kq8 g3 ffma2 nomask 128x4", q, cand, hits, sms);
    unsigned h; cudaMemcpy(&h, hits, 4, cudaMemcpyDeviceToHost);
    printf("hits %u (expect 0)\n", h);
    return 0;
}
