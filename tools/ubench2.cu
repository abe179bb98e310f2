// Candidate-evaluation step variants for the JFA flood pass (sm_100a): cycles per (candidate x 2 voxels) per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o tools/ubench2 tools/ubench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITER = 4096;

__device__ __forceinline__ float2 sq2(float2 x, float2 nz) { return __ffma2_rn(x, x, nz); }

// V = 0: FSETP/SEL/FMNMX (current kernel)      V = 1: packed distance + integer keys + VIMNMX3
// V = 2: scalar distance + keys + VIMNMX3       V = 3: packed ddz/sq, scalar sum
// V = 4: keys only on pre-made distances (FADD2 + keys + min)   V = 5: like 1, both keys forced to IMAD
// V = 6: like 1, both keys forced to shift+add  V = 7: packed distance + FMNMX3 only (no index)
template <int V>
__global__ void __launch_bounds__(256, 2) k(uint32_t* out, const float* in, long long* cyc, float nzr, uint32_t kc, uint32_t km) {
    float2 xy[9], fz[9];
    for (int c = 0; c < 9; ++c) {
        xy[c] = make_float2(in[c] + threadIdx.x, in[c + 9] + threadIdx.x);
        fz[c] = make_float2(in[c + 18], in[c + 27] + threadIdx.x);
    }
    const float2 nz = make_float2(nzr, nzr);
    uint32_t acc = 0;
    float nq = in[40];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        nq += 1.0f;
        const float2 nqz = make_float2(nq, nq);
        if (V == 0) {
            float2 best = make_float2(1e30f, 1e30f);
            int ia = 0, ib = 0;
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                const float2 ddz = __fadd2_rn(fz[c], nqz);
                const float2 d = __fadd2_rn(xy[c], sq2(ddz, nz));
                if (d.x < best.x) ia = c * 4 + 100;
                if (d.y < best.y) ib = c * 4 + 100;
                best.x = fminf(best.x, d.x);
                best.y = fminf(best.y, d.y);
            }
            acc ^= ia ^ (ib << 8) ^ __float_as_uint(best.x) ^ __float_as_uint(best.y);
        } else if (V == 7) {
            float2 best = make_float2(1e30f, 1e30f);
#pragma unroll
            for (int c = 0; c < 9; c += 1) {
                const float2 ddz = __fadd2_rn(fz[c], nqz);
                const float2 d = __fadd2_rn(xy[c], sq2(ddz, nz));
                best.x = fminf(best.x, d.x);
                best.y = fminf(best.y, d.y);
            }
            acc ^= __float_as_uint(best.x) ^ __float_as_uint(best.y);
        } else {
            uint32_t ka[9], kb[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                float2 d;
                if (V == 1 || V == 5 || V == 6 || V == 8 || V == 9) {
                    const float2 ddz = __fadd2_rn(fz[c], nqz);
                    d = __fadd2_rn(xy[c], sq2(ddz, nz));
                } else if (V == 2) {
                    const float dzx = __fadd_rn(fz[c].x, nq), dzy = __fadd_rn(fz[c].y, nq);
                    d.x = __fadd_rn(xy[c].x, __fmul_rn(dzx, dzx));
                    d.y = __fadd_rn(xy[c].y, __fmul_rn(dzy, dzy));
                } else if (V == 3) {
                    const float2 ddz = __fadd2_rn(fz[c], nqz);
                    const float2 s = sq2(ddz, nz);
                    d.x = __fadd_rn(xy[c].x, s.x);
                    d.y = __fadd_rn(xy[c].y, s.y);
                } else {
                    d = __fadd2_rn(xy[c], nqz);
                }
                if (V == 5) {
                    asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(ka[c]) : "r"(__float_as_uint(d.x)), "r"(kc + c));
                    asm("mad.lo.u32 %0, %1, 16, %2;" : "=r"(kb[c]) : "r"(__float_as_uint(d.y)), "r"(kc + c));
                } else if (V == 8) {
                    ka[c] = __float_as_uint(d.x) * km + (kc + c);
                    kb[c] = __float_as_uint(d.y) * 16u + (kc + c);
                } else if (V == 9) {
                    ka[c] = __float_as_uint(d.x) * km + (kc + c);
                    kb[c] = __float_as_uint(d.y) * km + (kc + c);
                } else if (V == 6) {
                    ka[c] = __funnelshift_l(0u, __float_as_uint(d.x), 4) | (uint32_t)c;
                    kb[c] = __funnelshift_l(0u, __float_as_uint(d.y), 4) | (uint32_t)c;
                } else {
                    ka[c] = __float_as_uint(d.x) * 16u + (kc + c);
                    kb[c] = __float_as_uint(d.y) * 16u + (kc + c);
                }
            }
            uint32_t ma = __vimin3_u32(ka[0], ka[1], ka[2]);
            ma = __vimin3_u32(ma, ka[3], ka[4]);
            ma = __vimin3_u32(ma, ka[5], ka[6]);
            ma = __vimin3_u32(ma, ka[7], ka[8]);
            uint32_t mb = __vimin3_u32(kb[0], kb[1], kb[2]);
            mb = __vimin3_u32(mb, kb[3], kb[4]);
            mb = __vimin3_u32(mb, kb[5], kb[6]);
            mb = __vimin3_u32(mb, kb[7], kb[8]);
            acc ^= ma ^ (mb << 1);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char* name) {
    const int grid = 148 * 2;
    uint32_t* out; float* in; long long* cyc;
    cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&in, 64 * 4); cudaMalloc(&cyc, grid * 8);
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 0.37f;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    k<V><<<grid, 256>>>(out, in, cyc, -0.0f, 12345u, 16u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<grid, 256>>>(out, in, cyc, -0.0f, 12345u, 16u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* hc = new long long[grid];
    cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += hc[i]; mean /= grid;
    // 16 warps per SM = 4 per SMSP; each iteration = 9 candidate-pairs per warp
    const double cyc_per_candpair = mean / ((double)ITER * 9 * 4);
    printf("%-34s %.3f ms  %.2f SMSP-cycles per (cand x 2 voxels x warp)  -> 27 cands: %.0f cyc per 64 voxels, %.2f voxels/clk/SM, %.2f ms/pass @1024^3  err=%d\n",
           name, ms, cyc_per_candpair, cyc_per_candpair * 27, 64.0 * 4 / (cyc_per_candpair * 27),
           1073741824.0 / (64.0 * 4 / (cyc_per_candpair * 27) * 148 * 1.965e9) * 1e3, (int)cudaGetLastError());
    delete[] hc; cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
    run<0>("V0 FSETP/SEL/FMNMX (current)");
    run<1>("V1 packed dist + keys + VIMNMX3");
    run<2>("V2 scalar dist + keys + VIMNMX3");
    run<3>("V3 packed ddz/sq, scalar sum");
    run<4>("V4 FADD2 + keys + VIMNMX3");
    run<5>("V5 V1 with IMAD keys");
    run<6>("V6 V1 with SHF/LOP keys");
    run<7>("V7 packed dist + FMNMX only");
    run<8>("V8 V1 with 1 IMAD + 1 LEA key");
    run<9>("V9 V1 with 2 IMAD keys");
    return 0;
}
