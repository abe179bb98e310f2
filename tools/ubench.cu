// Pipe-throughput microbenchmarks for the instructions the JFA flood pass is built from (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
// Prints warp-instructions per clock per SM for each mix (148 SMs, clock measured with clock64 inside the kernel).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

constexpr int ITER = 2048;
constexpr int CH = 8;   // independent chains per thread

enum Op { FADD, FFMA, FADD2, FFMA2, IMAD, LEA, IMNMX, IMNMX3, FMNMX, FMNMX3, SETPSEL, MIX_A, MIX_B, MIX_C, LDS64, LDS32, LOP3, IADD3, PRMT, FMUL2 };

__device__ __forceinline__ uint32_t lea4(uint32_t a, uint32_t b) { return (a << 4) + b; }

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, const uint32_t* in, long long* cyc) {
    __shared__ float2 sm[2048];
    uint32_t u[CH];
    float f[CH];
    float2 p[CH];
    const uint32_t a0 = in[0], a1 = in[1], a2 = in[2];
    const float fa = __uint_as_float(in[3]), fb = __uint_as_float(in[4]);
    for (int i = 0; i < CH; ++i) {
        u[i] = in[5 + i] + threadIdx.x;
        f[i] = __uint_as_float(in[16 + i]) + threadIdx.x;
        p[i] = make_float2(f[i], f[i] + 1.0f);
    }
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = make_float2((float)i, 1.0f);
    __syncthreads();
    const float2 pa = make_float2(fa, fa), pb = make_float2(fb, fb);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == FADD) f[i] = __fadd_rn(f[i], fa);
            if (OP == FFMA) f[i] = __fmaf_rn(f[i], fa, fb);
            if (OP == FADD2) p[i] = __fadd2_rn(p[i], pa);
            if (OP == FMUL2) p[i] = __fmul2_rn(p[i], pa);
            if (OP == FFMA2) p[i] = __ffma2_rn(p[i], pa, pb);
            if (OP == IMAD) u[i] = u[i] * a0 + a1;
            if (OP == LEA) u[i] = lea4(u[i], a1);
            if (OP == IMNMX) u[i] = min(u[i] ^ a0, a1);   // LOP3 + VIMNMX (2 instr)
            if (OP == IMNMX3) u[i] = __vimin3_u32(u[i], u[(i + 1) % CH] , a2) ;
            if (OP == FMNMX) f[i] = fminf(f[i], f[(i + 1) % CH]);
            if (OP == FMNMX3) f[i] = fminf(fminf(f[i], f[(i + 1) % CH]), f[(i + 2) % CH]);
            if (OP == SETPSEL) { if (f[i] < fa) u[i] = it; f[i] = fminf(f[i], fa + it); }
            if (OP == LOP3) u[i] = (u[i] & a0) ^ a1;
            if (OP == IADD3) u[i] = u[i] + a0 + a1;
            if (OP == PRMT) u[i] = __byte_perm(u[i], a0, a1);
            if (OP == MIX_A) {   // the proposed candidate step for 2 voxels: FADD2 + 2 key (IMAD) + 2x(1/2) VIMNMX3
                p[i] = __fadd2_rn(p[i], pa);
                const uint32_t k0 = __float_as_uint(p[i].x) * 16u + a1, k1 = __float_as_uint(p[i].y) * 16u + a2;
                u[i] = __vimin3_u32(u[i], k0, k1);
            }
            if (OP == MIX_B) {   // same with LEA keys
                p[i] = __fadd2_rn(p[i], pa);
                const uint32_t k0 = lea4(__float_as_uint(p[i].x), a1), k1 = lea4(__float_as_uint(p[i].y), a2);
                u[i] = __vimin3_u32(u[i], k0, k1);
            }
            if (OP == MIX_C) {   // one IMAD key + one LEA key
                p[i] = __fadd2_rn(p[i], pa);
                const uint32_t k0 = __float_as_uint(p[i].x) * 16u + a1, k1 = lea4(__float_as_uint(p[i].y), a2);
                u[i] = __vimin3_u32(u[i], k0, k1);
            }
            if (OP == LDS64) { const float2 v = sm[(u[i] + threadIdx.x) & 2047]; u[i] = __float_as_uint(v.x); p[i].y += v.y; }
            if (OP == LDS32) { const float v = reinterpret_cast<float*>(sm)[(u[i] + threadIdx.x) & 4095]; u[i] = __float_as_uint(v); }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < CH; ++i) acc ^= u[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(p[i].x) ^ __float_as_uint(p[i].y);
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double instr_per_step, int ctas_per_sm) {
    const int grid = 148 * ctas_per_sm;
    uint32_t* out; uint32_t* in; long long* cyc;
    cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&in, 64 * 4); cudaMalloc(&cyc, grid * 8);
    uint32_t h[64];
    for (int i = 0; i < 64; ++i) h[i] = 0x3f800000u + i * 7919u;
    h[0] = 3; h[1] = 12345; h[2] = 777; h[5] = 0; h[6] = 1;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    k<OP><<<grid, 256>>>(out, in, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(out, in, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long* hc = new long long[grid];
    cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += hc[i]; mean /= grid;
    const double warp_instr_per_sm = (double)ITER * CH * instr_per_step * 8 * ctas_per_sm;   // 8 warps per CTA
    printf("%-10s ctas/sm %d: %.3f ms, %.0f cyc/CTA, %.3f warp-instr/clk/SM (%.2f per SMSP), cudaErr=%d\n", name, ctas_per_sm, ms,
           mean, warp_instr_per_sm / mean, warp_instr_per_sm / mean / 4, (int)cudaGetLastError());
    delete[] hc; cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
    for (int c : {2, 4, 8}) {
        printf("---- %d CTAs (x8 warps) per SM\n", c);
        run<FADD>("FADD", 1, c);
        run<FFMA>("FFMA", 1, c);
        run<FADD2>("FADD2", 1, c);
        run<FMUL2>("FMUL2", 1, c);
        run<FFMA2>("FFMA2", 1, c);
        run<IMAD>("IMAD", 1, c);
        run<LEA>("LEA", 1, c);
        run<LOP3>("LOP3", 1, c);
        run<IADD3>("IADD3", 1, c);
        run<PRMT>("PRMT", 1, c);
        run<IMNMX>("LOP+IMNMX", 2, c);
        run<IMNMX3>("VIMNMX3", 1, c);
        run<FMNMX>("FMNMX", 1, c);
        run<FMNMX3>("FMNMX3", 1, c);
        run<SETPSEL>("SETP+SEL+MNMX+IADD", 4, c);
        run<MIX_A>("MIX_A(4)", 4, c);
        run<MIX_B>("MIX_B(4)", 4, c);
        run<MIX_C>("MIX_C(4)", 4, c);
        run<LDS64>("LDS64", 1, c);
        run<LDS32>("LDS32", 1, c);
    }
    return 0;
}
