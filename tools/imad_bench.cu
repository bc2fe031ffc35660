// Integer-pipe microbenchmark for the roofline denominators that MEASURED_PEAKS.json
// lacks: throughput of IMAD (lo / hi), IMAD.WIDE.U32 with and without carry chains,
// IADD3 alongside, and the library's own Montgomery multiplier.  Prints one JSON line
// per test with ops per clock per SM (clock64-based) and ops per second (event-based).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../halo2-experiments_b200/csrc/field.cuh"

using namespace b200zk;

#define ITERS 4096

template <int MODE> __global__ void __launch_bounds__(256) k_int(uint32_t* out, uint32_t a0, uint32_t b0, long long* cycles) {
    uint32_t x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a0 + threadIdx.x * 8 + i; y[i] = b0 + i; }
    uint32_t a = a0 | 1, b = b0 | 3;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {            // IMAD lo
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            } else if (MODE == 1) {     // IMAD.HI
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            } else if (MODE == 2) {     // IMAD.WIDE.U32 (64-bit accumulate)
                unsigned long long acc = ((unsigned long long)y[i] << 32) | x[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(x[i]));
                x[i] = (uint32_t)acc; y[i] = (uint32_t)(acc >> 32);
            } else if (MODE == 3) {     // lo.cc + hi chain with register operands (what field.cuh emits)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
            } else if (MODE == 4) {     // IADD3 only
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));
            } else if (MODE == 5) {     // IMAD lo + IADD3 interleaved (dual-pipe issue)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class F, int CHAINS> __global__ void __launch_bounds__(256) k_mul(fe_t* out, const fe_t* in, long long* cycles, int iters) {
    fe_t x[CHAINS], m = in[1];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { x[i] = in[0]; x[i].l[0] += threadIdx.x + i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = F::mul(x[i], m);
    }
    long long t1 = clock64();
    fe_t s = x[0];
#pragma unroll
    for (int i = 1; i < CHAINS; ++i) s = F::add(s, x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// FP64 pipe: DFMA alone (MIX = 0) and DFMA interleaved 1:1 with the lo.cc/hi IMAD.WIDE pair (MIX = 1)
template <int MIX> __global__ void __launch_bounds__(256) k_dfma(double* out, double a0, uint32_t b0, long long* cycles) {
    double x[8];
    uint32_t u[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = a0 + threadIdx.x + i; u[i] = b0 + i; v[i] = b0 * 3 + i; }
    double a = a0 * 1.0000001, b = a0 * 0.5;
    uint32_t ia = b0 | 1, ib = b0 | 3;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
            if (MIX) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(u[i]), "+r"(v[i]) : "r"(ia), "r"(ib));
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + (double)(u[i] ^ v[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double median_cycles(long long* d_cycles, int blocks) {
    long long* h = new long long[blocks];
    cudaMemcpy(h, d_cycles, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < blocks; ++i) s += (double)h[i];
    delete[] h; return s / blocks;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint32_t* d_out; long long* d_cycles; fe_t *d_fe_out, *d_fe_in;
    int max_blocks = sms * 8;
    cudaMalloc(&d_out, max_blocks * 256 * 4); cudaMalloc(&d_cycles, max_blocks * 8);
    cudaMalloc(&d_fe_out, (size_t)max_blocks * 256 * 32); cudaMalloc(&d_fe_in, 64);
    uint32_t hin[16] = {5, 7, 11, 13, 17, 19, 23, 0x0fffffffu, 3, 1, 4, 1, 5, 9, 2, 0x0aaaaaaau};
    cudaMemcpy(d_fe_in, hin, 64, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"imad_lo", "imad_hi", "imad_wide", "imad_lo_cc_hi_pair", "iadd_x2", "imad_lo_plus_iadd"};
    const double ops_per_inner[] = {1, 1, 1, 2, 2, 2};
    for (int bps = 1; bps <= 8; bps *= 2) {          // resident blocks per SM (256 threads each)
        int blocks = sms * bps;
        for (int mode = 0; mode < 6; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                switch (mode) {
                    case 0: k_int<0><<<blocks, 256>>>(d_out, 12345, 678, d_cycles); break;
                    case 1: k_int<1><<<blocks, 256>>>(d_out, 12345, 678, d_cycles); break;
                    case 2: k_int<2><<<blocks, 256>>>(d_out, 12345, 678, d_cycles); break;
                    case 3: k_int<3><<<blocks, 256>>>(d_out, 12345, 678, d_cycles); break;
                    case 4: k_int<4><<<blocks, 256>>>(d_out, 12345, 678, d_cycles); break;
                    case 5: k_int<5><<<blocks, 256>>>(d_out, 12345, 678, d_cycles); break;
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double cyc = median_cycles(d_cycles, blocks);
            double ops = (double)blocks * 256 * ITERS * 8 * ops_per_inner[mode];
            printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"ops_per_clk_per_sm\": %.2f, \"gops\": %.1f, \"ms\": %.4f, \"mhz_eff\": %.0f}\n",
                   names[mode], bps, ops / cyc / sms, ops / ms / 1e6, ms, cyc / ms / 1e3);
        }
    }
    {
        double* d_dout; cudaMalloc(&d_dout, (size_t)max_blocks * 256 * 8);
        for (int bps = 2; bps <= 8; bps *= 2) {
            int blocks = sms * bps;
            for (int mix = 0; mix < 2; ++mix) {
                for (int rep = 0; rep < 2; ++rep) {
                    cudaEventRecord(e0);
                    if (mix == 0) k_dfma<0><<<blocks, 256>>>(d_dout, 1.5, 77, d_cycles); else k_dfma<1><<<blocks, 256>>>(d_dout, 1.5, 77, d_cycles);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                double ops = (double)blocks * 256 * ITERS * 8;
                printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"g_dfma_per_s\": %.1f, \"g_imad_wide_per_s\": %.1f, \"ms\": %.4f}\n",
                       mix ? "dfma_plus_imad_wide" : "dfma", bps, ops / ms / 1e6, mix ? ops / ms / 1e6 : 0.0, ms);
            }
        }
    }
    for (int bps = 1; bps <= 4; bps *= 2) {
        int blocks = sms * bps, iters = 512;
        for (int which = 0; which < 4; ++which) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_mul<Fr, 1><<<blocks, 256>>>(d_fe_out, d_fe_in, d_cycles, iters);
                if (which == 1) k_mul<Fr, 2><<<blocks, 256>>>(d_fe_out, d_fe_in, d_cycles, iters);
                if (which == 2) k_mul<Fq, 1><<<blocks, 256>>>(d_fe_out, d_fe_in, d_cycles, iters);
                if (which == 3) k_mul<Fq, 2><<<blocks, 256>>>(d_fe_out, d_fe_in, d_cycles, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double cyc = median_cycles(d_cycles, blocks);
            int chains = (which & 1) ? 2 : 1;
            double muls = (double)blocks * 256 * iters * chains;
            printf("{\"test\": \"%s_mul_chains%d\", \"blocks_per_sm\": %d, \"clk_per_mul_per_sm\": %.3f, \"gmuls\": %.2f, \"ms\": %.4f, \"mhz_eff\": %.0f}\n",
                   which < 2 ? "fr" : "fq", chains, bps, cyc * sms / muls, muls / ms / 1e6, ms, cyc / ms / 1e3);
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
