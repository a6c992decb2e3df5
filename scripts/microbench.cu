// Microbenchmarks of the B200 ceilings that bound the per-ray kernels ("L2 gather bandwidth,
// FP64/FP32 non-tensor peak — the builder must microbenchmark them").  Prints one JSON object.
//
//   issue      : warp-instructions / s with FMA-pipe and ALU-pipe instructions interleaved (one per cycle per
//                scheduler is the architectural limit: 4 x 148 x f_SM)
//   ffma/ffma2 : FP32 FMA pipe, scalar and packed (FFMA2: two FMAs per lane per instruction)
//   dfma       : FP64 FMA pipe
//   mufu       : MUFU.RSQ (the transcendental unit the steppers and the transfer lean on)
//   gather     : one aligned 128-byte line per thread as 8 x LDG.128 — the access of the ray stepper's cell
//                fetch from the cell-major polynomial cube — random lines within a working set that fits L1
//                (96 KB per SM), L2 (48 MB) and HBM (8 GB)
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench scripts/microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                    \
    do {                                                                                         \
        cudaError_t e = (x);                                                                     \
        if (e != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e)); \
            exit(1);                                                                             \
        }                                                                                        \
    } while (0)

constexpr int kIters = 4096;

__global__ void k_ffma(float *out, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_ffma2(float *out, float a, float b)
{
    float2 x0 = make_float2(threadIdx.x, 1.f), x1 = make_float2(2.f, threadIdx.x), x2 = x0, x3 = x1, x4 = x0, x5 = x1, x6 = x0, x7 = x1;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        x0 = __ffma2_rn(x0, a2, b2); x1 = __ffma2_rn(x1, a2, b2); x2 = __ffma2_rn(x2, a2, b2); x3 = __ffma2_rn(x3, a2, b2);
        x4 = __ffma2_rn(x4, a2, b2); x5 = __ffma2_rn(x5, a2, b2); x6 = __ffma2_rn(x6, a2, b2); x7 = __ffma2_rn(x7, a2, b2);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0.x + x1.y + x2.x + x3.y + x4.x + x5.y + x6.x + x7.y;
}

__global__ void k_dfma(double *out, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_mufu(float *out, float a)
{
    float x0 = threadIdx.x + 1.f, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x0)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x2)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x3));
        asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x4)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x5));
        asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x6)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x7));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + a;
}

// FMA pipe + ALU pipe interleaved: 8 FFMA and 8 integer (LOP3 / IADD3) instructions per iteration, all independent
__global__ void k_issue(float *out, float a, float b, unsigned m)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3, u4 = u0 + 4, u5 = u0 + 5, u6 = u0 + 6, u7 = u0 + 7;
#pragma unroll 4
    for (int i = 0; i < kIters; ++i) {
        x0 = fmaf(x0, a, b); u0 = (u0 ^ m) + 0x9e37u;
        x1 = fmaf(x1, a, b); u1 = (u1 & m) ^ u0;
        x2 = fmaf(x2, a, b); u2 = (u2 ^ m) + 0x79b9u;
        x3 = fmaf(x3, a, b); u3 = (u3 | m) ^ u2;
        x4 = fmaf(x4, a, b); u4 = (u4 ^ m) + 0x7f4au;
        x5 = fmaf(x5, a, b); u5 = (u5 & m) ^ u4;
        x6 = fmaf(x6, a, b); u6 = (u6 ^ m) + 0x7c15u;
        x7 = fmaf(x7, a, b); u7 = (u7 | m) ^ u6;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] =
        x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float)(u0 ^ u1 ^ u2 ^ u3 ^ u4 ^ u5 ^ u6 ^ u7);
}

__device__ __forceinline__ uint32_t xorshift(uint32_t &s)
{
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s;
}

// Each thread reads `n_lines` random aligned 128-byte lines as 8 x LDG.128 from a table of `lines` lines.
// coherent = 1: the 32 lanes of a warp pick lines from a window of 4 neighbouring lines (what a warp of
// neighbouring rays does); 0: 32 unrelated lines.
__global__ void k_gather(const float4 *__restrict__ tab, uint32_t lines, int n_lines, int coherent, float *out)
{
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t sw = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2246822519u + 777u;
    float acc = 0.f;
    for (int i = 0; i < n_lines; ++i) {
        uint32_t line;
        if (coherent) line = (xorshift(sw) % lines + (xorshift(s) & 3u)) % lines;
        else line = xorshift(s) % lines;
        const float4 *p = tab + (size_t)line * 8;
        const float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
        const float4 q4 = __ldg(p + 4), q5 = __ldg(p + 5), q6 = __ldg(p + 6), q7 = __ldg(p + 7);
        acc += q0.x + q1.y + q2.z + q3.w + q4.x + q5.y + q6.z + q7.w;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
static double time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int block = 256, blocks = sms * 8;           // 64 warps per SM: every scheduler has 16 warps to pick from
    const double threads = (double)block * blocks, warps = threads / 32.0;
    float *outf; double *outd;
    CK(cudaMalloc(&outf, threads * sizeof(float))); CK(cudaMalloc(&outd, threads * sizeof(double)));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);

    printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f,\n", prop.name, sms, clk_khz / 1e3);
    double ms;
    ms = time_ms([&] { k_ffma<<<blocks, block>>>(outf, 1.0001f, 0.5f); });
    const double ffma_warp = warps * 8.0 * kIters / (ms * 1e-3);
    printf(" \"ffma\": {\"warp_inst_per_s\": %.4e, \"tflops\": %.2f, \"warp_inst_per_clk_per_sm_at_max_clock\": %.3f},\n", ffma_warp,
           ffma_warp * 64.0 / 1e12, ffma_warp / (sms * clk_khz * 1e3));
    ms = time_ms([&] { k_ffma2<<<blocks, block>>>(outf, 1.0001f, 0.5f); });
    const double ffma2_warp = warps * 8.0 * kIters / (ms * 1e-3);
    printf(" \"ffma2\": {\"warp_inst_per_s\": %.4e, \"tflops\": %.2f, \"warp_inst_per_clk_per_sm_at_max_clock\": %.3f},\n", ffma2_warp,
           ffma2_warp * 128.0 / 1e12, ffma2_warp / (sms * clk_khz * 1e3));
    ms = time_ms([&] { k_dfma<<<blocks, block>>>(outd, 1.0001, 0.5); });
    const double dfma_warp = warps * 8.0 * kIters / (ms * 1e-3);
    printf(" \"dfma\": {\"warp_inst_per_s\": %.4e, \"tflops\": %.2f, \"warp_inst_per_clk_per_sm_at_max_clock\": %.3f},\n", dfma_warp,
           dfma_warp * 64.0 / 1e12, dfma_warp / (sms * clk_khz * 1e3));
    ms = time_ms([&] { k_mufu<<<blocks, block>>>(outf, 0.f); });
    const double mufu_warp = warps * 8.0 * kIters / (ms * 1e-3);
    printf(" \"mufu_rsq\": {\"warp_inst_per_s\": %.4e, \"warp_inst_per_clk_per_sm_at_max_clock\": %.3f},\n", mufu_warp,
           mufu_warp / (sms * clk_khz * 1e3));
    ms = time_ms([&] { k_issue<<<blocks, block>>>(outf, 1.0001f, 0.5f, 0x5bd1e995u); });
    // SASS of the loop body (cuobjdump): 8 FFMA + 12 ALU (8 LOP3 + 4 IADD3/VIADD) per iteration
    const double issue_warp = warps * 20.0 * kIters / (ms * 1e-3);
    printf(" \"issue_fma_plus_alu\": {\"warp_inst_per_s\": %.4e, \"warp_inst_per_clk_per_sm_at_max_clock\": %.3f, "
           "\"architectural_peak_warp_inst_per_s_at_max_clock\": %.4e},\n",
           issue_warp, issue_warp / (sms * clk_khz * 1e3), 4.0 * sms * clk_khz * 1e3);

    // ---- gathers ----
    struct Case { const char *name; size_t bytes; int coherent; };
    const Case cases[] = {{"l1_96KB_per_sm_window", (size_t)96 << 10, 1}, {"l2_48MB_coherent", (size_t)48 << 20, 1},
                          {"l2_48MB_random", (size_t)48 << 20, 0},        {"hbm_8GB_coherent", (size_t)8 << 30, 1},
                          {"hbm_8GB_random", (size_t)8 << 30, 0}};
    float4 *tab;
    CK(cudaMalloc(&tab, (size_t)8 << 30));
    CK(cudaMemset(tab, 0, (size_t)8 << 30));
    printf(" \"gather_128B_lines\": {\n");
    for (size_t c = 0; c < sizeof(cases) / sizeof(cases[0]); ++c) {
        const uint32_t lines = (uint32_t)(cases[c].bytes / 128);
        const int n_lines = 256;
        const int gb = sms * 16, gt = 128;
        ms = time_ms([&] { k_gather<<<gb, gt>>>(tab, lines, n_lines, cases[c].coherent, outf); }, 3);
        const double bytes = (double)gb * gt * n_lines * 128.0;
        printf("  \"%s\": {\"thread_GBps\": %.1f, \"lines_per_s\": %.4e}%s\n", cases[c].name, bytes / (ms * 1e-3) / 1e9,
               bytes / 128.0 / (ms * 1e-3), c + 1 < sizeof(cases) / sizeof(cases[0]) ? "," : "");
    }
    printf(" },\n \"note\": \"thread_GBps counts 128 B per thread per line, i.e. the bytes the threads consume; lanes of a warp that share a line are served by one L1 wavefront\"\n}\n");
    return 0;
}
