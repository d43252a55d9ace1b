// pipe_probe.cu — instruction-level throughput probes for the packed FP32 forms the forward kernel uses (NOT product).
// Reports SMSP cycles per warp-instruction of the probed class at full occupancy, to learn what bounds FFMA2/FMUL2
// with scalar-broadcast / pair operands and how LOP3->P + FSEL and SHFL share issue and register-read bandwidth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -ftz=true -fmad=false -o tools/pipe_probe tools/pipe_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

constexpr int N = 8;      // independent chains per thread
constexpr int IT = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, const float* in, uint32_t m)
{
    float2 acc[N], c[N], d[N];
    float s[N], t[N];
    uint32_t w = in[threadIdx.x & 7] > 0.f ? m + threadIdx.x : m;
    uint32_t w2 = w ^ 0x5bd1e995u;
    float selacc = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float v = in[i] + (float)threadIdx.x * 1e-7f;
        acc[i] = make_float2(v, v * 0.5f); c[i] = make_float2(v * 1e-3f, v * 2e-3f); d[i] = make_float2(0.999f + v * 1e-6f, 0.998f);
        s[i] = 0.999f + v * 1e-6f; t[i] = v * 1e-3f;
    }
    for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 1) acc[i] = __ffma2_rn(acc[i], bc(s[i]), c[i]);            // pair * scalar + pair
                if (MODE == 2) acc[i] = __fmul2_rn(acc[i], bc(s[i]));                  // pair * scalar
                if (MODE == 3) acc[i] = __fmul2_rn(acc[i], d[i]);                      // pair * pair
                if (MODE == 4) acc[i] = __ffma2_rn(acc[i], d[i], c[i]);                // pair * pair + pair
                if (MODE == 5) { acc[i].x = __fmaf_rn(acc[i].x, s[i], t[i]); acc[i].y = __fmaf_rn(acc[i].y, s[(i + 1) % N], t[(i + 3) % N]); }   // scalar, distinct regs
                if (MODE == 6) acc[i] = __ffma2_rn(acc[i], bc(s[0]), c[i]);            // shared scalar
                if (MODE == 7) acc[i] = __ffma2_rn(acc[i], bc(s[i]), c[0]);            // shared addend
                if (MODE == 8) acc[i] = __ffma2_rn(c[i], bc(s[i]), acc[i]);            // accumulate form
                if (MODE == 9) acc[i] = __fadd2_rn(acc[i], c[i]);
                if (MODE == 10 || MODE == 11 || MODE == 12) {                          // FMUL2 by a selected pair (2 LOP3->P + 2 FSEL), alone / after 3 or 7 FP2
                    if (MODE == 10) { acc[i] = __ffma2_rn(acc[i], bc(s[i]), c[i]); acc[i] = __fmul2_rn(acc[i], bc(t[i])); acc[i] = __ffma2_rn(c[i], bc(s[(i + 1) % N]), acc[i]); }
                    if (MODE == 12) {
                        acc[i] = __ffma2_rn(acc[i], bc(s[i]), c[i]); acc[i] = __fmul2_rn(acc[i], bc(t[i])); acc[i] = __ffma2_rn(c[i], bc(s[(i + 1) % N]), acc[i]);
                        acc[i] = __fmul2_rn(acc[i], bc(s[(i + 2) % N])); acc[i] = __ffma2_rn(acc[i], bc(s[(i + 3) % N]), d[i]); acc[i] = __fmul2_rn(acc[i], bc(t[(i + 1) % N]));
                        acc[i] = __ffma2_rn(acc[i], bc(s[(i + 5) % N]), c[(i + 1) % N]);
                    }
                    float2 p;
                    p.x = (w & (0xFu << ((i + r) & 7) * 4)) ? s[i] : t[i];
                    p.y = (w2 & (0xFu << ((i + r) & 7) * 4)) ? s[i] : t[i];
                    acc[i] = __fmul2_rn(acc[i], p);
                }
                if (MODE == 13) {                                                        // FFMA2 + SHFL every 4th
                    acc[i] = __ffma2_rn(acc[i], bc(s[i]), c[i]);
                    if ((i & 3) == 3) acc[i].x = __shfl_up_sync(0xffffffffu, acc[i].x, 1);
                }
                if (MODE == 14) { acc[i].x = __fmul_rn(acc[i].x, s[i]); acc[i].y = __fmul_rn(acc[i].y, s[(i + 1) % N]); }   // scalar FMUL
            }
            w = w * 1664525u + 1013904223u; w2 = w2 * 22695477u + 1u;
        }
    }
    float r = selacc;
#pragma unroll
    for (int i = 0; i < N; ++i) r += acc[i].x + acc[i].y;
    if (r == 123.456f) out[0] = r;
}

template <int MODE>
void run(const char* name, double inst_per_inner, float* out, float* in)
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, probe<MODE>));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<MODE>, 256, 0));
    const int grid = prop.multiProcessorCount * occ;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        probe<MODE><<<grid, 256>>>(out, in, 0x13572468u);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double warps_per_smsp = occ * 8 / 4.0;
    const double inner = (double)IT * 4 * N;                                  // inner bodies per thread
    const double cyc = best * 1e-3 * khz * 1e3 / (inner * warps_per_smsp);   // SMSP cycles per inner body per warp
    printf("%-44s regs %3d occ %d  %7.3f ms  %6.3f cyc per body (%g probed instr/body -> %.3f cyc each)\n", name, fa.numRegs, occ, best, cyc,
           inst_per_inner, cyc / inst_per_inner);
    fflush(stdout);
}

int main()
{
    float *out, *in;
    CK(cudaMalloc(&out, 256)); CK(cudaMalloc(&in, 256));
    float h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0f + 0.01f * i;
    CK(cudaMemcpy(in, h, 256, cudaMemcpyHostToDevice));
    run<1>("FFMA2 pair*scalar+pair (distinct)", 1, out, in);
    run<8>("FFMA2 pair*scalar+acc", 1, out, in);
    run<6>("FFMA2 pair*scalar(shared)+pair", 1, out, in);
    run<7>("FFMA2 pair*scalar+pair(shared)", 1, out, in);
    run<4>("FFMA2 pair*pair+pair", 1, out, in);
    run<2>("FMUL2 pair*scalar", 1, out, in);
    run<3>("FMUL2 pair*pair", 1, out, in);
    run<9>("FADD2 pair+pair", 1, out, in);
    run<5>("2x scalar FFMA (distinct regs)", 2, out, in);
    run<14>("2x scalar FMUL", 2, out, in);
    run<11>("FMUL2 pair*sel + 2x(LOP3,FSEL)", 1, out, in);
    run<10>("3 FP2 + FMUL2 pair*sel + 2x(LOP3,FSEL)", 4, out, in);
    run<12>("7 FP2 + FMUL2 pair*sel + 2x(LOP3,FSEL)", 8, out, in);
    run<13>("FFMA2 + SHFL/4", 1, out, in);
    return 0;
}
