// pipes.cu -- per-SM issue throughput of the instruction classes the MPPI rollout kernel uses (B200, sm_100a).
// Each test runs ITER iterations of 8 independent dependency chains of ONE operation per thread, 1024 threads per SM
// (8 warps per scheduler), and reports warp-instructions per cycle per SM from clock64().  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench/pipes profiles/microbench/pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define ITER 2048

enum Op { FFMA3, FFMA_IMM, FMUL_, FADD_, IMADW, LOP3_, FMNMX_, SEL_, MUFU_LG2, MUFU_SQRT, MUFU_SIN, MUFU_COS, MUFU_EX2, MUFU_RCP,
          MUFU_RSQ, I2FP_, F2I_, FRND_, REDUX_, SHFL_, MIX_FFMA_LOP3, MIX_FFMA_FMNMX, MIX_FFMA_MUFU, LDS128_BCAST, PHILOX_RND, IMAD_LO, IMAD_HI, MIX_FFMA_IMADW, MIX_FFMA3_IMADW, MIX_LOP3_IMADW, FFMA_SAT, NOPS };
static const char* names[] = {"FFMA (3 registers)", "FFMA (immediate/const operands)", "FMUL", "FADD", "IMAD.WIDE.U32", "LOP3 (xor3)", "FMNMX",
                              "SEL", "MUFU.LG2", "sqrt.approx (MUFU.SQRT/RSQ)", "sin.approx (FMUL+MUFU.SIN)", "cos.approx (FMUL+MUFU.COS)",
                              "MUFU.EX2", "MUFU.RCP", "MUFU.RSQ", "I2FP.F32.U32", "F2I (rn)", "FRND (rint)", "REDUX.SUM",
                              "SHFL.BFLY", "FFMA + LOP3 interleaved (1:1)", "FFMA + FMNMX interleaved (1:1)", "FFMA + MUFU.LG2 (3:1)",
                              "LDS.128 broadcast", "Philox half-round (IMAD.WIDE + LOP3)", "IMAD (32-bit low)", "IMAD.HI.U32",
                              "FFMA + (IMAD.WIDE+LOP3) interleaved (1:1:1)", "3 FFMA + (IMAD.WIDE+LOP3)", "LOP3 + (IMAD.WIDE+LOP3)",
                              "FFMA.SAT"};
static const int ops_per_iter[] = {8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 16, 16, 8, 8, 16, 8, 8, 24, 40, 24, 8};

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cycles, float seed, int iseed) {
  __shared__ float4 sm4[64];
  if (threadIdx.x < 64) sm4[threadIdx.x] = make_float4(seed, seed, seed, seed);
  float a[8], b = seed + 1.0f, c = seed * 0.5f;
  unsigned u[8];
  unsigned long long w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + threadIdx.x * 1e-3f + i;
    u[i] = iseed + threadIdx.x * 977u + i;
    w[i] = u[i];
  }
  float b2 = b + threadIdx.x, c2 = c + threadIdx.x;   // genuinely register operands
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == FFMA3) a[i] = fmaf(a[i], b2, c2);
      if (OP == FFMA_IMM) a[i] = fmaf(a[i], 0.999f, 1e-3f);
      if (OP == FMUL_) a[i] = a[i] * b2;
      if (OP == FADD_) a[i] = a[i] + b2;
      if (OP == IMADW) w[i] = (unsigned long long)(unsigned)w[i] * 0xD2511F53u + (w[i] >> 32);
      if (OP == LOP3_) u[i] = u[i] ^ u[(i + 1) & 7] ^ (unsigned)iseed;
      if (OP == FMNMX_) a[i] = fminf(a[i], a[(i + 1) & 7] + 0.f) ;
      if (OP == SEL_) u[i] = (u[(i + 3) & 7] & 1) ? u[i] : u[(i + 1) & 7];
      if (OP == MUFU_LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == MUFU_SQRT) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == MUFU_SIN) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == MUFU_COS) asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == MUFU_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == MUFU_RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == MUFU_RSQ) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == I2FP_) { a[i] = (float)u[i]; u[i] = __float_as_uint(a[i]); }
      if (OP == F2I_) { u[i] = (unsigned)__float2int_rn(a[i]); a[i] = __uint_as_float(u[i] | 0x3f800000u); }
      if (OP == FRND_) a[i] = rintf(a[i]) ;
      if (OP == REDUX_) u[i] = __reduce_add_sync(0xffffffffu, u[i]) + threadIdx.x;
      if (OP == SHFL_) u[i] = __shfl_xor_sync(0xffffffffu, u[i], 1);
      if (OP == MIX_FFMA_LOP3) { a[i] = fmaf(a[i], b2, c2); u[i] = u[i] ^ u[(i + 1) & 7] ^ (unsigned)iseed; }
      if (OP == MIX_FFMA_FMNMX) { a[i] = fmaf(a[i], b2, c2); u[i] = min(u[i], u[(i + 1) & 7] + 1u); }
      if (OP == MIX_FFMA_MUFU) {
        if (i % 4 == 3) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        else a[i] = fmaf(a[i], b2, c2);
      }
      if (OP == PHILOX_RND) { const unsigned long long p = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(p >> 32) ^ (unsigned)p ^ (unsigned)iseed; }
      if (OP == IMAD_LO) u[i] = u[i] * 0xD2511F53u + (unsigned)iseed;
      if (OP == IMAD_HI) u[i] = __umulhi(u[i], 0xD2511F53u) + (unsigned)iseed;
      if (OP == MIX_FFMA_IMADW) { a[i] = fmaf(a[i], b2, c2); const unsigned long long p = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(p >> 32) ^ (unsigned)p ^ (unsigned)iseed; }
      if (OP == MIX_FFMA3_IMADW) { a[i] = fmaf(a[i], b2, c2); a[i] = fmaf(a[i], c2, b2); a[i] = fmaf(a[i], b2, c2); const unsigned long long p = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(p >> 32) ^ (unsigned)p ^ (unsigned)iseed; }
      if (OP == MIX_LOP3_IMADW) { u[(i + 4) & 7] = u[(i + 4) & 7] ^ u[(i + 5) & 7] ^ (unsigned)it; const unsigned long long p = (unsigned long long)u[i] * 0xD2511F53u; u[i] = (unsigned)(p >> 32) ^ (unsigned)p ^ (unsigned)iseed; }
      if (OP == FFMA_SAT) a[i] = __saturatef(fmaf(a[i], b2, c2));
      if (OP == LDS128_BCAST) { float4 v = sm4[(it + i) & 63]; a[i] += v.x + v.w; }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + (float)u[i] + (float)w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int nsm, float* out, long long* cyc) {
  bench<OP><<<nsm, 1024>>>(out, cyc, 1.25f, 12345);
  cudaDeviceSynchronize();
  bench<OP><<<nsm, 1024>>>(out, cyc, 1.25f, 12345);
  cudaDeviceSynchronize();
  std::vector<long long> h(nsm);
  cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double c = (double)h[nsm / 2];
  const double winst = (double)ITER * ops_per_iter[OP] * 32.0;   // warp-instructions per SM (32 warps)
  printf("%-40s %8.3f warp-inst/clk/SM  (%6.2f cycles per warp-inst per scheduler)\n", names[OP], winst / c, c / (winst / 4.0));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("%s, %d SMs, ITER=%d, 32 warps/SM, 8 independent chains per thread\n", p.name, nsm, ITER);
  float* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)nsm * 1024 * sizeof(float));
  cudaMalloc(&cyc, nsm * sizeof(long long));
  run<FFMA3>(nsm, out, cyc);
  run<FFMA_IMM>(nsm, out, cyc);
  run<FMUL_>(nsm, out, cyc);
  run<FADD_>(nsm, out, cyc);
  run<IMADW>(nsm, out, cyc);
  run<LOP3_>(nsm, out, cyc);
  run<FMNMX_>(nsm, out, cyc);
  run<SEL_>(nsm, out, cyc);
  run<MUFU_LG2>(nsm, out, cyc);
  run<MUFU_SQRT>(nsm, out, cyc);
  run<MUFU_SIN>(nsm, out, cyc);
  run<MUFU_COS>(nsm, out, cyc);
  run<MUFU_EX2>(nsm, out, cyc);
  run<MUFU_RCP>(nsm, out, cyc);
  run<MUFU_RSQ>(nsm, out, cyc);
  run<I2FP_>(nsm, out, cyc);
  run<F2I_>(nsm, out, cyc);
  run<FRND_>(nsm, out, cyc);
  run<REDUX_>(nsm, out, cyc);
  run<SHFL_>(nsm, out, cyc);
  run<MIX_FFMA_LOP3>(nsm, out, cyc);
  run<MIX_FFMA_FMNMX>(nsm, out, cyc);
  run<MIX_FFMA_MUFU>(nsm, out, cyc);
  run<LDS128_BCAST>(nsm, out, cyc);
  run<PHILOX_RND>(nsm, out, cyc);
  run<IMAD_LO>(nsm, out, cyc);
  run<IMAD_HI>(nsm, out, cyc);
  run<MIX_FFMA_IMADW>(nsm, out, cyc);
  run<MIX_FFMA3_IMADW>(nsm, out, cyc);
  run<MIX_LOP3_IMADW>(nsm, out, cyc);
  run<FFMA_SAT>(nsm, out, cyc);
  return 0;
}
