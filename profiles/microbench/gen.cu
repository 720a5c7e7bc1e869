// gen.cu -- what does the in-register noise generator cost NEXT TO the FP32 work of the rollout loop?  (B200, sm_100a)
//
// The rollout loop draws 128 random bits per three time steps (6 normals) and spends ~40 FP32 instructions per step on the
// ODE + cost.  This microbenchmark runs, per thread and iteration, ONE generator call (128 bits) beside NF dependent-chain
// FFMAs (4 independent chains, like the loop's ILP) and reports cycles per iteration per warp per scheduler, for
//   none        : the FFMAs alone
//   philox10    : Philox4x32-10 (20 IMAD.WIDE-class multiplies + xors; the round-1 generator)
//   philox7     : Philox4x32-7
//   threefry20  : Threefry4x32-20 (add / rotate / xor only; Random123 default)
//   threefry12  : Threefry4x32-12 (smallest Crush-resistant round count of the Random123 paper)
//   xoshiro     : four steps of xoshiro128++ (stateful, add / rotate / shift / xor only)
// at 4 and 8 warps per scheduler (512 / 1024 threads per SM).  Also plain issue rates of IADD3 / SHF beside FFMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench/gen profiles/microbench/gen.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <vector>

#define ITER 1024

enum Gen { NONE, PHILOX10, PHILOX7, THREEFRY20, THREEFRY12, XOSHIRO, IADD3_ONLY, SHF_ONLY, MIX_FFMA_IADD3, MIX_FFMA_SHF, NGEN };
static const char* gname[] = {"none", "philox4x32-10", "philox4x32-7", "threefry4x32-20", "threefry4x32-12", "xoshiro128++ x4",
                              "IADD3 x128", "SHF (rotate) x128", "FFMA + IADD3 (1:1) x64", "FFMA + SHF (1:1) x64"};

__device__ __forceinline__ uint32_t rotl(uint32_t x, int k) { return __funnelshift_l(x, x, k); }

template <int ROUNDS>
__device__ __forceinline__ uint4 philox(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < ROUNDS; ++i) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x, hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

template <int ROUNDS>
__device__ __forceinline__ uint4 threefry(uint4 c, uint4 k) {
  // Threefry-4x32 (Salmon et al., SC'11): rotation constants of the 4x32 variant, key injection every 4 rounds
  const int R[8][2] = {{10, 26}, {11, 21}, {13, 27}, {23, 5}, {6, 20}, {17, 11}, {25, 10}, {18, 20}};
  uint32_t ks[5] = {k.x, k.y, k.z, k.w, 0x1BD11BDAu ^ k.x ^ k.y ^ k.z ^ k.w};
  uint32_t x0 = c.x + ks[0], x1 = c.y + ks[1], x2 = c.z + ks[2], x3 = c.w + ks[3];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    if ((r & 1) == 0) {
      x0 += x1; x1 = rotl(x1, R[r & 7][0]) ^ x0;
      x2 += x3; x3 = rotl(x3, R[r & 7][1]) ^ x2;
    } else {
      x0 += x3; x3 = rotl(x3, R[r & 7][0]) ^ x0;
      x2 += x1; x1 = rotl(x1, R[r & 7][1]) ^ x2;
    }
    if ((r & 3) == 3) {
      const int s = r / 4 + 1;
      x0 += ks[s % 5]; x1 += ks[(s + 1) % 5]; x2 += ks[(s + 2) % 5]; x3 += ks[(s + 3) % 5] + s;
    }
  }
  return make_uint4(x0, x1, x2, x3);
}

__device__ __forceinline__ uint32_t xoshiro_next(uint32_t& s0, uint32_t& s1, uint32_t& s2, uint32_t& s3) {
  const uint32_t r = rotl(s0 + s3, 7) + s0;
  const uint32_t t = s1 << 9;
  s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t;
  s3 = rotl(s3, 11);
  return r;
}

template <int G, int NF>
__global__ void __launch_bounds__(1024, 1) bench(float* out, long long* cycles, float seed, unsigned iseed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  const float b = 0.999f + threadIdx.x * 1e-9f, c = 1e-3f + threadIdx.x * 1e-9f;
  uint32_t s0 = iseed + threadIdx.x * 977u, s1 = s0 * 31u + 1u, s2 = s1 * 17u + 3u, s3 = s2 ^ 0x9E3779B9u;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
    uint4 r = make_uint4(0, 0, 0, 0);
    if (G == PHILOX10) r = philox<10>(make_uint4(s0, s1, (uint32_t)it, iseed), make_uint2(s2, s3));
    if (G == PHILOX7) r = philox<7>(make_uint4(s0, s1, (uint32_t)it, iseed), make_uint2(s2, s3));
    if (G == THREEFRY20) r = threefry<20>(make_uint4(s0, s1, (uint32_t)it, iseed), make_uint4(s2, s3, 1u, 2u));
    if (G == THREEFRY12) r = threefry<12>(make_uint4(s0, s1, (uint32_t)it, iseed), make_uint4(s2, s3, 1u, 2u));
    if (G == XOSHIRO) {
      r.x = xoshiro_next(s0, s1, s2, s3);
      r.y = xoshiro_next(s0, s1, s2, s3);
      r.z = xoshiro_next(s0, s1, s2, s3);
      r.w = xoshiro_next(s0, s1, s2, s3);
    }
    if (G == IADD3_ONLY) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { s0 += s1 + iseed; s1 += s2 + iseed; s2 += s3 + iseed; s3 += s0 + iseed; }
    }
    if (G == SHF_ONLY) {
#pragma unroll
      for (int j = 0; j < 32; ++j) { s0 = __funnelshift_l(s0, s1, 7); s1 = __funnelshift_l(s1, s2, 9); s2 = __funnelshift_l(s2, s3, 11); s3 = __funnelshift_l(s3, s0, 13); }
    }
    if (G == MIX_FFMA_IADD3) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s0 += s1 + iseed; a0 = fmaf(a0, b, c); s1 += s2 + iseed; a1 = fmaf(a1, b, c); s2 += s3 + iseed; a2 = fmaf(a2, b, c); s3 += s0 + iseed; a3 = fmaf(a3, b, c);
      }
    }
    if (G == MIX_FFMA_SHF) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        s0 = __funnelshift_l(s0, s1, 7); a0 = fmaf(a0, b, c); s1 = __funnelshift_l(s1, s2, 9); a1 = fmaf(a1, b, c);
        s2 = __funnelshift_l(s2, s3, 11); a2 = fmaf(a2, b, c); s3 = __funnelshift_l(s3, s0, 13); a3 = fmaf(a3, b, c);
      }
    }
    acc ^= r.x ^ r.y ^ r.z ^ r.w;
#pragma unroll
    for (int j = 0; j < NF / 4; ++j) {
      a0 = fmaf(a0, b, c);
      a1 = fmaf(a1, b, c);
      a2 = fmaf(a2, b, c);
      a3 = fmaf(a3, b, c);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + (float)(acc ^ s0 ^ s1 ^ s2 ^ s3);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int G, int NF>
double run(int nsm, int threads, float* out, long long* cyc) {
  for (int rep = 0; rep < 2; ++rep) {
    bench<G, NF><<<nsm, threads>>>(out, cyc, 1.25f, 12345u);
    cudaDeviceSynchronize();
  }
  std::vector<long long> h(nsm);
  cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(h.begin(), h.end());
  const double warps_per_sched = threads / 32.0 / 4.0;
  return (double)h[nsm / 2] / ITER / warps_per_sched;   // cycles per iteration per warp, per scheduler
}

template <int G>
void row(int nsm, float* out, long long* cyc) {
  printf("%-24s", gname[G]);
  for (int threads : {512, 1024}) {
    const double alone = run<G, 0>(nsm, threads, out, cyc);
    const double with120 = run<G, 120>(nsm, threads, out, cyc);
    printf("  | %4d thr: alone %7.1f  with 120 FFMA %7.1f", threads, alone, with120);
  }
  printf("\n");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("%s, %d SMs; cycles per iteration (one generator call = 128 bits [+ 120 FFMA in 4 chains]) per warp per scheduler\n", p.name, nsm);
  float* out;
  long long* cyc;
  cudaMalloc(&out, (size_t)nsm * 1024 * sizeof(float));
  cudaMalloc(&cyc, nsm * sizeof(long long));
  row<NONE>(nsm, out, cyc);
  row<PHILOX10>(nsm, out, cyc);
  row<PHILOX7>(nsm, out, cyc);
  row<THREEFRY20>(nsm, out, cyc);
  row<THREEFRY12>(nsm, out, cyc);
  row<XOSHIRO>(nsm, out, cyc);
  row<IADD3_ONLY>(nsm, out, cyc);
  row<SHF_ONLY>(nsm, out, cyc);
  row<MIX_FFMA_IADD3>(nsm, out, cyc);
  row<MIX_FFMA_SHF>(nsm, out, cyc);
  return 0;
}
