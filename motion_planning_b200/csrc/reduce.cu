// reduce.cu -- instantiations + launch wrappers of the reduce / finalize / auxiliary kernels,
// and the kind-dispatch for the three rollout translation units.
#include <cstring>

#include "kernels_api.h"
#include "reduce_kernels.cuh"
#include "rollout_kernel.cuh"
#include "rollout_lean_kernel.cuh"

namespace mppi {

// launch with programmatic stream serialization: the kernel may start while its predecessor in the
// stream is still draining; it synchronises on-device with griddepcontrol.wait
template <typename Fn>
static cudaError_t launch_pdl(Fn f, int grid, int block, size_t smem, cudaStream_t st, const ReduceArgs& a) {
  // same L1 / shared-memory split as the rollout kernels (rollout_tu.inc): no SM reconfiguration inside a step
  static thread_local const void* carved[16] = {};
  bool seen = false;
  for (const void* p : carved) seen |= (p == (const void*)f);
  if (!seen) {
    cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    for (auto& p : carved)
      if (!p) {
        p = (const void*)f;
        break;
      }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, f, a);
}

#define DECL_TU(name)                                                                                          \
  cudaError_t rollout_##name##_prepare(int model, bool has_grid, int block, int variant, size_t smem, int* ctas, int* regs); \
  cudaError_t rollout_##name##_launch(int model, bool has_grid, int block, int variant, int grid, size_t smem, cudaStream_t st, \
                                      const RolloutArgs& a);
DECL_TU(f32_softmin)
DECL_TU(f32_screen)
DECL_TU(f64_softmin)
#undef DECL_TU

cudaError_t rollout_prepare(int kind, int model, bool has_grid, int block, int variant, size_t smem, int* ctas, int* regs) {
  switch (kind) {
    case ROLLOUT_F32_SOFTMIN: return rollout_f32_softmin_prepare(model, has_grid, block, variant, smem, ctas, regs);
    case ROLLOUT_F32_SCREEN: return rollout_f32_screen_prepare(model, has_grid, block, variant, smem, ctas, regs);
    default: return rollout_f64_softmin_prepare(model, has_grid, block, variant, smem, ctas, regs);
  }
}

cudaError_t rollout_launch(int kind, int model, bool has_grid, int block, int variant, int grid, size_t smem, cudaStream_t st,
                           const RolloutArgs& a) {
  switch (kind) {
    case ROLLOUT_F32_SOFTMIN: return rollout_f32_softmin_launch(model, has_grid, block, variant, grid, smem, st, a);
    case ROLLOUT_F32_SCREEN: return rollout_f32_screen_launch(model, has_grid, block, variant, grid, smem, st, a);
    default: return rollout_f64_softmin_launch(model, has_grid, block, variant, grid, smem, st, a);
  }
}

size_t rollout_smem(int kind, int T, int block, int variant, int grid_bytes_in_smem) {
  if (variant == ROLLOUT_LEAN && block == 512) return rollout_lean_sm_smem_bytes(T, grid_bytes_in_smem);
  if (variant == ROLLOUT_LEAN) return rollout_lean_smem_bytes(T, block, grid_bytes_in_smem);
  return kind == ROLLOUT_F64_SOFTMIN ? rollout_smem_bytes<double>(T, block, grid_bytes_in_smem)
                                     : rollout_smem_bytes<float>(T, block, grid_bytes_in_smem);
}

// shared memory of the finalize phase: the 4*T doubles of the update
static size_t finalize_smem(const StaticParams& sp, bool) { return (size_t)4 * sp.T * sizeof(double); }

template <typename Fn>
static cudaError_t opt_in_smem(Fn f, size_t smem) {
  if (smem <= 24 * 1024) return cudaSuccess;   // static (up to ~13 KB) + dynamic shared memory beyond 48 KB needs the opt-in
  // once per (function, device, size): the attribute call is not free and this sits on the launch path of every step
  static thread_local struct { const void* f; int dev; size_t smem; } done[16] = {};
  int dev = -1;
  cudaGetDevice(&dev);
  for (auto& d : done)
    if (d.f == (const void*)f && d.dev == dev && d.smem >= smem) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess)
    for (auto& d : done)
      if (!d.f || (d.f == (const void*)f && d.dev == dev)) {
        d.f = (const void*)f;
        d.dev = dev;
        d.smem = smem;
        break;
      }
  return e;
}

// grid: one block per time step (+ the finalizer block of a fused step)
cudaError_t reduce_softmin_launch(bool f64, int T, cudaStream_t st, const ReduceArgs& a) {
  const size_t smem = finalize_smem(a.sp, a.fused != 0);
  const int grid = T + (a.fused ? 1 : 0);
  cudaError_t e = f64 ? opt_in_smem(reduce_softmin_kernel<double>, smem) : opt_in_smem(reduce_softmin_kernel<float>, smem);
  if (e != cudaSuccess) return e;
  if (f64) return launch_pdl(reduce_softmin_kernel<double>, grid, 256, smem, st, a);
  return launch_pdl(reduce_softmin_kernel<float>, grid, 256, smem, st, a);
}

typedef void (*ScreenFn)(const ReduceArgs);
template <int M, bool G>
static ScreenFn sfn_() { return reduce_screen_kernel<M, G>; }

cudaError_t reduce_screen_launch(int model, bool has_grid, int T, cudaStream_t st, const ReduceArgs& a) {
  ScreenFn f;
  switch (model) {
    case MPPI_MODEL_DIFF_DRIVE: f = has_grid ? sfn_<MPPI_MODEL_DIFF_DRIVE, true>() : sfn_<MPPI_MODEL_DIFF_DRIVE, false>(); break;
    case MPPI_MODEL_UNICYCLE_EULER:
      f = has_grid ? sfn_<MPPI_MODEL_UNICYCLE_EULER, true>() : sfn_<MPPI_MODEL_UNICYCLE_EULER, false>();
      break;
    default: f = has_grid ? sfn_<MPPI_MODEL_BICYCLE, true>() : sfn_<MPPI_MODEL_BICYCLE, false>(); break;
  }
  size_t smem = (size_t)(8 * 7 + 4) * T * sizeof(double);   // scan scratch of 8 warps + nominal block
  if (finalize_smem(a.sp, a.fused != 0) > smem) smem = finalize_smem(a.sp, a.fused != 0);   // the finalizer block's needs
  cudaError_t e = opt_in_smem(f, smem);
  if (e != cudaSuccess) return e;
  return launch_pdl(f, T + (a.fused ? 1 : 0), 256, smem, st, a);
}

cudaError_t finalize_launch(cudaStream_t st, const FinalizeArgs& a) {
  const size_t smem = (size_t)4 * a.sp.T * sizeof(double);
  finalize_kernel<<<1, 256, smem, st>>>(a);
  return cudaGetLastError();
}

// measurement aid: flush L2.  The first half of the buffer (> L2) is OVERWRITTEN -- every line the step will touch, data and
// instructions, is evicted -- and then the second half (> L2, never written) is READ, so that what stays in the write-back L2
// are clean lines: the timed step starts cold, but it is not charged the write-back of the flush's own dirty lines (which a
// store-only flush leaves behind for whoever misses next).  A kernel of ours (not cudaMemsetAsync) so that it runs with the
// same L1 / shared-memory split as the step's kernels: the timed step after it starts without an SM reconfiguration.
__global__ void __launch_bounds__(256) flush_l2_kernel(uint4* p, size_t n16_half, unsigned int v, unsigned int* sink) {
  const uint4 val = make_uint4(v, v, v, v);
  const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = first; i < n16_half; i += stride) p[i] = val;
  unsigned int acc = 0;
  const uint4* q = p + n16_half;
  for (size_t i = first; i < n16_half; i += stride) {
    const uint4 r = __ldcg(q + i);
    acc ^= r.x ^ r.y ^ r.z ^ r.w;
  }
  if (acc == 0x9e3779b9u) *sink = acc;   // (keeps the loads alive; the second half is all zeros)
}
cudaError_t flush_l2_launch(cudaStream_t st, void* buf, size_t bytes, unsigned int value) {
  static thread_local bool carved = false;
  if (!carved) {
    cudaFuncSetAttribute(flush_l2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    carved = true;
  }
  // layout of `buf`: [bytes/2 overwritten][bytes/2 - 16 read][16 bytes sink]
  const size_t half16 = bytes / 32 - 1;
  flush_l2_kernel<<<148 * 8, 256, 0, st>>>(reinterpret_cast<uint4*>(buf), half16, value * 0x01010101u,
                                            reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(buf) + bytes - 16));
  return cudaGetLastError();
}

// measurement aid (mppi_bench, world > 1): align the ranks before a timed step.  Every rank raises its flag in every peer's
// row buffer tail (uint32 [2 parity][kMaxFusedWorld] behind the rows) and waits for all of them -- a device-side rendezvous over
// NVLink, launched OUTSIDE the timed interval, so that the ranks' L2 flushes (whose durations jitter) do not show up as waiting
// time inside the step of the rank that happened to finish its flush first.
__global__ void rendezvous_kernel(RendezvousArgs a) {
  const int g = threadIdx.x;
  if (g >= a.world) return;
  const int par = (int)(a.epoch & 1u);
  unsigned int* mine = reinterpret_cast<unsigned int*>(a.peers[g] + a.rows_uint2) + par * kMaxFusedWorld + a.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(mine), "r"(a.epoch) : "memory");
  const unsigned int* theirs = reinterpret_cast<const unsigned int*>(a.peers[a.rank] + a.rows_uint2) + par * kMaxFusedWorld + g;
  const long long t0 = clock64();
  unsigned int v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(theirs) : "memory");
  } while (v != a.epoch && clock64() - t0 < (1LL << 33));
}
cudaError_t rendezvous_launch(cudaStream_t st, const RendezvousArgs& a) {
  rendezvous_kernel<<<1, 32, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t prep_nominal_launch(cudaStream_t st, const DynState* dyn, const StaticParams& sp, const double* Umaster, float* nomF,
                                double* nomD) {
  prep_nominal_kernel<<<1, 128, 0, st>>>(dyn, sp.T, sp.u_max[0], sp.u_max[1], Umaster, nomF, nomD);
  return cudaGetLastError();
}

cudaError_t noise_export_launch(cudaStream_t st, const StaticParams& sp, const DynState* dyn, unsigned step, double* eps) {
  noise_export_kernel<<<(sp.K + 127) / 128, 128, 0, st>>>(sp, dyn, step, eps);
  return cudaGetLastError();
}

cudaError_t weights_from_v_launch(cudaStream_t st, const StaticParams& sp, const DynState* dyn, const double* V,
                                  const double* eps, double* record) {
  weights_from_v_kernel<<<sp.T, 128, 0, st>>>(sp, dyn, V, eps, record);
  return cudaGetLastError();
}

cudaError_t model_step_launch(cudaStream_t st, const StaticParams& sp, const double* x, const double* u, int n, double* out) {
  model_step_kernel<<<(n + 127) / 128, 128, 0, st>>>(sp, x, u, n, out);
  return cudaGetLastError();
}

cudaError_t fp32_peak_launch(cudaStream_t st, int blocks, int threads, float* out, int iters) {
  fp32_peak_kernel<<<blocks, threads, 0, st>>>(out, iters);
  return cudaGetLastError();
}

}  // namespace mppi
