// kernels_api.h -- host-side launch interface between engine.cu and the kernel translation units.
#pragma once
#include <string>

#include "common.cuh"
#include "reduce_kernels_args.h"

namespace mppi {

// which instantiation family of rollout_kernel
enum RolloutKind { ROLLOUT_F32_SOFTMIN = 0, ROLLOUT_F32_SCREEN = 1, ROLLOUT_F64_SOFTMIN = 2 };

// which code path of that family (engine.cu: try_configure picks the leanest one the parameters admit)
//   GENERAL: replayed noise from HBM and/or large yaw increments (full-range trig, multi-turn wrap)
//   FAST   : Philox noise in registers, |dt * yaw rate| <= pi/4: branch-free four-step blocks (rollout_kernel.cuh)
//   LEAN   : fp32 only, additionally Q[2] == 0 and |dt * yaw rate| <= 1/8 (rollout_lean_kernel.cuh)
enum RolloutVariant { ROLLOUT_GENERAL = 0, ROLLOUT_FAST = 1, ROLLOUT_LEAN = 2 };

// occupancy query + opt-in to large dynamic shared memory for one instantiation
cudaError_t rollout_prepare(int kind, int model, bool has_grid, int block, int variant, size_t smem, int* ctas_per_sm, int* regs);
cudaError_t rollout_launch(int kind, int model, bool has_grid, int block, int variant, int grid, size_t smem, cudaStream_t st,
                           const RolloutArgs& a);
size_t rollout_smem(int kind, int T, int block, int variant, int grid_bytes_in_smem);

cudaError_t reduce_softmin_launch(bool f64, int T, cudaStream_t st, const ReduceArgs& a);
cudaError_t reduce_screen_launch(int model, bool has_grid, int T, cudaStream_t st, const ReduceArgs& a);
cudaError_t finalize_launch(cudaStream_t st, const FinalizeArgs& a);
cudaError_t prep_nominal_launch(cudaStream_t st, const DynState* dyn, const StaticParams& sp, const double* Umaster, float* nomF, double* nomD);   // nomF: 8*T floats
cudaError_t flush_l2_launch(cudaStream_t st, void* buf, size_t bytes, unsigned int value);
cudaError_t rendezvous_launch(cudaStream_t st, const RendezvousArgs& a);
cudaError_t noise_export_launch(cudaStream_t st, const StaticParams& sp, const DynState* dyn, unsigned step, double* eps);
cudaError_t weights_from_v_launch(cudaStream_t st, const StaticParams& sp, const DynState* dyn, const double* V,
                                  const double* eps, double* record);
cudaError_t model_step_launch(cudaStream_t st, const StaticParams& sp, const double* x, const double* u, int n, double* out);
cudaError_t fp32_peak_launch(cudaStream_t st, int blocks, int threads, float* out, int iters);

// ---- caller-supplied dynamics / cost functor: the step's kernels instantiated at run time (user_model.cu) ----------------
struct UserKernels;
mppi_status user_kernels_build(const mppi_user_model* um, UserKernels** out, std::string* err);
void user_kernels_free(UserKernels* uk);
mppi_status user_model_check(const mppi_user_model* um, std::string* err, size_t* cubin_bytes);
cudaError_t user_rollout_prepare(const UserKernels* uk, int kind, bool has_grid, size_t smem, int* ctas_per_sm, int* regs);
cudaError_t user_rollout_launch(const UserKernels* uk, int kind, bool has_grid, int grid, size_t smem, cudaStream_t st, const RolloutArgs& a);
cudaError_t user_reduce_screen_launch(const UserKernels* uk, bool has_grid, int T, cudaStream_t st, const ReduceArgs& a);
cudaError_t user_reduce_softmin_launch(const UserKernels* uk, bool f64, int T, cudaStream_t st, const ReduceArgs& a);
cudaError_t user_finalize_launch(const UserKernels* uk, cudaStream_t st, const FinalizeArgs& a);
cudaError_t user_model_step_launch(const UserKernels* uk, cudaStream_t st, const StaticParams& sp, const double* x, const double* u, int n,
                                   double* out);

}  // namespace mppi
