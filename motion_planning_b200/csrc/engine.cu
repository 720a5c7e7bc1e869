// engine.cu -- host side of libmppi_b200.so: the C ABI declared in include/mppi_b200.h.
//
// One engine = one CUDA device.  A step (= MPPI.get_path, control/src/mppi:85-102) is TWO kernel launches and no copy:
//   rollout_{lean,lean_sm,}_kernel --PDL--> reduce_{softmin,screen}_kernel (T row blocks with the row exchange + 1 finalizer block)
// x0 / goal ride in the kernels' argument buffers, the result block is stored by the finalize phase into mapped pinned
// host memory; all controller state (nominal U, noise step counter) stays resident in HBM.  mppi_bench issues the same two
// launches with x0 resident on the device (closed loop on the model; optionally from a captured CUDA graph).
#include <cuda_runtime.h>

#include <atomic>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kernels_api.h"

using namespace mppi;

static thread_local char g_err[512] = "";

static void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#define CK(call)                                                                                 \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));             \
      return MPPI_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

#define CKS(call)                       \
  do {                                  \
    mppi_status _s = (call);            \
    if (_s != MPPI_OK) return _s;       \
  } while (0)

// Host <-> device copies of the set-up / debug entry points run ON the engine's stream and are waited for there: a
// synchronous cudaMemcpy from pageable memory goes through the legacy stream, which a cudaStreamNonBlocking stream is not
// ordered with (the copy may return before its DMA has landed while a kernel of ours is already being launched).
#define COPY_SYNC(e, dst, src, bytes, kind)                                             \
  do {                                                                                  \
    CK(cudaMemcpyAsync((dst), (src), (bytes), (kind), (e)->stream));                    \
    CK(cudaStreamSynchronize((e)->stream));                                             \
  } while (0)

struct mppi_engine;
static cudaError_t memcpy_on(mppi_engine* e, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);

struct KindCfg {
  int variant = ROLLOUT_GENERAL;
  int block = 0, grid = 0, ntiles = 0, ctas_per_sm = 0, regs = 0;
  int nparts = 0;   // partial records per time step the rollout kernel writes (== grid, or 7 per CTA for the SM-wide kernel)
  size_t smem = 0;
  bool ready = false;
};

struct mppi_engine {
  mppi_params p;
  StaticParams sp;
  int dev = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  // device state
  DynState* d_dyn = nullptr;
  double *d_Umaster = nullptr, *d_Ulast = nullptr, *d_nomD = nullptr, *d_Utmp = nullptr;
  float* d_nomF = nullptr;
  double sg_a = 0, sg_b = 0, sg_inv_norm[4] = {0, 0, 0, 0};
  double *d_record = nullptr, *d_gather = nullptr, *d_record_tmp = nullptr;
  unsigned long long* d_debug_ts = nullptr;
  unsigned long long* d_debug_rts = nullptr;   // rollout kernel stamps (MPPI_EXP_TIMELINE builds)
  size_t debug_rts_ctas = 0;
  // row exchange of the fused step (reduce_kernels.cuh): this rank's flag-in-data buffer (exported through CUDA IPC when
  // world > 1) and the device array of all ranks' buffer pointers (own buffer + IPC mappings of the peers')
  uint2* d_ll = nullptr;
  uint2* d_ll2 = nullptr;          // merged rows of this rank: [2 parity][T][kRow2Words]
  size_t ll_bytes = 0, ll_rows_uint2 = 0;
  unsigned int rdv_epoch = 0;
  uint2* ll_peers[kMaxFusedWorld] = {};   // host copy of the ranks' buffer pointers: they travel in the kernel arguments
  std::vector<void*> p2p_opened;    // IPC mappings to close
  bool p2p_on = false;
  void* d_part = nullptr;
  double* d_epart = nullptr;
  float4* d_cand_meta = nullptr;
  uint2* d_cand = nullptr;
  size_t part_capacity_ctas = 0;
  // host-side phases of mppi_step (profiling aid, mppi_debug_host_timing): entry -> rollout launched -> reduce launched
  // -> result seen -> return, accumulated in nanoseconds
  double host_ns[4] = {0, 0, 0, 0};
  long long host_calls = 0;
  std::chrono::steady_clock::time_point tp_launch1, tp_launch2;
  // MIXED back-off: after a candidate-list overflow (the soft-min support has grown beyond what the fp32 screen lists,
  // typically within centimetres of the goal) the next `f64_holdoff` steps go straight to the fp64 pipeline instead of
  // paying for a doomed mixed attempt plus its redo; the hold-off doubles (8 .. 64 steps) while overflows persist
  int f64_holdoff = 0, f64_backoff = 8;
  bool attempted_mixed = false;
  int lean_split = 0;   // MPPI_B200_SPLIT: first step of the second pair of warps of the SM-wide kernel's shared tile (0 = default)
  signed char* d_grid = nullptr;
  signed char* h_grid_stage = nullptr;   // pinned staging of mppi_update_grid patches
  size_t grid_stage_cap = 0;
  cudaEvent_t grid_stage_ev = nullptr;
  double* d_eps_ext = nullptr;
  void* d_vcap = nullptr;
  void* d_flush = nullptr;
  size_t flush_bytes = 0;
  KindCfg cfg[3];
  bool user_has_cost = false;
  int user_kind = 0;                        // 0 ODE functor, 1 kinematic functor (mppi_user_model.kind)
  double user_speed_max = 0, user_yaw_max = 0;   // kind 1: bounds of |speed|, |yaw rate| (the screening window of MIXED)
  UserKernels* user = nullptr;   // MPPI_MODEL_USER: the step's kernels instantiated at run time for the caller's functors
  // pinned host staging
  double* h_in = nullptr;      // x0[3], goal[3]
  DynState* h_out = nullptr;
  // zero-copy result hand-over of mppi_step: mapped pinned block written by the finalize phase, published by seq
  HostWire* h_res = nullptr;     // mapped pinned block the finalize phase writes (flag-in-data words)
  HostWire* d_res = nullptr;     // device alias of h_res
  HostResult res{};              // the last step's result, decoded by wait_result
  unsigned long long seq = 0;
  bool wait_block = false;       // MPPI_B200_WAIT=block: cudaStreamSynchronize instead of polling h_res->seq
  // graphs
  cudaGraphExec_t g_loop = nullptr;   // device-resident closed loop (mppi_bench)
  bool dirty = true;
  // host mirrors
  double goal[3] = {0, 0, 0};
  double last_x0[3] = {0, 0, 0}, last_goal[3] = {0, 0, 0};
  std::vector<double> U_prev;
  int last_capture_kind = -1;
  bool local_pending = false;
  mppi_timing last{};
};

static cudaError_t memcpy_on(mppi_engine* e, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  cudaError_t ce = cudaMemcpyAsync(dst, src, bytes, kind, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  return ce;
}

static int kind_of(int precision) {
  return precision == MPPI_PRECISION_F32 ? ROLLOUT_F32_SOFTMIN
                                         : (precision == MPPI_PRECISION_F64 ? ROLLOUT_F64_SOFTMIN : ROLLOUT_F32_SCREEN);
}

// ------------------------------------------------------------------------------------------------
extern "C" const char* mppi_last_error(void) { return g_err; }
extern "C" const char* mppi_version(void) { return "mppi_b200 0.1 (sm_100a)"; }
extern "C" int32_t mppi_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" mppi_status mppi_default_params(mppi_params* p) {
  if (!p) return MPPI_ERR_INVALID;
  memset(p, 0, sizeof(*p));
  p->struct_size = sizeof(mppi_params);
  p->abi_version = MPPI_B200_ABI_VERSION;
  p->K = 10;                                       // control/src/mppi:62
  p->T = 100;                                      // control/src/mppi:62
  p->model = MPPI_MODEL_DIFF_DRIVE;                // model=rk4, control/src/mppi:62
  p->weighting = MPPI_WEIGHT_COST_TO_GO;
  p->precision = MPPI_PRECISION_MIXED;
  p->device = 0;
  p->dt = 0.0;                                     // -> 1/T, control/src/mppi:67
  p->q[0] = 1e3; p->q[1] = 1e3; p->q[2] = 0.0;     // control/src/mppi:69
  p->r[0] = 1.0; p->r[3] = 1.0;                    // control/src/mppi:71
  p->p1[0] = p->p1[1] = p->p1[2] = 1e3;            // control/src/mppi:73
  p->sig[0] = 0.9; p->sig[3] = 0.9;                // control/src/mppi:88
  p->noise_std[0] = p->noise_std[1] = 0.9;         // sig[0,0], control/src/mppi:144-146
  p->lambda = 1e-3;                                // control/src/mppi:89
  p->u_max[0] = p->u_max[1] = 6.35492;             // WHEEL_VEL_MAX, control/src/mppi:18
  p->wheel_radius = 0.033;                         // control/src/mppi:19
  p->wheel_base = 0.16;                            // control/src/mppi:20
  p->eps_floor = 1e-8;                             // control/src/mppi:193
  p->seed = 0;                                     // np.random.seed(0), control/src/mppi:15
  p->k_offset = 0;
  p->k_total = 0;
  p->world_size = 1;
  p->rank = 0;
  p->stream = nullptr;
  p->refine_margin = 0.0;
  return MPPI_OK;
}

// The SCREEN window = support + head-room (common.cuh: screen_window).  Support: e^-40 ~ 4e-18 relative weight is far
// below the 1e-8 floor.  Head-room: the fp32 error of the screened cost-to-go grows with the magnitude of the costs being
// summed, so it is scaled with a bound Vmax of |V| in delta form over the horizon (distance to the goal enters through
// d (d + 2a)): head = max(2e-3, 4 * 2^-23 * Vmax) -- 5..8x the measured max |V32 - V64| of all BASELINE configs (profiles/),
// and the reduce kernel redoes the step in fp64 when the deviation it can observe exceeds half of it.
static void set_window(mppi_engine* e, double lam) {
  StaticParams& sp = e->sp;
  if (e->p.refine_margin > 0) {   // caller's static window
    sp.margin = e->p.refine_margin;
    sp.head_scale = sp.head_min = sp.vm_c0 = sp.vm_ca = sp.vm_cth = 0.0;
    return;
  }
  sp.margin = 40.0 * lam;
  sp.head_scale = 4.0;
  sp.head_min = 2e-3;
  const double horizon = sp.T * sp.dt;
  double v, w;   // largest forward speed / yaw rate of a clipped control
  if (sp.model == MPPI_MODEL_DIFF_DRIVE) {
    v = 0.5 * sp.wheel_r * (sp.u_max[0] + sp.u_max[1]);
    w = sp.wheel_r / sp.wheel_L * (sp.u_max[0] + sp.u_max[1]);
  } else if (sp.model == MPPI_MODEL_UNICYCLE_EULER) {
    v = sp.u_max[0];
    w = sp.u_max[1];
  } else if (sp.model == MPPI_MODEL_BICYCLE) {
    v = sp.u_max[0];
    w = sp.u_max[0] * std::tan(std::fmin(sp.u_max[1], 1.55)) / sp.wheel_L;
  } else {   // MPPI_MODEL_USER: a kinematic functor states its bounds (an ODE functor has no fp32 screen; the window is not used)
    v = e->user_speed_max;
    w = e->user_yaw_max;
  }
  const double D = v * horizon, Th = w * horizon;
  const double cA = sp.T * 0.5 * std::fmax(sp.q[0], sp.q[1]) + std::fmax(sp.p1[0], sp.p1[1]);
  const double cTh = sp.T * 0.5 * sp.q[2] + sp.p1[2];
  sp.vm_c0 = cA * D * D + cTh * Th * Th + std::fabs(sp.w_obs) * sp.T;
  sp.vm_ca = 2.0 * cA * D;
  sp.vm_cth = 2.0 * cTh * Th;
}

static void free_partials(mppi_engine* e) {
  cudaFree(e->d_part);
  cudaFree(e->d_epart);
  cudaFree(e->d_cand_meta);
  cudaFree(e->d_cand);
  e->d_part = nullptr;
  e->d_epart = nullptr;
  e->d_cand_meta = nullptr;
  e->d_cand = nullptr;
  e->part_capacity_ctas = 0;
}

static mppi_status drop_graphs(mppi_engine* e) {
  if (e->g_loop) cudaGraphExecDestroy(e->g_loop);
  e->g_loop = nullptr;
  e->dirty = true;
  return MPPI_OK;
}

// (re)compute launch configurations for the three rollout families and size the partial buffers
// largest |dt * yaw rate| any admissible (clipped) control can produce
static double max_yaw_increment(const StaticParams& sp) {
  double w;
  if (sp.model == MPPI_MODEL_USER) return 1e9;   // unknown: only the general code path applies
  if (sp.model == MPPI_MODEL_DIFF_DRIVE)
    w = sp.wheel_r / sp.wheel_L * (sp.u_max[0] + sp.u_max[1]);
  else if (sp.model == MPPI_MODEL_UNICYCLE_EULER)
    w = sp.u_max[1];
  else
    w = sp.u_max[0] * std::tan(std::fmin(sp.u_max[1], 1.55)) / sp.wheel_L;
  return sp.dt * w;
}

static bool try_configure(mppi_engine* e, int gin, size_t* max_ctas) {
  StaticParams& sp = e->sp;
  const bool has_grid = sp.has_grid != 0;
  // FAST kernels: in-register Philox noise and yaw increments small enough for the branch-free step;
  // LEAN (fp32 families only): additionally Q[2] == 0, Q[0] == Q[1] > 0 and |dt * yaw rate| <= 1/8 (rollout_lean_kernel.cuh).
  // MPPI_B200_BLOCK=<64|128> forces one tile shape, MPPI_B200_VARIANT=<general|fast|lean> caps the code path
  // (experiments / tests of the fallback paths).
  const double yaw_inc = max_yaw_increment(sp);
  int variant_max = ROLLOUT_LEAN;
  if (const char* envv = getenv("MPPI_B200_VARIANT")) {
    if (!strcmp(envv, "general")) variant_max = ROLLOUT_GENERAL;
    else if (!strcmp(envv, "fast")) variant_max = ROLLOUT_FAST;
  }
  const bool fast = !sp.noise_external && yaw_inc <= 0.78 && variant_max >= ROLLOUT_FAST;
  const bool lean = fast && sp.q[2] == 0.0 && sp.q[0] == sp.q[1] && sp.q[0] > 0.0 && sp.u_max[0] > 0.0 && sp.u_max[1] > 0.0 && (sp.model != MPPI_MODEL_BICYCLE || sp.u_max[1] <= 0.785) && yaw_inc <= (sp.model == MPPI_MODEL_UNICYCLE_EULER ? 0.5 : 1.0) * kLeanMaxYawInc && variant_max >= ROLLOUT_LEAN;
  const char* envb = getenv("MPPI_B200_BLOCK");
  *max_ctas = 0;
  for (int kind = 0; kind < 3; ++kind) {
    KindCfg best;
    double best_cost = 1e300;
    if (e->user) {   // run-time instantiation: general code path, tiles of 64; the screen family for kinematic functors only
      if (kind != ROLLOUT_F32_SCREEN || e->user_kind == 1) {
        KindCfg c;
        c.variant = ROLLOUT_GENERAL;
        c.block = 64;
        c.ntiles = (sp.K + 63) / 64;
        c.smem = rollout_smem(kind, sp.T, 64, ROLLOUT_GENERAL, gin);
        if (c.smem <= 227 * 1024 &&
            user_rollout_prepare(e->user, kind, has_grid, c.smem, &c.ctas_per_sm, &c.regs) == cudaSuccess &&
            c.ctas_per_sm >= 1) {
          const long long resident = (long long)e->num_sms * c.ctas_per_sm;
          c.grid = (int)((c.ntiles < resident) ? c.ntiles : resident);
          c.nparts = c.grid;
          c.ready = true;
          best = c;
        } else {
          cudaGetLastError();
        }
      }
      e->cfg[kind] = best;
      if (best.ready && (size_t)best.nparts > *max_ctas) *max_ctas = best.nparts;
      continue;
    }
    const int variant = !fast ? ROLLOUT_GENERAL : ((lean && kind != ROLLOUT_F64_SOFTMIN) ? ROLLOUT_LEAN : ROLLOUT_FAST);
    int shapes[3], nshapes = 0;
    if (variant == ROLLOUT_GENERAL) {
      shapes[nshapes++] = 64;
    } else if (envb && (atoi(envb) == 64 || atoi(envb) == 128 || (atoi(envb) == 512 && variant == ROLLOUT_LEAN))) {
      shapes[nshapes++] = atoi(envb);
    } else {
      shapes[nshapes++] = 64;
      shapes[nshapes++] = 128;
      if (variant == ROLLOUT_LEAN) shapes[nshapes++] = 512;
    }
    for (int si = 0; si < nshapes; ++si) {
      const int block = shapes[si];
      KindCfg c;
      c.variant = variant;
      c.block = block;
      const bool sm_wide = block == 512;   // rollout_lean_sm_kernel: 7 tiles of 64 per CTA, at most one CTA per SM
      const int rollouts = sm_wide ? 64 * 7 : block;
      c.ntiles = (sp.K + rollouts - 1) / rollouts;
      if (sm_wide && (c.ntiles > e->num_sms || sp.T < 12)) continue;
      c.smem = rollout_smem(kind, sp.T, block, variant, gin);
      if (c.smem > 227 * 1024) continue;
      cudaError_t ce = rollout_prepare(kind, sp.model, has_grid, block, variant, c.smem, &c.ctas_per_sm, &c.regs);
      if (ce != cudaSuccess || c.ctas_per_sm < 1) {
        cudaGetLastError();
        continue;
      }
      // The T-step loop is throughput (issue) bound -- measured 123 cycles per warp-step and scheduler at 2 warps per
      // scheduler, 136 at 3.5 -- so an SM's time is (its tiles) x (warps per tile) / 4 schedulers, TIMES two penalties that
      // depend on how many CTAs are made resident per SM (a choice, <= what fits; tiles are assigned statically):
      //   imbalance: r CTAs of w warps put ceil(r w / 4) warps on the busiest scheduler and every tile runs at its pace
      //              (5 CTAs of 2 warps = 3,3,2,2: K = 1048576, T = 128 measured 510 us that way, 448 us expected balanced);
      //   latency:   below ~1.6 warps per scheduler the issue slots cannot be kept busy.
      // The SM-wide kernel cuts its seventh tile in time and carries 3.5 on every scheduler.
      const int wpb = block / 32;
      const int tiles_per_sm = (c.ntiles + e->num_sms - 1) / e->num_sms;
      int best_r = c.ctas_per_sm;
      double cost = 1e300;
      if (sm_wide) {
        best_r = 1;
        cost = 3.5;
      } else {
        for (int r = 1; r <= c.ctas_per_sm; ++r) {
          const double W = 0.25 * r * wpb;
          const double imb = std::ceil(W) / W, lat = W < 1.6 ? 1.6 / W : 1.0;
          const double cr = tiles_per_sm * 0.25 * wpb * imb * lat;
          if (cr <= cost) {   // ties: more resident warps
            cost = cr;
            best_r = r;
          }
        }
      }
      c.ctas_per_sm = best_r;
      const long long resident = (long long)e->num_sms * c.ctas_per_sm;
      c.grid = (int)((c.ntiles < resident) ? c.ntiles : resident);
      c.nparts = sm_wide ? (sp.K + 63) / 64 : c.grid;
      // ties between shapes: less thread-work on the busiest SM, then the larger tile
      cost = cost * 1e6 + (double)tiles_per_sm * rollouts;
      c.ready = true;
      if (cost < best_cost || (cost == best_cost && block > best.block)) {
        best = c;
        best_cost = cost;
      }
    }
    e->cfg[kind] = best;
    if (best.ready && (size_t)best.nparts > *max_ctas) *max_ctas = best.nparts;
  }
  // the fp64 family is needed by precision F64 itself and by MIXED (redo of an overflowing step); an F32 engine only needs
  // it for the debug entry point mppi_cost_to_go, which then reports MPPI_ERR_UNSUPPORTED by itself
  return e->cfg[kind_of(e->p.precision)].ready && (e->p.precision == MPPI_PRECISION_F32 || e->cfg[ROLLOUT_F64_SOFTMIN].ready);
}

static mppi_status configure(mppi_engine* e) {
  StaticParams& sp = e->sp;
  size_t max_ctas = 0;
  int gin = (sp.has_grid && sp.grid_bytes_padded <= 96 * 1024) ? sp.grid_bytes_padded : 0;
  bool ok = try_configure(e, gin, &max_ctas);
  if (!ok && gin) {   // the grid does not fit beside the cost tile: read it through L1/L2 instead
    gin = 0;
    ok = try_configure(e, 0, &max_ctas);
  }
  if (!ok) {
    set_err("no launch configuration fits: the cost tile of T=%d steps needs too much shared memory for precision %s "
            "(fp64 rollouts: T <= ~400; fp32 rollouts: T <= ~800)", sp.T,
            e->p.precision == MPPI_PRECISION_F32 ? "F32" : (e->p.precision == MPPI_PRECISION_F64 ? "F64" : "MIXED"));
    return MPPI_ERR_UNSUPPORTED;
  }
  sp.grid_in_smem = gin > 0 ? 1 : 0;
  if (max_ctas > e->part_capacity_ctas) {
    free_partials(e);
    const size_t n = (size_t)sp.T * max_ctas;
    CK(cudaMalloc(&e->d_part, n * sizeof(double4)));
    CK(cudaMalloc(&e->d_epart, n * 2 * sizeof(double)));
    CK(cudaMalloc(&e->d_cand_meta, n * sizeof(float4)));
    CK(cudaMalloc(&e->d_cand, n * kMaxCand * sizeof(uint2)));
    e->part_capacity_ctas = max_ctas;
  }
  return drop_graphs(e);
}

static void savgol_basis(int T, double& a, double& b, double inv_norm[4]) {
  // Gram (discrete orthogonal) cubic basis on z = -h..h, window w = T-1 (control/src/mppi:202):
  // p0 = 1, p1 = z, p2 = z^2 - a, p3 = z^3 - b z;  projection coefficient i = sum_j p_i(z_j) u_j / sum_j p_i(z_j)^2
  const int W = T - 1, h = W / 2;
  long double s2 = 0, s4 = 0;
  for (int j = 0; j < W; ++j) {
    long double z = j - h;
    s2 += z * z;
    s4 += z * z * z * z;
  }
  const long double la = s2 / W, lb = s4 / s2;
  long double n[4] = {0, 0, 0, 0};
  for (int j = 0; j < W; ++j) {
    long double z = j - h;
    long double pz[4] = {1.0L, z, z * z - la, z * z * z - lb * z};
    for (int i = 0; i < 4; ++i) n[i] += pz[i] * pz[i];
  }
  for (int i = 0; i < 4; ++i) inv_norm[i] = (double)(1.0L / n[i]);
  a = (double)la;
  b = (double)lb;
}

static mppi_status upload_dyn_sampling(mppi_engine* e, const double sig[4], double lam, const double nstd[2]) {
  DynState tmp;
  COPY_SYNC(e, &tmp, e->d_dyn, sizeof(DynState), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 4; ++i) tmp.sig[i] = sig[i];
  tmp.lam = lam;
  tmp.noise_std[0] = nstd[0];
  tmp.noise_std[1] = nstd[1];
  tmp.neg_inv_lam_f = (float)(-1.0 / lam);
  tmp.noise_std_f[0] = (float)nstd[0];
  tmp.noise_std_f[1] = (float)nstd[1];
  COPY_SYNC(e, e->d_dyn, &tmp, sizeof(DynState), cudaMemcpyHostToDevice);
  return MPPI_OK;
}

static mppi_status prep_nominal(mppi_engine* e) {
  CK(prep_nominal_launch(e->stream, e->d_dyn, e->sp, e->d_Umaster, e->d_nomF, e->d_nomD));
  CK(cudaStreamSynchronize(e->stream));
  return MPPI_OK;
}

static mppi_status create_impl(const mppi_params* pin, const mppi_user_model* um, mppi_handle* out);

extern "C" mppi_status mppi_create(const mppi_params* pin, mppi_handle* out) {
  if (pin && pin->model == MPPI_MODEL_USER) {
    set_err("model MPPI_MODEL_USER needs its functor source: use mppi_create_user");
    return MPPI_ERR_INVALID;
  }
  return create_impl(pin, nullptr, out);
}

extern "C" mppi_status mppi_create_user(const mppi_params* pin, const mppi_user_model* um, mppi_handle* out) {
  if (!um) {
    set_err("null mppi_user_model");
    return MPPI_ERR_INVALID;
  }
  return create_impl(pin, um, out);
}

extern "C" mppi_status mppi_check_user_model(const mppi_user_model* um) {
  std::string log;
  size_t bytes = 0;
  const mppi_status s = user_model_check(um, &log, &bytes);
  if (s != MPPI_OK)
    set_err("%s", log.c_str());
  else
    set_err("user model compiled for sm_100a: %zu bytes of cubin", bytes);
  return s;
}

static mppi_status create_impl(const mppi_params* pin, const mppi_user_model* um, mppi_handle* out) {
  if (!pin || !out) {
    set_err("null argument");
    return MPPI_ERR_INVALID;
  }
  *out = nullptr;
  if (pin->struct_size != sizeof(mppi_params) || pin->abi_version != MPPI_B200_ABI_VERSION) {
    set_err("mppi_params ABI mismatch (size %u vs %zu, version %u vs %d)", pin->struct_size, sizeof(mppi_params),
            pin->abi_version, MPPI_B200_ABI_VERSION);
    return MPPI_ERR_INVALID;
  }
  mppi_params p = *pin;
  if (um) p.model = MPPI_MODEL_USER;
  if (p.K < 1 || p.T < 6 || (p.T & 1) || p.T > 1024) {
    set_err("need K >= 1 and even 6 <= T <= 1024 (savgol window T-1 must be odd, control/src/mppi:202); got K=%d T=%d", p.K, p.T);
    return MPPI_ERR_INVALID;
  }
  if (p.model < 0 || p.model > 3 || p.weighting < 0 || p.weighting > 1 || p.precision < 0 || p.precision > 2) {
    set_err("bad model/weighting/precision enum");
    return MPPI_ERR_INVALID;
  }
  if (um && p.precision == MPPI_PRECISION_MIXED) {
    if (um->kind != 1 || um->has_cost) {
      set_err("precision MIXED needs a KINEMATIC functor (mppi_user_model.kind 1) and the built-in cost: its fp32 screen (delta-form "
              "cost, time-parallel fp64 re-evaluation) relies on xdot = s(u) cos(theta), ydot = s(u) sin(theta), thetadot = w(u); "
              "use precision F64 or F32");
      return MPPI_ERR_UNSUPPORTED;
    }
    if (!(um->speed_max > 0) || !(um->yaw_rate_max >= 0) || !std::isfinite(um->speed_max) || !std::isfinite(um->yaw_rate_max)) {
      set_err("precision MIXED with a kinematic functor needs speed_max > 0 and yaw_rate_max >= 0 (bounds over the clipped controls)");
      return MPPI_ERR_INVALID;
    }
  }
  if (p.precision == MPPI_PRECISION_MIXED && p.T > 400) {
    set_err("precision MIXED supports T <= 400 (fp64 refinement scratch is 56*T bytes per warp); use F64 or F32");
    return MPPI_ERR_UNSUPPORTED;
  }
  if (!(p.lambda > 0) || !(p.wheel_base > 0) || !(p.u_max[0] > 0) || !(p.u_max[1] > 0)) {
    set_err("lambda, wheel_base and u_max must be positive");
    return MPPI_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
    cudaGetLastError();
    set_err("no CUDA device: the MPPI engine has no CPU fallback");
    return MPPI_ERR_NO_DEVICE;
  }
  if (p.device < 0 || p.device >= ndev) {
    set_err("device %d out of range (have %d)", p.device, ndev);
    return MPPI_ERR_INVALID;
  }
  CK(cudaSetDevice(p.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, p.device));
  if (prop.major < 10) {
    set_err("device %d is sm_%d%d; this library is built for sm_100a only", p.device, prop.major, prop.minor);
    return MPPI_ERR_UNSUPPORTED;
  }
  if (p.dt <= 0) p.dt = 1.0 / (double)p.T;
  if (p.k_total <= 0) p.k_total = p.K;
  if (p.world_size <= 0) p.world_size = 1;

  mppi_engine* e = new mppi_engine();
  e->p = p;
  e->dev = p.device;
  e->num_sms = prop.multiProcessorCount;
  if (p.stream) {
    e->stream = (cudaStream_t)p.stream;
    e->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_err("cudaStreamCreate failed");
      delete e;
      return MPPI_ERR_CUDA;
    }
    e->own_stream = true;
  }
  StaticParams& sp = e->sp;
  memset(&sp, 0, sizeof(sp));
  sp.K = p.K;
  sp.T = p.T;
  sp.k_offset = p.k_offset;
  sp.k_total = p.k_total;
  sp.model = p.model;
  sp.weighting = p.weighting;
  sp.world = p.world_size;
  sp.dt = p.dt;
  for (int i = 0; i < 3; ++i) {
    sp.q[i] = p.q[i];
    sp.p1[i] = p.p1[i];
  }
  sp.u_max[0] = p.u_max[0];
  sp.u_max[1] = p.u_max[1];
  sp.wheel_r = p.wheel_radius;
  sp.wheel_L = p.wheel_base;
  sp.eps_floor = p.eps_floor;
  sp.seed = p.seed;
  sp.g_inv_res = 1.0;
  if (um) {
    e->user_kind = um->kind;
    e->user_speed_max = um->speed_max;
    e->user_yaw_max = um->yaw_rate_max;
  }
  set_window(e, p.lambda);

  const int T = p.T;
  auto fail = [&](mppi_status s) {
    mppi_destroy(e);
    return s;
  };
#define CKF(call)                                                                                \
  do {                                                                                           \
    cudaError_t _e = (call);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));             \
      return fail(MPPI_ERR_CUDA);                                                                \
    }                                                                                            \
  } while (0)
  CKF(cudaMalloc(&e->d_dyn, sizeof(DynState)));
  CKF(cudaMalloc(&e->d_Umaster, 2 * T * sizeof(double)));
  CKF(cudaMalloc(&e->d_Ulast, 2 * T * sizeof(double)));
  CKF(cudaMalloc(&e->d_Utmp, 4 * T * sizeof(double)));
  CKF(cudaMalloc(&e->d_nomD, 4 * T * sizeof(double)));
  CKF(cudaMalloc(&e->d_nomF, 8 * T * sizeof(float)));   // planar [4][T] + interleaved float4[T] (LEAN)
  CKF(cudaMalloc(&e->d_record, (size_t)T * kRecordStride * sizeof(double)));
  CKF(cudaMalloc(&e->d_record_tmp, (size_t)T * kRecordStride * sizeof(double)));
  CKF(cudaMalloc(&e->d_gather, (size_t)p.world_size * T * kRecordStride * sizeof(double)));
  e->ll_rows_uint2 = (size_t)2 * p.world_size * T * kRowWords;
  e->ll_bytes = e->ll_rows_uint2 * sizeof(uint2) + (size_t)2 * kMaxFusedWorld * sizeof(unsigned int);   // rows + rendezvous flags
  CKF(cudaMalloc(&e->d_ll, e->ll_bytes));
  CKF(cudaMemset(e->d_ll, 0, e->ll_bytes));       // flag 0 never matches an epoch + 1
  if (p.world_size == 1) e->ll_peers[0] = e->d_ll;
  CKF(cudaMalloc(&e->d_ll2, (size_t)2 * T * kRow2Words * sizeof(uint2)));
  CKF(cudaMemset(e->d_ll2, 0, (size_t)2 * T * kRow2Words * sizeof(uint2)));
  CKF(cudaMallocHost(&e->h_in, 6 * sizeof(double)));
  CKF(cudaMallocHost(&e->h_out, sizeof(DynState)));
  CKF(cudaHostAlloc(&e->h_res, sizeof(HostWire), cudaHostAllocMapped));
  memset(e->h_res, 0, sizeof(HostWire));
  CKF(cudaHostGetDevicePointer((void**)&e->d_res, e->h_res, 0));
  if (const char* w = getenv("MPPI_B200_WAIT")) e->wait_block = !strcmp(w, "block");
  if (const char* w = getenv("MPPI_B200_SPLIT")) e->lean_split = atoi(w);
  CKF(cudaMemset(e->d_Umaster, 0, 2 * T * sizeof(double)));     // uvec_init = zeros, control/src/mppi:65
  CKF(cudaMemset(e->d_Ulast, 0, 2 * T * sizeof(double)));
  CKF(cudaMemset(e->d_record, 0, (size_t)T * kRecordStride * sizeof(double)));
  CKF(cudaMemset(e->d_gather, 0, (size_t)p.world_size * T * kRecordStride * sizeof(double)));
  savgol_basis(T, e->sg_a, e->sg_b, e->sg_inv_norm);
  {
    DynState d;
    memset(&d, 0, sizeof(d));
    d.lam = p.lambda;
    for (int i = 0; i < 4; ++i) {
      d.sig[i] = p.sig[i];
      d.R[i] = p.r[i];
    }
    d.noise_std[0] = p.noise_std[0];
    d.noise_std[1] = p.noise_std[1];
    d.neg_inv_lam_f = (float)(-1.0 / p.lambda);
    d.noise_std_f[0] = (float)p.noise_std[0];
    d.noise_std_f[1] = (float)p.noise_std[1];
    CKF(memcpy_on(e, e->d_dyn, &d, sizeof(d), cudaMemcpyHostToDevice));
  }
  e->U_prev.assign(2 * T, 0.0);
  if (um) {   // compile the caller's functors into their own instantiation of the step's kernels
    std::string log;
    const mppi_status us = user_kernels_build(um, &e->user, &log);
    if (us != MPPI_OK) {
      set_err("%s", log.c_str());
      return fail(us);
    }
    e->user_has_cost = um->has_cost != 0;
  }
  mppi_status s = configure(e);
  if (s != MPPI_OK) return fail(s);
  s = prep_nominal(e);
  if (s != MPPI_OK) return fail(s);
  *out = e;
  return MPPI_OK;
}

extern "C" mppi_status mppi_destroy(mppi_handle e) {
  if (!e) return MPPI_OK;
  cudaSetDevice(e->dev);
  if (e->stream) cudaStreamSynchronize(e->stream);
  drop_graphs(e);
  free_partials(e);
  user_kernels_free(e->user);
  cudaFree(e->d_dyn);
  cudaFree(e->d_Umaster);
  cudaFree(e->d_Ulast);
  cudaFree(e->d_Utmp);
  cudaFree(e->d_nomD);
  cudaFree(e->d_nomF);
  cudaFree(e->d_record);
  cudaFree(e->d_record_tmp);
  cudaFree(e->d_gather);
  cudaFree(e->d_debug_ts);
  cudaFree(e->d_debug_rts);
  for (void* q : e->p2p_opened) cudaIpcCloseMemHandle(q);
  cudaFree(e->d_ll);
  cudaFree(e->d_ll2);
  cudaFree(e->d_grid);
  if (e->h_grid_stage) cudaFreeHost(e->h_grid_stage);
  if (e->grid_stage_ev) cudaEventDestroy(e->grid_stage_ev);
  cudaFree(e->d_eps_ext);
  cudaFree(e->d_vcap);
  cudaFree(e->d_flush);
  if (e->h_in) cudaFreeHost(e->h_in);
  if (e->h_out) cudaFreeHost(e->h_out);
  if (e->h_res) cudaFreeHost(e->h_res);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return MPPI_OK;
}

#define ENTER(e)                                  \
  if (!(e)) {                                     \
    set_err("null handle");                       \
    return MPPI_ERR_INVALID;                      \
  }                                               \
  {                                               \
    int cur_ = -1;                                \
    if (cudaGetDevice(&cur_) != cudaSuccess || cur_ != (e)->dev) CK(cudaSetDevice((e)->dev)); \
  }

extern "C" mppi_status mppi_reset(mppi_handle e) {
  ENTER(e);
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemsetAsync(e->d_Umaster, 0, 2 * e->sp.T * sizeof(double), e->stream));   // control/src/mppi:81
  e->local_pending = false;
  e->f64_holdoff = 0;
  e->f64_backoff = 8;
  return prep_nominal(e);
}

extern "C" mppi_status mppi_set_goal(mppi_handle e, const double goal[3]) {
  ENTER(e);
  if (!goal) return MPPI_ERR_INVALID;
  for (int i = 0; i < 3; ++i) e->goal[i] = goal[i];
  return MPPI_OK;
}

extern "C" mppi_status mppi_set_sampling(mppi_handle e, const double sig[4], double lambda) {
  ENTER(e);
  if (!sig || !(lambda > 0)) {
    set_err("sig must be non-null and lambda > 0");
    return MPPI_ERR_INVALID;
  }
  CK(cudaStreamSynchronize(e->stream));
  const double nstd[2] = {sig[0], sig[0]};   // reference: normal(0, sig[0,0]) for both rows, control/src/mppi:144-146
  CKS(upload_dyn_sampling(e, sig, lambda, nstd));
  for (int i = 0; i < 4; ++i) e->p.sig[i] = sig[i];
  e->p.lambda = lambda;
  e->p.noise_std[0] = e->p.noise_std[1] = sig[0];
  const double m = e->sp.margin;
  set_window(e, lambda);
  if (m != e->sp.margin) drop_graphs(e);
  return prep_nominal(e);
}

extern "C" mppi_status mppi_set_noise_std(mppi_handle e, const double nstd[2]) {
  ENTER(e);
  if (!nstd) return MPPI_ERR_INVALID;
  CK(cudaStreamSynchronize(e->stream));
  CKS(upload_dyn_sampling(e, e->p.sig, e->p.lambda, nstd));
  e->p.noise_std[0] = nstd[0];
  e->p.noise_std[1] = nstd[1];
  return prep_nominal(e);   // the LEAN nominal block carries std * g
}

extern "C" mppi_status mppi_get_nominal(mppi_handle e, double* U) {
  ENTER(e);
  if (!U) return MPPI_ERR_INVALID;
  COPY_SYNC(e, U, e->d_Umaster, 2 * e->sp.T * sizeof(double), cudaMemcpyDeviceToHost);
  return MPPI_OK;
}

extern "C" mppi_status mppi_set_nominal(mppi_handle e, const double* U) {
  ENTER(e);
  if (!U) return MPPI_ERR_INVALID;
  COPY_SYNC(e, e->d_Umaster, U, 2 * e->sp.T * sizeof(double), cudaMemcpyHostToDevice);
  return prep_nominal(e);
}

extern "C" mppi_status mppi_get_last_update(mppi_handle e, double* U) {
  ENTER(e);
  if (!U) return MPPI_ERR_INVALID;
  COPY_SYNC(e, U, e->d_Ulast, 2 * e->sp.T * sizeof(double), cudaMemcpyDeviceToHost);
  return MPPI_OK;
}

extern "C" mppi_status mppi_set_grid(mppi_handle e, const int8_t* cells, int32_t W, int32_t H, double res, double x_min,
                                     double y_min, double w_obs) {
  ENTER(e);
  if (!cells || W < 1 || H < 1 || !(res > 0) || (long long)W * H > (1LL << 30)) {
    set_err("bad grid");
    return MPPI_ERR_INVALID;
  }
  CK(cudaStreamSynchronize(e->stream));
  const size_t n = (size_t)W * H, padded = (n + 15) & ~(size_t)15;
  // the new grid is complete on the device before the old one is released: a failure leaves the engine as it was
  signed char* fresh = nullptr;
  CK(cudaMalloc(&fresh, padded));
  cudaError_t ce = cudaMemsetAsync(fresh, 100, padded, e->stream);
  if (ce == cudaSuccess) ce = memcpy_on(e, fresh, cells, n, cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    cudaFree(fresh);
    set_err("mppi_set_grid: upload failed: %s", cudaGetErrorString(ce));
    return MPPI_ERR_CUDA;
  }
  cudaFree(e->d_grid);
  e->d_grid = fresh;
  StaticParams& sp = e->sp;
  sp.has_grid = 1;
  sp.gW = W;
  sp.gH = H;
  sp.grid_bytes_padded = (int)padded;
  sp.g_inv_res = 1.0 / res;
  sp.g_x0 = x_min;
  sp.g_y0 = y_min;
  sp.w_obs = w_obs;
  set_window(e, e->p.lambda);
  return configure(e);
}

// Patch a rectangle of the resident grid (the incrementally revealed map of the planners' simulated sensor,
// map/src/map/grid.cpp:155-199: Grid::update_grid / fake_occupancy_grid; or any nav_msgs/OccupancyGrid update): rows
// y0 .. y0+h-1, columns x0 .. x0+w-1, `cells` row-major (h, w).  Asynchronous: the patch is staged in pinned host memory and
// copied on the engine's stream, so it is ordered before the next step's kernels and the call does not wait for the device;
// no reallocation, no reconfiguration (the rollout kernel re-reads the grid from HBM -- through TMA into shared memory when
// it fits -- at every launch).
extern "C" mppi_status mppi_update_grid(mppi_handle e, const int8_t* cells, int32_t x0, int32_t y0, int32_t w, int32_t h) {
  ENTER(e);
  const StaticParams& sp = e->sp;
  if (!sp.has_grid || !e->d_grid) {
    set_err("mppi_update_grid: no grid is resident (call mppi_set_grid first)");
    return MPPI_ERR_STATE;
  }
  if (!cells || w < 1 || h < 1 || x0 < 0 || y0 < 0 || (long long)x0 + w > sp.gW || (long long)y0 + h > sp.gH) {
    set_err("mppi_update_grid: patch [%d,%d)+(%d x %d) outside the %d x %d grid", x0, y0, w, h, sp.gW, sp.gH);
    return MPPI_ERR_INVALID;
  }
  const size_t bytes = (size_t)w * h;
  if (e->grid_stage_ev) CK(cudaEventSynchronize(e->grid_stage_ev));   // the previous patch has left the staging buffer
  if (bytes > e->grid_stage_cap) {
    if (e->h_grid_stage) cudaFreeHost(e->h_grid_stage);
    e->h_grid_stage = nullptr;
    e->grid_stage_cap = 0;
    CK(cudaMallocHost(&e->h_grid_stage, bytes < 4096 ? 4096 : bytes));
    e->grid_stage_cap = bytes < 4096 ? 4096 : bytes;
  }
  if (!e->grid_stage_ev) CK(cudaEventCreateWithFlags(&e->grid_stage_ev, cudaEventDisableTiming));
  memcpy(e->h_grid_stage, cells, bytes);
  CK(cudaMemcpy2DAsync(e->d_grid + (size_t)y0 * sp.gW + x0, (size_t)sp.gW, e->h_grid_stage, (size_t)w, (size_t)w, (size_t)h,
                       cudaMemcpyHostToDevice, e->stream));
  CK(cudaEventRecord(e->grid_stage_ev, e->stream));
  return MPPI_OK;
}

extern "C" mppi_status mppi_clear_grid(mppi_handle e) {
  ENTER(e);
  CK(cudaStreamSynchronize(e->stream));
  e->sp.has_grid = 0;
  e->sp.grid_in_smem = 0;
  e->sp.w_obs = 0;
  set_window(e, e->p.lambda);
  return configure(e);
}

extern "C" mppi_status mppi_set_noise(mppi_handle e, const double* eps) {
  ENTER(e);
  if (!eps) return MPPI_ERR_INVALID;
  CK(cudaStreamSynchronize(e->stream));
  const size_t n = (size_t)e->sp.T * 2 * e->sp.K;
  if (!e->d_eps_ext) CK(cudaMalloc(&e->d_eps_ext, n * sizeof(double)));
  COPY_SYNC(e, e->d_eps_ext, eps, n * sizeof(double), cudaMemcpyHostToDevice);
  if (!e->sp.noise_external) {
    e->sp.noise_external = 1;
    return configure(e);     // replayed noise runs on the GENERAL kernels
  }
  return MPPI_OK;
}

extern "C" mppi_status mppi_use_philox(mppi_handle e, uint64_t seed) {
  ENTER(e);
  CK(cudaStreamSynchronize(e->stream));
  e->sp.noise_external = 0;
  e->sp.seed = seed;
  unsigned int zero = 0;
  COPY_SYNC(e, &e->d_dyn->step, &zero, sizeof(zero), cudaMemcpyHostToDevice);
  return configure(e);
}

extern "C" mppi_status mppi_get_noise(mppi_handle e, double* eps) {
  ENTER(e);
  if (!eps) return MPPI_ERR_INVALID;
  CK(cudaStreamSynchronize(e->stream));
  const size_t n = (size_t)e->sp.T * 2 * e->sp.K;
  if (e->sp.noise_external) {
    COPY_SYNC(e, eps, e->d_eps_ext, n * sizeof(double), cudaMemcpyDeviceToHost);
    return MPPI_OK;
  }
  unsigned int step = 0;
  COPY_SYNC(e, &step, &e->d_dyn->step, sizeof(step), cudaMemcpyDeviceToHost);
  if (step == 0) {
    set_err("mppi_get_noise: no step has been run since the noise stream was (re)seeded");
    return MPPI_ERR_STATE;
  }
  double* tmp = nullptr;
  CK(cudaMalloc(&tmp, n * sizeof(double)));
  cudaError_t ce = noise_export_launch(e->stream, e->sp, e->d_dyn, step - 1, tmp);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(eps, tmp, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  cudaFree(tmp);
  CK(ce);
  return MPPI_OK;
}

extern "C" mppi_status mppi_set_capture(mppi_handle e, int32_t on) {
  ENTER(e);
  CK(cudaStreamSynchronize(e->stream));
  if (on && !e->d_vcap) CK(cudaMalloc(&e->d_vcap, (size_t)e->sp.T * e->sp.K * sizeof(double)));
  if ((on != 0) != (e->sp.capture != 0)) {
    e->sp.capture = on ? 1 : 0;
    drop_graphs(e);
  }
  return MPPI_OK;
}

// ---- the three-kernel pipeline ---------------------------------------------------------------------
struct KernelEvents {
  cudaEvent_t ev[4];
  bool on = false;
};

static FinalizeArgs make_fin(mppi_engine* e, bool closed_loop) {
  FinalizeArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.sp = e->sp;
  fa.dyn = e->d_dyn;
  fa.gather = (e->sp.world > 1) ? e->d_gather : e->d_record;
  fa.ll2_local = nullptr;
  fa.Umaster = e->d_Umaster;
  fa.Ulast = e->d_Ulast;
  fa.nomF = e->d_nomF;
  fa.nomD = e->d_nomD;
  fa.sg_a = e->sg_a;
  fa.sg_b = e->sg_b;
  for (int i = 0; i < 4; ++i) fa.sg_inv_norm[i] = e->sg_inv_norm[i];
  fa.mode = 0;
  fa.closed_loop = closed_loop ? 1 : 0;
  return fa;
}

enum FuseMode { FUSE_NONE = 0, FUSE_STEP = 1, FUSE_LOOP = 2 };

// rollout + reduce (+ row exchange and the finalizer block inside the reduce kernel when fuse != FUSE_NONE)
// `in` != nullptr: x0 / goal travel as kernel arguments and the result is published to e->h_res under e->seq
static mppi_status launch_local(mppi_engine* e, cudaStream_t st, int precision, int fuse, KernelEvents* kev,
                                const StepInput* in = nullptr) {
  const int kind = kind_of(precision);
  const KindCfg& c = e->cfg[kind];
  RolloutArgs ra;
  memset(&ra, 0, sizeof(ra));
  ra.sp = e->sp;
  ra.dyn = e->d_dyn;
  ra.nom = (kind == ROLLOUT_F64_SOFTMIN) ? (const void*)e->d_nomD
                                         : (const void*)(e->d_nomF + (c.variant == ROLLOUT_LEAN ? 4 * e->sp.T : 0));
  ra.grid = e->d_grid;
  ra.eps_ext = e->d_eps_ext;
  ra.part = e->d_part;
  ra.epart = e->d_epart;
  ra.cand_meta = e->d_cand_meta;
  ra.cand = e->d_cand;
  ra.vcap = e->d_vcap;
  ra.ntiles = (c.block == 512) ? c.nparts : c.ntiles;   // SM-wide kernel: the number of 64-rollout tiles
  if (in) ra.in = *in;
  ra.debug_ts = e->d_debug_rts;
  if (c.variant == ROLLOUT_LEAN) {
    const StaticParams& sp = e->sp;
    LeanStatic& ls = ra.lean;
    const double hq = 0.5 * sp.q[0], sq = std::sqrt(hq);
    const double um0 = sp.u_max[0], um1 = sp.u_max[1];
    double A0 = 0, A1 = 0, Ac = 0, G0 = 0, G1 = 0, Gc = 0, bk = 0;
    if (sp.model == MPPI_MODEL_DIFF_DRIVE) {          // u = u_max (2 s - 1); a = (dt r / 2L)(u1 - u0), g = (dt r / 12)(u0 + u1)
      const double ca = 0.5 * sp.dt * sp.wheel_r / sp.wheel_L, cg = sq * sp.dt * sp.wheel_r * 0.5 / 6.0;
      A0 = -2.0 * ca * um0;
      A1 = 2.0 * ca * um1;
      Ac = ca * (um0 - um1);
      G0 = 2.0 * cg * um0;
      G1 = 2.0 * cg * um1;
      Gc = -cg * (um0 + um1);
    } else if (sp.model == MPPI_MODEL_UNICYCLE_EULER) {   // a = dt u1 (full increment), g = dt u0
      A1 = 2.0 * sp.dt * um1;
      Ac = -sp.dt * um1;
      G0 = 2.0 * sq * sp.dt * um0;
      Gc = -sq * sp.dt * um0;
    } else {                                          // bicycle: a = (dt / 2L) v tan(delta), g = (dt / 6) v
      bk = 0.5 * sp.dt / sp.wheel_L;
      G0 = sq * sp.dt / 6.0;
    }
    ls.A0 = (float)A0;
    ls.A1 = (float)A1;
    ls.Ac = (float)Ac;
    ls.G0 = (float)G0;
    ls.G1 = (float)G1;
    ls.Gc = (float)Gc;
    ls.um0 = (float)um0;
    ls.um1 = (float)um1;
    ls.bk = (float)bk;
    ls.inv2um0 = (float)(0.5 / um0);
    ls.inv2um1 = (float)(0.5 / um1);
    ls.sq = (float)sq;
    ls.p1x = (float)(sp.p1[0] / hq);
    ls.p1y = (float)(sp.p1[1] / hq);
    ls.p1th = (float)sp.p1[2];
    ls.g_inv_res = (float)(sp.g_inv_res / sq);
    ls.w_obs_100 = (float)(sp.w_obs / 100.0);
    ls.split = e->lean_split > 0 ? e->lean_split : ((sp.T * 9 / 16 + 5) / 6) * 6;
    for (int i = 0; i < MPPI_PHILOX_ROUNDS; ++i) {
      ls.pkx[i] = (uint32_t)sp.seed + (uint32_t)i * 0x9E3779B9u;
      ls.pky[i] = (uint32_t)(sp.seed >> 32) + (uint32_t)i * 0xBB67AE85u;
    }
  }
  if (kev && kev->on) CK(cudaEventRecord(kev->ev[0], st));
  if (e->user)
    CK(user_rollout_launch(e->user, kind, e->sp.has_grid != 0, c.grid, c.smem, st, ra));
  else
    CK(rollout_launch(kind, e->sp.model, e->sp.has_grid != 0, c.block, c.variant, c.grid, c.smem, st, ra));
  e->tp_launch1 = std::chrono::steady_clock::now();
  if (kev && kev->on) CK(cudaEventRecord(kev->ev[1], st));
  ReduceArgs rd;
  memset(&rd, 0, sizeof(rd));
  rd.sp = e->sp;
  rd.fin = make_fin(e, fuse == FUSE_LOOP);
  if (in) {
    rd.fin.in = *in;
    if (fuse != FUSE_NONE) {
      rd.fin.host_res = e->d_res;
      rd.fin.seq = e->seq;
    }
  }
  rd.fused = fuse != FUSE_NONE ? 1 : 0;   // (callers have checked: world == 1, or the peers' row buffers are mapped)
  if (rd.fused) {
    rd.fin.ll2_local = e->d_ll2;
    rd.fin.gather = nullptr;
    rd.fin.debug_ts = e->d_debug_ts ? e->d_debug_ts + (size_t)e->sp.T * 8 : nullptr;
  }
  rd.rank = e->p.rank;
  rd.ll2_local = e->d_ll2;
  for (int g = 0; g < kMaxFusedWorld; ++g) rd.ll_peers[g] = e->ll_peers[g];
  rd.debug_ts = e->d_debug_ts;
  rd.part = e->d_part;
  rd.epart = e->d_epart;
  rd.cand_meta = e->d_cand_meta;
  rd.cand = e->d_cand;
  rd.nomD = e->d_nomD;
  rd.grid = e->d_grid;
  rd.eps_ext = e->d_eps_ext;
  rd.record = e->d_record;
  rd.nCTA = c.nparts;
  if (e->user && kind == ROLLOUT_F32_SCREEN)
    CK(user_reduce_screen_launch(e->user, e->sp.has_grid != 0, e->sp.T, st, rd));
  else if (e->user)
    CK(user_reduce_softmin_launch(e->user, kind == ROLLOUT_F64_SOFTMIN, e->sp.T, st, rd));
  else if (kind == ROLLOUT_F32_SCREEN)
    CK(reduce_screen_launch(e->sp.model, e->sp.has_grid != 0, e->sp.T, st, rd));
  else
    CK(reduce_softmin_launch(kind == ROLLOUT_F64_SOFTMIN, e->sp.T, st, rd));
  e->tp_launch2 = std::chrono::steady_clock::now();
  if (kev && kev->on) CK(cudaEventRecord(kev->ev[2], st));
  e->last_capture_kind = kind;
  return MPPI_OK;
}

// stand-alone finalize (sharded steps: after the exchange)
static mppi_status launch_finalize(mppi_engine* e, cudaStream_t st, bool closed_loop, KernelEvents* kev, bool publish = false) {
  FinalizeArgs fa = make_fin(e, closed_loop);
  if (publish) {
    fa.host_res = e->d_res;
    fa.seq = e->seq;
  }
  CK(e->user ? user_finalize_launch(e->user, st, fa) : finalize_launch(st, fa));
  if (kev && kev->on) CK(cudaEventRecord(kev->ev[3], st));
  return MPPI_OK;
}

// the device-resident closed loop of mppi_bench: one graph = { rollout, reduce + finalize }
static mppi_status build_graphs(mppi_engine* e) {
  if (!e->dirty) return MPPI_OK;
  drop_graphs(e);
  cudaGraph_t g = nullptr;
  CK(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
  mppi_status s = launch_local(e, e->stream, e->p.precision, FUSE_LOOP, nullptr);
  cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
  if (s != MPPI_OK || ce != cudaSuccess) {
    if (g) cudaGraphDestroy(g);
    if (s == MPPI_OK) set_err("cudaStreamEndCapture: %s", cudaGetErrorString(ce));
    return MPPI_ERR_CUDA;
  }
  cudaGraphExec_t ge = nullptr;
  ce = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (ce != cudaSuccess) {
    set_err("cudaGraphInstantiate: %s", cudaGetErrorString(ce));
    return MPPI_ERR_CUDA;
  }
  e->g_loop = ge;
  e->dirty = false;
  return MPPI_OK;
}

// wait until the finalize phase has published the result of step e->seq into mapped host memory (flag-in-data words,
// common.cuh: HostWire).  Polling host memory costs ~0.1 us of latency against ~5 us for a blocking stream synchronisation; the stream is
// queried now and then so that a faulted kernel cannot hang the caller.
static mppi_status wait_result(mppi_engine* e) {
  volatile unsigned long long* w = e->h_res->w;
  const unsigned long long tag = e->seq & 0xffffffffull;
  auto complete = [&]() {
    for (int i = 0; i < kHostWords; ++i)
      if ((w[i] >> 32) != tag) return false;
    return true;
  };
  if (e->wait_block) {
    CK(cudaStreamSynchronize(e->stream));
  } else {
    for (unsigned int spin = 1;; ++spin) {
      if ((w[kHostWords - 1] >> 32) == tag && complete()) break;
      if ((spin & 0xfffu) == 0) {
        const cudaError_t q = cudaStreamQuery(e->stream);
        if (q == cudaSuccess) break;             // everything retired: the words are checked below
        if (q != cudaErrorNotReady) {
          set_err("mppi_step: %s", cudaGetErrorString(q));
          return MPPI_ERR_CUDA;
        }
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (!complete()) {
    set_err("mppi_step: the step retired without publishing its result (sequence %llu)", e->seq);
    return MPPI_ERR_STATE;
  }
  unsigned int pay[kHostWords];
  for (int i = 0; i < kHostWords; ++i) pay[i] = (unsigned int)(w[i] & 0xffffffffull);
  auto dbl = [&](int i) {
    const unsigned long long bits = (unsigned long long)pay[2 * i] | ((unsigned long long)pay[2 * i + 1] << 32);
    double v;
    memcpy(&v, &bits, sizeof(v));
    return v;
  };
  HostResult& r = e->res;
  r.out_u[0] = dbl(0);
  r.out_u[1] = dbl(1);
  r.out_x[0] = dbl(2);
  r.out_x[1] = dbl(3);
  r.out_x[2] = dbl(4);
  r.max_dev = dbl(5);
  r.head = dbl(6);
  r.status = (int)pay[14];
  r.candidates = (int)pay[15];
  r.overflow_total = (int)pay[16];
  r.seq = e->seq;
  return MPPI_OK;
}

// world > 1: align the ranks on the device before a timed step (outside the timed interval; see reduce.cu)
static mppi_status rendezvous(mppi_engine* e) {
  if (e->sp.world <= 1 || !e->p2p_on) return MPPI_OK;
  RendezvousArgs ra;
  memset(&ra, 0, sizeof(ra));
  for (int g = 0; g < kMaxFusedWorld; ++g) ra.peers[g] = e->ll_peers[g];
  ra.rows_uint2 = e->ll_rows_uint2;
  ra.epoch = ++e->rdv_epoch;
  ra.world = e->sp.world;
  ra.rank = e->p.rank;
  CK(rendezvous_launch(e->stream, ra));
  return MPPI_OK;
}

static StepInput step_input(const mppi_engine* e) {
  StepInput in;
  for (int i = 0; i < 6; ++i) in.x0g[i] = e->h_in[i];
  in.from_args = 1;
  return in;
}

static mppi_status finish_outputs(mppi_engine* e, double u_out[2], double x_next[3]) {
  const HostResult* o = &e->res;
  if (o->status == kStatusRedoF64) {
    // MIXED: a candidate list overflowed; redo this step entirely in fp64 (same noise: the step
    // counter was not advanced, U and x0 are untouched).
    if (e->sp.world > 1 && !e->p2p_on) {
      // split-phase step with an external exchange: the host owns the exchange, so the redo is a second round trip of
      // mppi_step_local (now fp64: the hold-off is armed) / exchange / mppi_step_finish.  Every rank merged the same
      // records, so every rank lands here for the same step and keeps the same hold-off schedule.
      e->last.refine_overflow += 1;
      e->f64_holdoff = e->f64_backoff + 1;   // + 1: the redo itself consumes one
      e->f64_backoff = e->f64_backoff * 2 > 64 ? 64 : e->f64_backoff * 2;
      set_err("the fp32 screen overflowed on some rank: repeat mppi_step_local / exchange / mppi_step_finish for this step (it runs in fp64)");
      return MPPI_ERR_RETRY;
    }
    const StepInput in = step_input(e);
    e->seq += 1;
    CKS(launch_local(e, e->stream, MPPI_PRECISION_F64, FUSE_STEP, nullptr, &in));
    CKS(wait_result(e));
    e->last.refine_overflow += 1;
    e->f64_holdoff = e->f64_backoff;
    e->f64_backoff = e->f64_backoff * 2 > 64 ? 64 : e->f64_backoff * 2;
  } else if (e->attempted_mixed) {
    e->f64_backoff = 8;   // a mixed step went through: the next overflow starts from the short hold-off again
  }
  e->last.refine_candidates = o->candidates;
  e->last.refine_max_dev = o->max_dev;
  e->last.refine_head_room = o->head;
  if (u_out) {
    u_out[0] = o->out_u[0];
    u_out[1] = o->out_u[1];
  }
  if (x_next) {
    x_next[0] = o->out_x[0];
    x_next[1] = o->out_x[1];
    x_next[2] = o->out_x[2];
  }
  if (o->status == MPPI_ERR_NONFINITE) {
    set_err("non-finite control update (NaN/Inf input propagated, as in the reference)");
    return MPPI_ERR_NONFINITE;
  }
  if (o->status == MPPI_ERR_STATE) {
    set_err("peer-to-peer exchange timed out (a rank did not deliver its record)");
    return MPPI_ERR_STATE;
  }
  return MPPI_OK;
}

static mppi_status pre_step(mppi_engine* e, const double x0[3]) {
  if (!x0) {
    set_err("x0 is null");
    return MPPI_ERR_INVALID;
  }
  for (int i = 0; i < 3; ++i) {
    e->h_in[i] = x0[i];
    e->h_in[3 + i] = e->goal[i];
    e->last_x0[i] = x0[i];
    e->last_goal[i] = e->goal[i];
  }
  if (e->sp.capture) {   // debug: remember the nominal this step starts from (offset of get_cost_to_go)
    COPY_SYNC(e, e->U_prev.data(), e->d_Umaster, 2 * e->sp.T * sizeof(double), cudaMemcpyDeviceToHost);
  }
  return MPPI_OK;
}

extern "C" mppi_status mppi_step(mppi_handle e, const double x0[3], double u_out[2], double x_next[3]) {
  ENTER(e);
  if (e->local_pending) {
    set_err("mppi_step called between mppi_step_local and mppi_step_finish");
    return MPPI_ERR_STATE;
  }
  if (e->sp.world > 1 && !e->p2p_on) {
    set_err("world_size > 1: connect the peer-to-peer exchange (mppi_p2p_connect) or use mppi_step_local / mppi_step_finish");
    return MPPI_ERR_STATE;
  }
  const auto tp0 = std::chrono::steady_clock::now();
  CKS(pre_step(e, x0));
  // two launches, no copies: x0 / goal ride in the kernel arguments, the result comes back through mapped host memory
  const StepInput in = step_input(e);
  e->seq += 1;
  int precision = e->p.precision;
  if (precision == MPPI_PRECISION_MIXED) {
    if (e->f64_holdoff > 0) {
      precision = MPPI_PRECISION_F64;   // same result (mixed == f64 to rounding), without the screen that would overflow
      e->f64_holdoff -= 1;
    }
  }
  e->attempted_mixed = precision == MPPI_PRECISION_MIXED;
  CKS(launch_local(e, e->stream, precision, FUSE_STEP, nullptr, &in));
  CKS(wait_result(e));
  const auto tp3 = std::chrono::steady_clock::now();
  const mppi_status st = finish_outputs(e, u_out, x_next);
  const auto tp4 = std::chrono::steady_clock::now();
  typedef std::chrono::duration<double, std::nano> ns;
  e->host_ns[0] += ns(e->tp_launch1 - tp0).count();
  e->host_ns[1] += ns(e->tp_launch2 - e->tp_launch1).count();
  e->host_ns[2] += ns(tp3 - e->tp_launch2).count();
  e->host_ns[3] += ns(tp4 - tp3).count();
  e->host_calls += 1;
  return st;
}

// profiling aid: mean host-side duration (us) of the phases of mppi_step since the last call of this function:
// out[0] entry -> rollout kernel launched, [1] -> reduce kernel launched, [2] -> result seen in mapped memory, [3] -> return
extern "C" mppi_status mppi_debug_host_timing(mppi_handle e, double out[4]) {
  ENTER(e);
  for (int i = 0; i < 4; ++i) {
    if (out) out[i] = e->host_calls ? e->host_ns[i] / (double)e->host_calls * 1e-3 : 0.0;
    e->host_ns[i] = 0;
  }
  e->host_calls = 0;
  return MPPI_OK;
}

extern "C" mppi_status mppi_step_local(mppi_handle e, const double x0[3]) {
  ENTER(e);
  if (e->local_pending) {
    set_err("mppi_step_local called twice");
    return MPPI_ERR_STATE;
  }
  CKS(pre_step(e, x0));
  CK(cudaMemcpyAsync(e->d_dyn, e->h_in, 6 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  int precision = e->p.precision;
  if (precision == MPPI_PRECISION_MIXED && e->f64_holdoff > 0) {   // same back-off as mppi_step (identical on every rank)
    precision = MPPI_PRECISION_F64;
    e->f64_holdoff -= 1;
  }
  e->attempted_mixed = precision == MPPI_PRECISION_MIXED;
  CKS(launch_local(e, e->stream, precision, FUSE_NONE, nullptr));
  e->local_pending = true;
  return MPPI_OK;
}

extern "C" mppi_status mppi_exchange_buffers(mppi_handle e, void** record, size_t* record_bytes, void** gather,
                                             size_t* gather_bytes) {
  ENTER(e);
  const size_t rb = (size_t)e->sp.T * kRecordStride * sizeof(double);
  if (record) *record = e->d_record;
  if (record_bytes) *record_bytes = rb;
  if (gather) *gather = e->d_gather;
  if (gather_bytes) *gather_bytes = rb * e->sp.world;
  return MPPI_OK;
}

extern "C" mppi_status mppi_read_record(mppi_handle e, double* record) {
  ENTER(e);
  if (!record) return MPPI_ERR_INVALID;
  CK(cudaStreamSynchronize(e->stream));
  CK(memcpy_on(e, record, e->d_record, (size_t)e->sp.T * kRecordStride * sizeof(double), cudaMemcpyDeviceToHost));
  return MPPI_OK;
}

extern "C" mppi_status mppi_write_gather(mppi_handle e, const double* all) {
  ENTER(e);
  if (!all) return MPPI_ERR_INVALID;
  CK(cudaMemcpyAsync(e->d_gather, all, (size_t)e->sp.world * e->sp.T * kRecordStride * sizeof(double),
                     cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return MPPI_OK;
}

extern "C" mppi_status mppi_p2p_export(mppi_handle e, void* handle64) {
  ENTER(e);
  if (!handle64) return MPPI_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t hnd;
  CK(cudaIpcGetMemHandle(&hnd, e->d_ll));
  memcpy(handle64, &hnd, 64);
  return MPPI_OK;
}

extern "C" mppi_status mppi_p2p_connect(mppi_handle e, const void* handles) {
  ENTER(e);
  if (!handles) {
    set_err("mppi_p2p_connect: call mppi_p2p_export on every rank first and pass all world_size handles");
    return MPPI_ERR_STATE;
  }
  const int world = e->sp.world;
  if (world > kMaxFusedWorld) {
    set_err("the fused peer-to-peer exchange supports up to %d ranks; use mppi_step_local / mppi_step_finish", kMaxFusedWorld);
    return MPPI_ERR_UNSUPPORTED;
  }
  for (int g = 0; g < world; ++g) {
    if (g == e->p.rank) {
      e->ll_peers[g] = e->d_ll;
      continue;
    }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, (const char*)handles + (size_t)g * 64, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    e->p2p_opened.push_back(ptr);
    e->ll_peers[g] = (uint2*)ptr;
  }
  e->p2p_on = true;
  return drop_graphs(e);
}

extern "C" mppi_status mppi_step_finish(mppi_handle e, double u_out[2], double x_next[3]) {
  ENTER(e);
  if (!e->local_pending) {
    set_err("mppi_step_finish without mppi_step_local");
    return MPPI_ERR_STATE;
  }
  e->local_pending = false;
  e->seq += 1;
  CKS(launch_finalize(e, e->stream, false, nullptr, true));
  CKS(wait_result(e));
  return finish_outputs(e, u_out, x_next);
}

// ---- value-function offset dropped by the delta-cost formulation (see common.cuh) -----------------
static void cost_offsets(const mppi_engine* e, const double* U, const double x0[3], const double goal[3],
                         std::vector<double>& off) {
  const int T = e->sp.T;
  if (e->user_has_cost) {   // a user cost functor is evaluated on the absolute state: nothing was dropped
    off.assign(T, 0.0);
    return;
  }
  const double a[3] = {x0[0] - goal[0], x0[1] - goal[1], x0[2] - goal[2]};
  const double still = 0.5 * (e->sp.q[0] * a[0] * a[0] + e->sp.q[1] * a[1] * a[1] + e->sp.q[2] * a[2] * a[2]);
  const double term = e->sp.p1[0] * a[0] * a[0] + e->sp.p1[1] * a[1] * a[1] + e->sp.p1[2] * a[2] * a[2];
  off.assign(T, 0.0);
  double acc = term;
  const double* R = e->p.r;
  for (int t = T - 1; t >= 0; --t) {
    const double u0 = U[t], u1 = U[T + t];
    // 1/2 u'Ru, control/src/mppi:183
    acc += 0.5 * (u0 * (R[0] * u0 + R[1] * u1) + u1 * (R[2] * u0 + R[3] * u1)) + still;
    off[t] = acc;
  }
  if (e->sp.weighting == MPPI_WEIGHT_TOTAL_COST)
    for (int t = 1; t < T; ++t) off[t] = off[0];
}

static mppi_status read_vcap(mppi_engine* e, int kind, const std::vector<double>& off, double* V) {
  const int T = e->sp.T, K = e->sp.K;
  const size_t n = (size_t)T * K;
  if (kind == ROLLOUT_F64_SOFTMIN) {
    CK(memcpy_on(e, V, e->d_vcap, n * sizeof(double), cudaMemcpyDeviceToHost));
  } else {
    std::vector<float> tmp(n);
    CK(memcpy_on(e, tmp.data(), e->d_vcap, n * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) V[i] = (double)tmp[i];
  }
  for (int t = 0; t < T; ++t)
    for (int k = 0; k < K; ++k) V[(size_t)t * K + k] += off[t];
  return MPPI_OK;
}

extern "C" mppi_status mppi_get_cost_to_go(mppi_handle e, double* V) {
  ENTER(e);
  if (!V) return MPPI_ERR_INVALID;
  if (!e->sp.capture || e->last_capture_kind < 0) {
    set_err("mppi_get_cost_to_go: enable mppi_set_capture(h, 1) before the step");
    return MPPI_ERR_STATE;
  }
  CK(cudaStreamSynchronize(e->stream));
  std::vector<double> off;
  cost_offsets(e, e->U_prev.data(), e->last_x0, e->last_goal, off);
  return read_vcap(e, e->last_capture_kind, off, V);
}

// ---- standalone reference-API ops (always fp64) ---------------------------------------------------
extern "C" mppi_status mppi_cost_to_go(mppi_handle e, const double x0[3], const double* U, const double goal[3],
                                       const double* eps, double* V) {
  ENTER(e);
  if (!x0 || !U || !goal || !eps || !V) return MPPI_ERR_INVALID;
  CK(cudaStreamSynchronize(e->stream));
  const int T = e->sp.T;
  const size_t ne = (size_t)T * 2 * e->sp.K;
  // save
  StaticParams sp_save = e->sp;
  double* eps_save = e->d_eps_ext;
  DynState dyn_save;
  CK(memcpy_on(e, &dyn_save, e->d_dyn, sizeof(DynState), cudaMemcpyDeviceToHost));
  CK(memcpy_on(e, e->d_Utmp, e->d_Umaster, 2 * T * sizeof(double), cudaMemcpyDeviceToDevice));
  double* eps_tmp = nullptr;
  CK(cudaMalloc(&eps_tmp, ne * sizeof(double)));
  mppi_status s = MPPI_OK;
  do {
    if (!e->d_vcap && cudaMalloc(&e->d_vcap, (size_t)T * e->sp.K * sizeof(double)) != cudaSuccess) {
      s = MPPI_ERR_CUDA;
      break;
    }
    DynState d = dyn_save;
    for (int i = 0; i < 3; ++i) {
      d.x0[i] = x0[i];
      d.goal[i] = goal[i];
    }
    cudaError_t ce = memcpy_on(e, eps_tmp, eps, ne * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = memcpy_on(e, e->d_Umaster, U, 2 * T * sizeof(double), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = memcpy_on(e, e->d_dyn, &d, sizeof(d), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
      set_err("mppi_cost_to_go: staging copy failed: %s", cudaGetErrorString(ce));
      s = MPPI_ERR_CUDA;
      break;
    }
    e->d_eps_ext = eps_tmp;
    e->sp.noise_external = 1;
    e->sp.capture = 1;
    if ((s = configure(e)) != MPPI_OK) break;
    if (!e->cfg[ROLLOUT_F64_SOFTMIN].ready) {
      set_err("mppi_cost_to_go: the fp64 rollout kernel has no launch configuration for T=%d (its cost tile needs too much shared memory)", T);
      s = MPPI_ERR_UNSUPPORTED;
      break;
    }
    if ((s = prep_nominal(e)) != MPPI_OK) break;
    if ((s = launch_local(e, e->stream, MPPI_PRECISION_F64, FUSE_NONE, nullptr)) != MPPI_OK) break;
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) {
      set_err("cost_to_go kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
      s = MPPI_ERR_CUDA;
      break;
    }
    std::vector<double> off;
    cost_offsets(e, U, x0, goal, off);
    s = read_vcap(e, ROLLOUT_F64_SOFTMIN, off, V);
  } while (0);
  // restore
  e->sp = sp_save;
  e->d_eps_ext = eps_save;
  e->last_capture_kind = -1;
  configure(e);
  memcpy_on(e, e->d_dyn, &dyn_save, sizeof(DynState), cudaMemcpyHostToDevice);
  memcpy_on(e, e->d_Umaster, e->d_Utmp, 2 * T * sizeof(double), cudaMemcpyDeviceToDevice);
  cudaFree(eps_tmp);
  mppi_status s2 = prep_nominal(e);
  return s != MPPI_OK ? s : s2;
}

extern "C" mppi_status mppi_update_action(mppi_handle e, const double* U_in, const double* eps, const double* V,
                                          double* U_out) {
  ENTER(e);
  if (!U_in || !eps || !V || !U_out) return MPPI_ERR_INVALID;
  CK(cudaStreamSynchronize(e->stream));
  const int T = e->sp.T, K = e->sp.K;
  double *dV = nullptr, *deps = nullptr;
  CK(cudaMalloc(&dV, (size_t)T * K * sizeof(double)));
  if (cudaMalloc(&deps, (size_t)T * 2 * K * sizeof(double)) != cudaSuccess) {
    cudaFree(dV);
    set_err("cudaMalloc failed");
    return MPPI_ERR_CUDA;
  }
  mppi_status s = MPPI_OK;
  cudaError_t ce = memcpy_on(e, dV, V, (size_t)T * K * sizeof(double), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = memcpy_on(e, deps, eps, (size_t)T * 2 * K * sizeof(double), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = memcpy_on(e, e->d_Utmp, e->d_Umaster, 2 * T * sizeof(double), cudaMemcpyDeviceToDevice);
  if (ce == cudaSuccess) ce = memcpy_on(e, e->d_Utmp + 2 * T, e->d_Ulast, 2 * T * sizeof(double), cudaMemcpyDeviceToDevice);
  if (ce == cudaSuccess) ce = memcpy_on(e, e->d_Umaster, U_in, 2 * T * sizeof(double), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = weights_from_v_launch(e->stream, e->sp, e->d_dyn, dV, deps, e->d_record_tmp);
  if (ce == cudaSuccess) {
    FinalizeArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.sp = e->sp;
    fa.sp.world = 1;
    fa.sp.k_total = K;
    fa.dyn = e->d_dyn;
    fa.gather = e->d_record_tmp;
    fa.Umaster = e->d_Umaster;
    fa.Ulast = e->d_Ulast;
    fa.nomF = e->d_nomF;
    fa.nomD = e->d_nomD;
    fa.sg_a = e->sg_a;
    fa.sg_b = e->sg_b;
    for (int i = 0; i < 4; ++i) fa.sg_inv_norm[i] = e->sg_inv_norm[i];
    fa.mode = 1;
    ce = e->user ? user_finalize_launch(e->user, e->stream, fa) : finalize_launch(e->stream, fa);
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce == cudaSuccess) ce = memcpy_on(e, U_out, e->d_Ulast, 2 * T * sizeof(double), cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) {
    set_err("mppi_update_action: %s", cudaGetErrorString(ce));
    s = MPPI_ERR_CUDA;
  }
  memcpy_on(e, e->d_Umaster, e->d_Utmp, 2 * T * sizeof(double), cudaMemcpyDeviceToDevice);
  memcpy_on(e, e->d_Ulast, e->d_Utmp + 2 * T, 2 * T * sizeof(double), cudaMemcpyDeviceToDevice);
  cudaFree(dV);
  cudaFree(deps);
  return s;
}

extern "C" mppi_status mppi_model_step(mppi_handle e, const double* x, const double* u, int32_t n, double* x_out) {
  ENTER(e);
  if (!x || !u || !x_out || n < 1) return MPPI_ERR_INVALID;
  double* d = nullptr;
  CK(cudaMalloc(&d, (size_t)8 * n * sizeof(double)));
  cudaError_t ce = memcpy_on(e, d, x, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = memcpy_on(e, d + 3 * (size_t)n, u, (size_t)2 * n * sizeof(double), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess)
    ce = e->user ? user_model_step_launch(e->user, e->stream, e->sp, d, d + 3 * (size_t)n, n, d + 5 * (size_t)n)
                 : model_step_launch(e->stream, e->sp, d, d + 3 * (size_t)n, n, d + 5 * (size_t)n);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
  if (ce == cudaSuccess) ce = memcpy_on(e, x_out, d + 5 * (size_t)n, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d);
  CK(ce);
  return MPPI_OK;
}

extern "C" mppi_status mppi_perform_action(mppi_handle e, const double x0[3], const double* U, double x_out[3]) {
  if (!e || !x0 || !U || !x_out) return MPPI_ERR_INVALID;
  const double u[2] = {U[0], U[e->sp.T]};   // uvec[:,0], control/src/mppi:212
  return mppi_model_step(e, x0, u, 1, x_out);
}

// ---- measurement -------------------------------------------------------------------------------------
extern "C" mppi_status mppi_last_stats(mppi_handle e, mppi_timing* out) {
  ENTER(e);
  if (!out) return MPPI_ERR_INVALID;
  *out = e->last;
  return MPPI_OK;
}

extern "C" mppi_status mppi_bench(mppi_handle e, const double x0[3], int32_t steps, int32_t warmup, int32_t flush_l2,
                                  int32_t per_kernel, mppi_timing* out) {
  ENTER(e);
  if (!x0 || !out || steps < 1 || warmup < 0) return MPPI_ERR_INVALID;
  if (!e->own_stream || (e->sp.world > 1 && !e->p2p_on)) {
    set_err("mppi_bench needs an engine-owned stream and world_size 1 (or a connected p2p exchange)");
    return MPPI_ERR_STATE;
  }
  CKS(pre_step(e, x0));
  {
    const char* bm0 = getenv("MPPI_B200_BENCH");
    if (bm0 && !strcmp(bm0, "graph")) CKS(build_graphs(e));
  }
  CK(memcpy_on(e, e->d_dyn, e->h_in, 6 * sizeof(double), cudaMemcpyHostToDevice));
  int ovf_before = 0;
  CK(memcpy_on(e, &ovf_before, &e->d_dyn->overflow_total, sizeof(int), cudaMemcpyDeviceToHost));
  if (flush_l2 && !e->d_flush) {
    e->flush_bytes = (size_t)512 << 20;   // 256 MiB overwritten + 256 MiB read, each > 126 MB L2 (reduce.cu: flush_l2_kernel)
    CK(cudaMalloc(&e->d_flush, e->flush_bytes));
    CK(cudaMemset(e->d_flush, 0, e->flush_bytes));
  }
  // The two kernels of a step are launched directly; MPPI_B200_BENCH=graph replays them from the captured CUDA graph instead.
  // The host runs far ahead of the device in this loop either way, and measured on this part the graph's device-side start
  // latency is the longer one: 39.2 us per step through the graph against 38.8 us eager at N = 1, 54.0 against 52.9 us at
  // N = 8 (profiles/README.md) -- a graph saves HOST launch time, which is not on this loop's critical path.
  const char* bm = getenv("MPPI_B200_BENCH");
  const bool eager = !(bm && !strcmp(bm, "graph"));
  auto launch_step = [&]() -> mppi_status {
    if (eager) return launch_local(e, e->stream, e->p.precision, FUSE_LOOP, nullptr);
    CK(cudaGraphLaunch(e->g_loop, e->stream));
    return MPPI_OK;
  };
  for (int i = 0; i < warmup; ++i) CKS(launch_step());
  CK(cudaStreamSynchronize(e->stream));
  std::vector<cudaEvent_t> ev((size_t)2 * steps);
  for (auto& v : ev) CK(cudaEventCreate(&v));
  for (int i = 0; i < steps; ++i) {
    if (flush_l2) CK(flush_l2_launch(e->stream, e->d_flush, e->flush_bytes, (unsigned)(i & 0xff)));
    CKS(rendezvous(e));
    CK(cudaEventRecord(ev[2 * i], e->stream));
    CKS(launch_step());
    CK(cudaEventRecord(ev[2 * i + 1], e->stream));
  }
  CK(cudaStreamSynchronize(e->stream));
  double total = 0;
  for (int i = 0; i < steps; ++i) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
    total += ms;
  }
  for (auto& v : ev) cudaEventDestroy(v);
  mppi_timing t{};
  t.step_ms = (float)(total / steps);
  t.steps = steps;
  t.launches = 2 * steps;   // rollout + reduce (row exchange and finalize run inside the reduce kernel)
  if (per_kernel) {
    KernelEvents kev;
    kev.on = true;
    for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&kev.ev[i]));
    double acc[3] = {0, 0, 0};
    for (int i = 0; i < steps; ++i) {
      if (flush_l2) CK(flush_l2_launch(e->stream, e->d_flush, e->flush_bytes, (unsigned)(i & 0xff)));
      CKS(launch_local(e, e->stream, e->p.precision, FUSE_LOOP, &kev));
      CK(cudaEventRecord(kev.ev[3], e->stream));
      CK(cudaStreamSynchronize(e->stream));
      for (int j = 0; j < 3; ++j) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, kev.ev[j], kev.ev[j + 1]));
        acc[j] += ms;
      }
    }
    for (int i = 0; i < 4; ++i) cudaEventDestroy(kev.ev[i]);
    t.rollout_ms = (float)(acc[0] / steps);
    t.reduce_ms = (float)(acc[1] / steps);
    t.finalize_ms = (float)(acc[2] / steps);
  }
  CK(memcpy_on(e, e->h_out, e->d_dyn, sizeof(DynState), cudaMemcpyDeviceToHost));
  t.refine_candidates = e->h_out->last_candidates;
  t.refine_overflow = e->h_out->overflow_total;
  t.refine_max_dev = e->h_out->last_max_dev;
  t.refine_head_room = e->h_out->last_head;
  e->last = t;
  *out = t;
  if (e->h_out->overflow_total != ovf_before) {
    // the device-resident loop has no host in it to redo an overflowing MIXED step in fp64: such a step leaves U, x0 and the
    // noise counter untouched and every later graph launch would replay it -- the timing would be of a loop that stands still
    set_err("mppi_bench: the fp32 screen of precision MIXED overflowed inside the device-resident loop (%d step(s)); the timed "
            "region is invalid -- start further from the goal, or bench precision F64 / F32", e->h_out->overflow_total - ovf_before);
    return MPPI_ERR_STATE;
  }
  return MPPI_OK;
}

// profiling aid: globaltimer stamps (ns) of the reduce kernel phases of the LAST step, [T][8];
// the first call only arms the stamps (and invalidates the CUDA graph)
extern "C" mppi_status mppi_debug_reduce_timestamps(mppi_handle e, unsigned long long* out) {
  ENTER(e);
  CK(cudaStreamSynchronize(e->stream));
  const size_t n = (size_t)(e->sp.T + 1) * 8;   // row T: the finalizer block
  if (!e->d_debug_ts) {
    CK(cudaMalloc(&e->d_debug_ts, n * sizeof(unsigned long long)));
    CK(cudaMemset(e->d_debug_ts, 0, n * sizeof(unsigned long long)));
    drop_graphs(e);
  }
  if (out) CK(memcpy_on(e, out, e->d_debug_ts, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return MPPI_OK;
}

// profiling aid (libraries built with -DMPPI_EXP_TIMELINE): globaltimer stamps of the rollout kernel's CTAs, [n_ctas][8]
// (entry, loads issued, prologue done, loop done, total stored, barrier, tile done, SM id); first call arms
extern "C" mppi_status mppi_debug_rollout_timestamps(mppi_handle e, unsigned long long* out, size_t n_ctas) {
  ENTER(e);
  CK(cudaStreamSynchronize(e->stream));
  if (!e->d_debug_rts) {
    e->debug_rts_ctas = e->part_capacity_ctas;   // one row per partial record (CTA, or tile of the SM-wide kernel)
    CK(cudaMalloc(&e->d_debug_rts, e->debug_rts_ctas * 8 * sizeof(unsigned long long)));
    CK(cudaMemset(e->d_debug_rts, 0, e->debug_rts_ctas * 8 * sizeof(unsigned long long)));
    drop_graphs(e);
  }
  if (out) {
    if (n_ctas > e->debug_rts_ctas) n_ctas = e->debug_rts_ctas;
    CK(memcpy_on(e, out, e->d_debug_rts, n_ctas * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  return MPPI_OK;
}

extern "C" mppi_status mppi_io_bytes(mppi_handle e, size_t* h2d, size_t* d2h) {
  ENTER(e);
  // x0 + goal ride in the argument buffers of the two kernels; the result block is stored into mapped host memory
  if (h2d) *h2d = 2 * 6 * sizeof(double);
  if (d2h) *d2h = sizeof(HostWire);
  return MPPI_OK;
}

// measurement aid: overwrite a buffer larger than L2 so that the next step starts from a cold cache
extern "C" mppi_status mppi_debug_flush_l2(mppi_handle e) {
  ENTER(e);
  if (!e->d_flush) {
    e->flush_bytes = (size_t)512 << 20;   // 256 MiB overwritten + 256 MiB read, each > 126 MB L2 (reduce.cu: flush_l2_kernel)
    CK(cudaMalloc(&e->d_flush, e->flush_bytes));
    CK(cudaMemset(e->d_flush, 0, e->flush_bytes));
  }
  CK(flush_l2_launch(e->stream, e->d_flush, e->flush_bytes, (unsigned)(e->seq & 0xff)));
  CKS(rendezvous(e));
  CK(cudaStreamSynchronize(e->stream));
  return MPPI_OK;
}

extern "C" mppi_status mppi_launch_info(mppi_handle e, int32_t info[8]) {
  ENTER(e);
  if (!info) return MPPI_ERR_INVALID;
  const KindCfg& c = e->cfg[kind_of(e->p.precision)];
  info[0] = c.block;
  info[1] = c.grid;
  info[2] = c.ntiles;
  info[3] = (int32_t)c.smem;
  info[4] = c.ctas_per_sm;
  info[5] = c.regs;
  info[6] = c.variant;
  info[7] = c.nparts;
  return MPPI_OK;
}

extern "C" mppi_status mppi_measure_fp32_peak(int32_t device, double* tflops, double* sm_clock_mhz) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_err("no such CUDA device");
    return MPPI_ERR_NO_DEVICE;
  }
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
  float* d = nullptr;
  CK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(a, 0));
    CK(fp32_peak_launch(0, blocks, threads, d, iters));
    CK(cudaEventRecord(b, 0));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    const double fl = (double)blocks * threads * (double)iters * 16 * 8 * 2;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  if (tflops) *tflops = best;
  if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
  return MPPI_OK;
}
