// common.cuh -- shared device/host definitions of the B200 MPPI engine.
//
// Hot path restated: MPPI.get_path of the reference (control/src/mppi:85-102), i.e.
// get_cost2go (:127-178) -> update_action (:186-208) -> perform_action (:210-213) -> shift.
// Everything here is new code written for sm_100a; reference lines are cited for semantics only.
#pragma once
#ifndef __CUDACC_RTC__   // (run-time compilation of a user model gets the CUDA built-ins without host headers)
#include <cuda_runtime.h>
#include <stdint.h>
#endif

#include "../../include/mppi_b200.h"

#ifndef MPPI_PHILOX_ROUNDS
#define MPPI_PHILOX_ROUNDS 10   // Philox4x32-10 (the cuRAND default); 7 is the smallest Crush-resistant count
#endif

namespace mppi {

constexpr int kMaxCand = 16;        // SCREEN: candidate slots per (CTA, t)
constexpr int kMaxRefine = 256;     // SCREEN: fp64 re-evaluations per t after the global filter
constexpr int kRecordStride = 6;    // doubles per t in an exchange record: m, S, N0, N1, E0, E1
constexpr int kRowDoubles = 8;      // fused step: a row = the record + (max |V32 - V64|, candidates) of that row
constexpr int kMaxFusedWorld = 16;  // ranks of a fused (peer-to-peer) exchange; larger groups use the split-phase step
constexpr int kRowWords = 2 * kRowDoubles;   // 32-bit payload words per row, each travelling with its own flag (8-byte stores)
constexpr int kRow2Doubles = 6;     // fused step, merged row of one t: clip(U0 + dU0), clip(U1 + dU1), flag bits, candidates, max dev, -
constexpr int kRow2Words = 2 * kRow2Doubles;
constexpr int kRow2Overflow = 1, kRow2Bad = 2, kRow2Timeout = 4;   // flag bits of a merged row
constexpr double kZFixScale = 1048576.0;   // 2^20: fixed-point scale of the floor-term noise sums
constexpr double kLeanMaxYawInc = 0.125;   // LEAN rollout kernel: admission bound on |dt * yaw rate| (half of it for Euler)

enum RolloutMode { MODE_SOFTMIN = 0, MODE_SCREEN = 1 };

// ---- static (engine-lifetime) parameters: passed by value in the constant bank -----------------
struct StaticParams {
  int K;                 // rollouts on this device
  int T;
  long long k_offset;    // global id of rollout 0 of this device (Philox counter)
  long long k_total;     // global K (floor term 1e-8*K, control/src/mppi:193-195)
  int model;
  int weighting;
  int noise_external;    // 0: Philox in registers, 1: replay eps (T,2,K) f64 from HBM
  int capture;           // 1: write cost-to-go (T,K) to HBM (debug get_cost2go)
  int has_grid;
  int grid_in_smem;
  int gW, gH;
  int grid_bytes_padded; // multiple of 16 (TMA bulk copy granularity)
  int world;             // ranks exchanging records
  double dt;
  double q[3];
  double p1[3];
  double u_max[2];
  double wheel_r, wheel_L;
  double eps_floor;
  double g_inv_res, g_x0, g_y0, w_obs;
  double margin;         // SCREEN window (cost units): the soft-min SUPPORT part, 40 lam (or the caller's whole window)
  // head-room of the window for the fp32 error of the screened cost-to-go, scaled with the cost magnitude of the step:
  //   head = max(head_min, head_scale * 2^-23 * Vmax),  Vmax = vm_c0 + vm_ca * |pos(x0) - pos(goal)| + vm_cth * |th0 - th_goal|
  // (Vmax bounds |V| in delta form over the horizon; head_scale == 0: static window = margin, the caller's refine_margin)
  double head_scale, head_min, vm_c0, vm_ca, vm_cth;
  unsigned long long seed;
};

// the SCREEN window of a step and its head-room part.  Evaluated with explicitly rounded operations so that the rollout
// kernel (which lists candidates against it) and the reduce kernel (which filters them against the global minimum) get the
// bit-identical value from the same (x0, goal).
__device__ __forceinline__ double screen_window(const StaticParams& sp, const double x0[3], const double g[3], double* head_out = nullptr) {
  if (sp.head_scale == 0.0) {
    if (head_out) *head_out = 0.5 * sp.margin;
    return sp.margin;
  }
  const double ax = __dsub_rn(x0[0], g[0]), ay = __dsub_rn(x0[1], g[1]), at = fabs(__dsub_rn(x0[2], g[2]));
  const double a = __dsqrt_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)));
  const double v = __dadd_rn(__dadd_rn(sp.vm_c0, __dmul_rn(sp.vm_ca, a)), __dmul_rn(sp.vm_cth, at));
  const double head = fmax(sp.head_min, __dmul_rn(sp.head_scale * 1.1920928955078125e-07, v));
  if (head_out) *head_out = head;
  return __dadd_rn(sp.margin, head);
}

// ---- dynamic state living in HBM (changes every step; lets the CUDA graph stay static) ---------
struct DynState {
  double x0[3];          // start state of this step          (get_path arg, control/src/mppi:86)
  double goal[3];        // goal                              (get_path arg, :87)
  double lam;            // :89
  double sig[4];         // :88
  double R[4];           // :71
  double noise_std[2];   // :144-146
  unsigned int step;     // Philox step counter (advanced by finalize)
  unsigned int pad0;
  // outputs of finalize
  double out_u[2];       // uvec[-1] = U[:,0] before shift    (:96-97)
  double out_x[3];       // predicted next state              (:94,102)
  int status;            // mppi_status seen on device (NONFINITE), or kStatusRedoF64
  int refine_candidates; // accumulators of the running step (reduce_screen_kernel)
  int refine_overflow;
  int last_candidates;   // statistics of the last finished step (finalize_kernel)
  double refine_max_dev;
  double last_max_dev;
  double last_head;      // head-room of the screening window of the last finished step
  int overflow_total;    // steps that hit a candidate-list overflow since creation
  unsigned int xchg;     // row-exchange epoch: +1 per step, never rewound (the rows' flag-in-data words carry xchg + 1)
  // fp32 mirrors of lam / noise_std kept by the host (LEAN rollout prologue: no fp64 division, no conversions)
  float neg_inv_lam_f;
  float noise_std_f[2];
  float pad1;
};
// result of one step as the host sees it.  On the wire (MAPPED pinned host memory, written by the finalize phase's owner
// thread) it is flag-in-data like the rows of the exchange: kHostWords 8-byte stores, each 4 bytes of payload + the low 32 bits
// of the step's sequence number -- an aligned 8-byte store is single-copy atomic, so the host needs no fence on the device side
// (no __threadfence_system round trip over PCIe before a separate `seq` word): it polls until every word carries the
// sequence number it waits for, then decodes.
constexpr int kHostWords = 17;   // out_u 4, out_x 6, max_dev 2, head 2, status 1, candidates 1, overflow_total 1
struct HostWire {
  unsigned long long w[kHostWords];
};
struct HostResult {
  double out_u[2];
  double out_x[3];
  double max_dev;
  double head;           // head-room of the screening window of this step (max_dev is checked against half of it)
  int status;
  int candidates;
  int overflow_total;
  int pad;
  unsigned long long seq;
};
// x0 / goal of a step either travel as kernel ARGUMENTS (mppi_step: no H2D copy on the critical path) or live in
// DynState (device-resident closed loop, sharded local/finish steps)
struct StepInput {
  double x0g[6];         // x0[3], goal[3]
  int from_args;
};
__device__ __forceinline__ void load_step_input(const StepInput& in, const DynState* __restrict__ ds, double x0[3], double goal[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    x0[i] = in.from_args ? in.x0g[i] : ds->x0[i];
    goal[i] = in.from_args ? in.x0g[3 + i] : ds->goal[i];
  }
}
constexpr int kStatusRedoF64 = 100;   // MIXED: support list overflowed, host must redo the step in fp64

// ---- small math layer ---------------------------------------------------------------------------
template <typename R> struct Math;

template <> struct Math<float> {
  typedef float4 Vec4;
  static __device__ __forceinline__ float pi() { return 3.14159274101257324f; }
  static __device__ __forceinline__ float inv_2pi() { return 0.159154943091895336f; }
  static __device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
  static __device__ __forceinline__ float ceil_(float a) { return ceilf(a); }
  static __device__ __forceinline__ float floor_(float a) { return floorf(a); }
  static __device__ __forceinline__ float exp_(float a) { return __expf(a); }
  static __device__ __forceinline__ float min_(float a, float b) { return fminf(a, b); }
  static __device__ __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
  // sin/cos for |x| up to ~1e4: 3-term Cody-Waite to [-pi/4, pi/4] + minimax polynomials.
  // No MUFU (keeps the SFU pipe for the Box-Muller of the noise), ~1.5 ulp.
  static __device__ __forceinline__ void sincos_(float x, float& s, float& c) {
    float j = rintf(x * 0.636619772367581343f);
    int q = __float2int_rn(j);
    float r = fmaf(-j, 1.57079601e+00f, x);
    r = fmaf(-j, 3.13916473e-07f, r);
    r = fmaf(-j, 5.39030253e-15f, r);
    float z = r * r;
    float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * r, r);
    float pc = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f),
                    z * z, fmaf(-0.5f, z, 1.0f));
    float ss = (q & 1) ? pc : ps;
    float cc = (q & 1) ? ps : pc;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
  }
  // sin/cos of an angle INCREMENT: polynomials only when |a| <= pi/4 (always, for sane dt * yaw rate)
  static __device__ __noinline__ void sincos_cold_(float x, float& s, float& c) { sincos_(x, s, c); }
  static __device__ __forceinline__ void sincos_small_(float a, float& s, float& c) {
    if (fabsf(a) > 0.78539816f) {
      sincos_cold_(a, s, c);
      return;
    }
    const float z = a * a;
    s = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * a, a);
    c = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z,
             fmaf(-0.5f, z, 1.0f));
  }
  // polynomial sin/cos, |a| <= pi/4 guaranteed by the caller (FAST path: no branch at all)
  static __device__ __forceinline__ void sincos_poly_(float a, float& s, float& c) {
    const float z = a * a;
    s = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * a, a);
    c = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f), z * z,
             fmaf(-0.5f, z, 1.0f));
  }
  // the reference's wrap for |theta| < 3 pi (one turn at most), branch-free: n = +1 above pi, -1 at or
  // below -pi, else 0  ==  ceil((theta+pi)/(2pi)) - 1 on that range (control/src/mppi:52-53)
  static __device__ __forceinline__ float wrap_once_(float th) {
    const float n = (th > pi()) ? 1.0f : ((th <= -pi()) ? -1.0f : 0.0f);
    const float r = fmaf(-n, 6.28318548202514648f, th);
    return fmaf(n, 1.74845553146951715e-07f, r);
  }
  // theta - (ceil((theta+pi)/(2pi)) - 1) * 2pi   (control/src/mppi:52-53), 2pi split hi/lo.
  // The expression is the identity on (-pi, pi], so it is only evaluated outside that interval.
  static __device__ __noinline__ float wrap_slow_(float th) {
    float n = ceilf((th + pi()) * inv_2pi()) - 1.0f;
    float r = fmaf(-n, 6.28318548202514648f, th);
    return fmaf(n, 1.74845553146951715e-07f, r);
  }
  static __device__ __forceinline__ float wrap_(float th) {
    if (th > pi() || th <= -pi()) th = wrap_slow_(th);
    return th;
  }
};

template <> struct Math<double> {
  typedef double4 Vec4;
  static __device__ __forceinline__ double pi() { return 3.14159265358979323846; }
  static __device__ __forceinline__ double inv_2pi() { return 0.15915494309189533577; }
  static __device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }
  static __device__ __forceinline__ double ceil_(double a) { return ceil(a); }
  static __device__ __forceinline__ double floor_(double a) { return floor(a); }
  static __device__ __forceinline__ double exp_(double a) { return exp(a); }
  static __device__ __forceinline__ double min_(double a, double b) { return fmin(a, b); }
  static __device__ __forceinline__ double max_(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  // fdlibm-style kernels on [-pi/4, pi/4] (~1 ulp)
  static __device__ __forceinline__ void sincos_kernel_(double r, double& s, double& c) {
    const double z = r * r;
    const double ps = fma(fma(fma(fma(fma(1.58969099521155010221e-10, z, -2.50507602534068634195e-08), z,
                                      2.75573137070700676789e-06), z, -1.98412698298579493134e-04), z,
                              8.33333333332248946124e-03), z, -1.66666666666666324348e-01);
    const double pc = fma(fma(fma(fma(fma(-1.13596475577881948265e-11, z, 2.08757232129817482790e-09), z,
                                      -2.75573143513906633035e-07), z, 2.48015872894767294178e-05), z,
                              -1.38888888888741095749e-03), z, 4.16666666666666019037e-02);
    s = fma(ps * z, r, r);
    c = fma(pc * z, z, fma(-0.5, z, 1.0));
  }
  // sin/cos for |x| up to ~1e5: two-term FMA Cody-Waite reduction (no Payne-Hanek slow path, which
  // the CUDA library routine drags in together with a local-memory stack frame)
  static __device__ __forceinline__ void sincos_(double x, double& s, double& c) {
    const double j = rint(x * 0.63661977236758138);
    const int q = __double2int_rn(j);
    double r = fma(-j, 1.5707963267948966, x);
    r = fma(-j, 6.123233995736766e-17, r);
    double ps, pc;
    sincos_kernel_(r, ps, pc);
    const double ss = (q & 1) ? pc : ps;
    const double cc = (q & 1) ? ps : pc;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
  }
  static __device__ __forceinline__ void sincos_poly_(double a, double& s, double& c) { sincos_kernel_(a, s, c); }
  static __device__ __forceinline__ double wrap_once_(double th) {
    const double n = (th > pi()) ? 1.0 : ((th <= -pi()) ? -1.0 : 0.0);
    return th - n * 2.0 * pi();
  }
  static __device__ __forceinline__ void sincos_small_(double a, double& s, double& c) {
    if (fabs(a) > 0.78539816339744828) {
      sincos_(a, s, c);
      return;
    }
    sincos_kernel_(a, s, c);
  }
  // same expression as the reference, evaluated in f64 (control/src/mppi:52-53)
  static __device__ __forceinline__ double wrap_(double th) {
    if (th > pi() || th <= -pi()) {   // identity on (-pi, pi]
      double n = ceil((th + pi()) / (2.0 * pi())) - 1.0;
      th = th - n * 2.0 * pi();
    }
    return th;
  }
};

template <typename R>
__device__ __forceinline__ R clamp_(R v, R lim) {
  return Math<R>::max_(-lim, Math<R>::min_(v, lim));
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter-based: the noise of rollout k at the step triple
// `call` of engine step `step` is a pure function of (seed, k_global, call, step) -- independent of how
// K is sharded over GPUs and regenerable anywhere (reduction kernels, noise export).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < MPPI_PHILOX_ROUNDS; ++i) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// 6 standard normals (fp32) per Philox call = the z of the three time steps 3c, 3c+1, 3c+2 (channel 0, channel 1 each),
// c = the call index (counter word z).  The 128 random bits make THREE Box-Muller pairs: pair j takes the 22 high bits
// of word j as the radius uniform (tail to 5.6 sigma, 4 M levels) and 20 angle bits (the 10 low bits of word j over 10 bits
// of word 3) -- fp32 normals carry 24 bits, so nothing an fp32 rollout can see is lost against 32 + 32 bits per pair, and
// the generator (43 % of the rollout loop's issue time at two steps per call) is called a third less often.
// Box-Muller on the SFU pipe (lg2, sqrt, sin, cos).  Bit-identical wherever it is called from (intrinsics only,
// nothing for the compiler to contract).
struct Normal6 {
  float v[6];
};
// one Box-Muller pair from the 32 bits of its own word and its 10 bits of word 3 (already shifted down and masked)
__device__ __forceinline__ void normal_pair_from_bits(uint32_t wj, uint32_t w10, float& z0, float& z1) {
  const uint32_t rb = wj >> 10;                                                            // 22 bits
  const uint32_t ab = ((wj & 0x3FFu) << 10) | w10;                                         // 20 bits
  const float u = __fmaf_rn((float)rb, 2.384185791015625e-07f, 1.1920928955078125e-07f);   // (rb + 1/2) 2^-22 in (0, 1)
  // sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u): MUFU.LG2 + FMUL + MUFU.SQRT
  const float ra = sqrt_approx(__fmul_rn(-1.38629436111989062f, lg2_approx(u)));
  float sn, cs;
  __sincosf(__fmul_rn((float)ab, 5.992112452678286e-06f), &sn, &cs);                       // 2 pi 2^-20
  z0 = __fmul_rn(ra, cs);
  z1 = __fmul_rn(ra, sn);
}
__device__ __forceinline__ Normal6 normal6_from_bits(uint4 r) {
  Normal6 o;
  normal_pair_from_bits(r.x, r.w >> 22, o.v[0], o.v[1]);
  normal_pair_from_bits(r.y, (r.w >> 12) & 0x3FFu, o.v[2], o.v[3]);
  normal_pair_from_bits(r.z, (r.w >> 2) & 0x3FFu, o.v[4], o.v[5]);
  return o;
}
__device__ __forceinline__ Normal6 philox_normal6(unsigned long long seed, unsigned long long kglobal, unsigned int call,
                                                  unsigned int step) {
  uint4 ctr = make_uint4((uint32_t)kglobal, (uint32_t)(kglobal >> 32), call, step);
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return normal6_from_bits(philox4x32_10(ctr, key));
}

// eps = noise_std * z, rounded once in fp32: THE sample value (what mppi_get_noise exports)
__device__ __forceinline__ float eps_from_z(float std_, float z) { return __fmul_rn(std_, z); }

// eps of (rollout kglobal, time t, channel pair) regenerated from counters: only the pair of step t is expanded
__device__ __forceinline__ void philox_eps(unsigned long long seed, unsigned long long kglobal, int t,
                                           unsigned int step, float std0, float std1, float& e0, float& e1) {
  const unsigned int call = (unsigned)t / 3u, j = (unsigned)t - 3u * call;
  const uint4 ctr = make_uint4((uint32_t)kglobal, (uint32_t)(kglobal >> 32), call, step);
  const uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t wj = (j == 0u) ? r.x : ((j == 1u) ? r.y : r.z);
  float z0, z1;
  normal_pair_from_bits(wj, (r.w >> (22u - 10u * j)) & 0x3FFu, z0, z1);
  e0 = eps_from_z(std0, z0);
  e1 = eps_from_z(std1, z1);
}

// ---- vehicle models: every supported model has a state-independent yaw rate, so one step is
//   s = forward speed(u), w = yaw rate(u);  RK4 on (x, y, theta) with u held (control/src/mppi:39-54)
template <typename R>
struct ModelConsts {
  R dt, half_r, r_over_L, inv_L;
  R x0, y0;     // start position of the step: a user ODE (MPPI_MODEL_USER) sees absolute coordinates
};

#if defined(MPPI_USER_MODEL) && !defined(MPPI_USER_KINEMATIC)
// ---- caller-supplied functors (mppi_create_user; the text is compiled together with these headers by NVRTC) ----------
//   template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]);
//   (MPPI_USER_COST)  mppi_user_running_cost<R>(x, goal, u_nom, eps, t),  mppi_user_terminal_cost<R>(x, goal)
// One integrator step of the user's ODE on the absolute state, as the reference's integrator functors do it:
// MPPI_USER_INTEGRATOR 0 = rk4 with the control held (control/src/mppi:39-50; generic RK4 of control/src/rk4.cpp),
// 1 = explicit Euler (:57-58); MPPI_USER_WRAP = the theta wrap of :52-53.
template <typename R>
__device__ __forceinline__ void user_integrate(R dt, const R x[3], const R u[2], R xn[3]) {
#if MPPI_USER_INTEGRATOR == 1
  R k1[3];
  mppi_user_ode<R>(x, u, k1);
#pragma unroll
  for (int i = 0; i < 3; ++i) xn[i] = x[i] + dt * k1[i];
#else
  R k1[3], k2[3], k3[3], k4[3], xt[3];
  mppi_user_ode<R>(x, u, k1);
#pragma unroll
  for (int i = 0; i < 3; ++i) { k1[i] *= dt; xt[i] = x[i] + k1[i] / R(2); }
  mppi_user_ode<R>(xt, u, k2);
#pragma unroll
  for (int i = 0; i < 3; ++i) { k2[i] *= dt; xt[i] = x[i] + k2[i] / R(2); }
  mppi_user_ode<R>(xt, u, k3);
#pragma unroll
  for (int i = 0; i < 3; ++i) { k3[i] *= dt; xt[i] = x[i] + k3[i]; }
  mppi_user_ode<R>(xt, u, k4);
#pragma unroll
  for (int i = 0; i < 3; ++i) xn[i] = x[i] + (R(1.0) / R(6.0)) * (k1[i] + R(2) * k2[i] + R(2) * k3[i] + dt * k4[i]);
#endif
#if MPPI_USER_WRAP
  xn[2] = Math<R>::wrap_(xn[2]);
#endif
}
#endif

// true for the models integrated by explicit Euler without a theta wrap (control/src/mppi:57-58): the reference's unicycle, and a
// caller's kinematic functor that asks for it
template <int MODEL>
struct EulerLike {
#if defined(MPPI_USER_KINEMATIC) && MPPI_USER_INTEGRATOR == 1
  static constexpr bool value = MODEL == MPPI_MODEL_UNICYCLE_EULER || MODEL == MPPI_MODEL_USER;
#else
  static constexpr bool value = MODEL == MPPI_MODEL_UNICYCLE_EULER;
#endif
};

template <typename R, int MODEL>
__device__ __forceinline__ void speed_yaw(const ModelConsts<R>& mc, R u0, R u1, R& s, R& w) {
#ifdef MPPI_USER_KINEMATIC
  // a caller's KINEMATIC functor (mppi_create_user, kind 1): forward speed and yaw rate as functions of the controls --
  //   xdot = s(u) cos(theta), ydot = s(u) sin(theta), thetadot = w(u)
  // the family all built-in models belong to; everything else (integrator, screen, fp64 re-evaluation) is the built-in code
  if (MODEL == MPPI_MODEL_USER) {
    const R u[2] = {u0, u1};
    mppi_user_speed_yaw<R>(u, &s, &w);
    return;
  }
#endif
  if (MODEL == MPPI_MODEL_DIFF_DRIVE) {          // dd_dynamics, control/src/mppi:23-30
    s = mc.half_r * (u0 + u1);
    w = mc.r_over_L * (u1 - u0);
  } else if (MODEL == MPPI_MODEL_UNICYCLE_EULER) {   // unicycle_dynamics, control/src/mppi:33-36
    s = u0;
    w = u1;
  } else {                                        // NEW bicycle: thdot = v tan(delta) / L
    R sd, cd;
    Math<R>::sincos_(u1, sd, cd);
    s = u0;
    w = u0 * (sd / cd) * mc.inv_L;
  }
}

// One integrator step in DISPLACEMENT coordinates (dx, dy relative to x0; theta absolute, wrapped),
// carrying (c, s) = (cos theta, sin theta) so that the three RK4 trig evaluations become one
// small-angle sincos of the half increment plus two plane rotations (angle-addition):
//   RK4: k1 uses theta, k2 == k3 use theta + k/2, k4 uses theta + k  (theta-dot does not depend on the
//   state)  =>  x+ = x + dt*s/6 * (c1 + 4 c2 + c4).
// The caller re-synchronises (c, s) from theta every few steps (resync_trig) to stop rounding drift.
// FAST = the engine has checked on the host that |dt * yaw rate| <= pi/4 for every admissible control, so
// the increment's sin/cos are plain polynomials and one wrap turn suffices: no branch in the step.
template <typename R, int MODEL, bool FAST = false>
__device__ __forceinline__ void model_step(const ModelConsts<R>& mc, R u0, R u1, R& dx, R& dy, R& th, R& c, R& s) {
#if defined(MPPI_USER_MODEL) && !defined(MPPI_USER_KINEMATIC)
  if (MODEL == MPPI_MODEL_USER) {   // the caller's ODE through the generic integrator; (c, s) are not carried
    const R x[3] = {mc.x0 + dx, mc.y0 + dy, th}, u[2] = {u0, u1};
    R xn[3];
    user_integrate<R>(mc.dt, x, u, xn);
    dx = xn[0] - mc.x0;
    dy = xn[1] - mc.y0;
    th = xn[2];
    return;
  }
#endif
  R spd, w;
  speed_yaw<R, MODEL>(mc, u0, u1, spd, w);
  const R kth = mc.dt * w;
  if (EulerLike<MODEL>::value) {                  // euler, control/src/mppi:57-58 (no wrap)
    dx = Math<R>::fma_(mc.dt * spd, c, dx);
    dy = Math<R>::fma_(mc.dt * spd, s, dy);
    th = th + kth;
    R sa, ca;
    if (FAST)
      Math<R>::sincos_poly_(kth, sa, ca);
    else
      Math<R>::sincos_small_(kth, sa, ca);
    const R cn = c * ca - s * sa;
    s = Math<R>::fma_(s, ca, c * sa);
    c = cn;
    return;
  }
  R sa, ca;
  if (FAST)
    Math<R>::sincos_poly_(R(0.5) * kth, sa, ca);
  else
    Math<R>::sincos_small_(R(0.5) * kth, sa, ca);
  const R c2 = Math<R>::fma_(c, ca, -(s * sa)), s2 = Math<R>::fma_(s, ca, c * sa);
  const R c4 = Math<R>::fma_(c2, ca, -(s2 * sa)), s4 = Math<R>::fma_(s2, ca, c2 * sa);
  const R g = mc.dt * spd * R(1.0 / 6.0);
  dx = Math<R>::fma_(g, Math<R>::fma_(R(4), c2, c + c4), dx);
  dy = Math<R>::fma_(g, Math<R>::fma_(R(4), s2, s + s4), dy);
  th = FAST ? Math<R>::wrap_once_(th + kth) : Math<R>::wrap_(th + kth);
  c = c4;
  s = s4;
}

constexpr int kTrigResyncMask = 3;   // (c, s) <- sincos(theta) after every 4th step

// ---- cost in delta form --------------------------------------------------------------------------
// The reference subtracts min_k V[t,k] per t before exponentiating (control/src/mppi:189), so any
// k-independent term of the running cost is irrelevant to the weights.  We drop the control cost
// 1/2 u'Ru (u is the NOMINAL control, :160,183) and the cost of "standing still at x0", i.e. we
// accumulate  1/2 Q (e^2 - a^2) = 1/2 Q d (d + 2a)  with a = x0 - goal, d = displacement.  This keeps
// fp32 magnitudes ~1e2 instead of ~3e4.  mppi_get_cost_to_go adds the offset back.
template <typename R>
struct CostConsts {
  R hqx, hqy, hqth;      // Q/2
  R p1x, p1y, p1th;
  R ax2, ay2;            // 2*(x0 - goal)
  R th0, gth2;           // theta0, 2*goal_theta
  R gx, gy, x0, y0;      // goal / start position (absolute; a user cost functor sees absolute coordinates)
  // grid
  R g_inv_res, g_ox, g_oy, w_obs_100;   // (x0 - origin) folded into g_ox/g_oy
  int gW, gH;
};

template <typename R>
__device__ __forceinline__ R running_cost(const CostConsts<R>& cc, R dx, R dy, R th, R g0, R g1, R e0, R e1) {
  // get_cost, control/src/mppi:180-184:  1/2 (x-g)'Q(x-g) [+ 1/2 u'Ru dropped] + lam * u.sig.eps
  R c = cc.hqx * dx * (dx + cc.ax2);
  c = Math<R>::fma_(cc.hqy * dy, dy + cc.ay2, c);
  if (cc.hqth != R(0)) c = Math<R>::fma_(cc.hqth * (th - cc.th0), th + cc.th0 - cc.gth2, c);
  c = Math<R>::fma_(g0, e0, c);
  c = Math<R>::fma_(g1, e1, c);
  return c;
}

template <typename R>
__device__ __forceinline__ R terminal_cost(const CostConsts<R>& cc, R dx, R dy, R th) {
  // control/src/mppi:165-171: (x_T - g)' P1 (x_T - g), no 1/2, theta difference NOT wrapped
  R c = cc.p1x * dx * (dx + cc.ax2);
  c = Math<R>::fma_(cc.p1y * dy, dy + cc.ay2, c);
  c = Math<R>::fma_(cc.p1th * (th - cc.th0), th + cc.th0 - cc.gth2, c);
  return c;
}

// NEW occupancy-grid term: w_obs * cell/100, outside the map = 100.
template <typename R>
__device__ __forceinline__ R grid_cost(const CostConsts<R>& cc, const signed char* __restrict__ cells, R dx, R dy) {
  R fx = Math<R>::floor_((dx + cc.g_ox) * cc.g_inv_res);
  R fy = Math<R>::floor_((dy + cc.g_oy) * cc.g_inv_res);
  int v = 100;
  if (fx >= R(0) && fy >= R(0) && fx < R(cc.gW) && fy < R(cc.gH)) v = cells[(int)fy * cc.gW + (int)fx];
  return cc.w_obs_100 * R(v);
}

template <typename R>
__device__ __forceinline__ void make_consts(const StaticParams& sp, const StepInput& in, const DynState* __restrict__ ds,
                                            ModelConsts<R>& mc, CostConsts<R>& cc) {
  mc.dt = R(sp.dt);
  mc.half_r = R(sp.wheel_r * 0.5);
  mc.r_over_L = R(sp.wheel_r / sp.wheel_L);
  mc.inv_L = R(1.0 / sp.wheel_L);
  double xs[3], gs[3];
  load_step_input(in, ds, xs, gs);
  const double x0 = xs[0], y0 = xs[1], th0 = xs[2];
  const double gx = gs[0], gy = gs[1], gth = gs[2];
  cc.hqx = R(0.5 * sp.q[0]);
  cc.hqy = R(0.5 * sp.q[1]);
  cc.hqth = R(0.5 * sp.q[2]);
  cc.p1x = R(sp.p1[0]);
  cc.p1y = R(sp.p1[1]);
  cc.p1th = R(sp.p1[2]);
  mc.x0 = R(x0);
  mc.y0 = R(y0);
  cc.gx = R(gx);
  cc.gy = R(gy);
  cc.x0 = R(x0);
  cc.y0 = R(y0);
  cc.ax2 = R(2.0 * (x0 - gx));
  cc.ay2 = R(2.0 * (y0 - gy));
  cc.th0 = R(th0);
  cc.gth2 = R(2.0 * gth);
  cc.g_inv_res = R(sp.g_inv_res);
  cc.g_ox = R(x0 - sp.g_x0);
  cc.g_oy = R(y0 - sp.g_y0);
  cc.w_obs_100 = R(sp.w_obs / 100.0);
  cc.gW = sp.gW;
  cc.gH = sp.gH;
}

// ---- warp helpers --------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ R warp_min(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = Math<R>::min_(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- TMA 1-D bulk copy + mbarrier (global -> shared), sm_90+/sm_100a -----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace mppi
