// rollout_lean_kernel.cuh -- the product instantiation of kernel 1 (fp32, in-register Philox noise).
//
// Same contract as rollout_kernel<float, ..., FAST> (rollout_kernel.cuh): noise -> rollout -> cost -> per-t
// partials, identical Philox counters and partial-record formats; the reduce kernels cannot tell the two
// apart.  The rollout kernel is ISSUE bound (one warp instruction per scheduler cycle),
// so this variant is written to minimise executed instructions per state-step under the conditions the host
// has checked for it (engine.cu: try_configure):
//   * Q[2] == 0 (the reference's Q = diag(1e3, 1e3, 0), control/src/mppi:69): theta enters the running cost
//     nowhere and is needed by the terminal cost only (:165-171);
//   * |dt * yaw rate| <= 1/8 for every admissible control: sin/cos of the HALF increment (|a| <= 1/16) are short
//     polynomials (sin: a (1 - a^2/6), phase error 0.008 a^5 < 8e-9 rad per rotation; cos: next term a^6/720 < 1e-10).
// Differences from the generic FAST step:
//   * nominal block interleaved as float4 per t (U0', U1', std0*g0, std1*g1): one broadcast LDS.128 per step; the
//     controls live in CLIP UNITS s = u / (2 u_max) + 1/2, so sample + clip (:147-152) is ONE instruction,
//     s = sat(fmaf(std / (2 u_max), z, U')) (FFMA.SAT instead of FFMA + 2 FMNMX on the half-rate ALU pipe), and the
//     noise term of the cost (:184) is fmaf(std*g, z, .): neither eps nor u is ever materialised;
//   * model constants folded into affine maps of the clip-unit controls: half yaw increment a = A0 s0 + A1 s1 + Ac,
//     Simpson factor g = G0 s0 + G1 s1 + Gc;
//   * positions carried in COST UNITS d' = sqrt(Q/2) d (admission: Q[0] == Q[1] > 0, the reference's Q): the running
//     cost in delta form is d'(d' + a') per axis -- FADD + FFMA instead of FMUL + FADD + FFMA;
//   * theta is not carried at all by the wrapping models (rk4 wraps every step, :52-53, so theta_T = atan2(sin, cos) of
//     the carried pair, once per rollout); the Euler model (no wrap, :57-58) keeps a plain sum;
//   * Simpson weights through the mid-point rotation only: c1 + 4 c2 + c4 = c2 (4 + 2 cos a) because
//     c1 + c4 = 2 c2 cos a -- one rotation feeds the position update, a second one advances (cos, sin);
//   * (cos, sin) is never re-derived from theta: one first-order renormalisation per iteration (six steps) keeps the
//     pair on the unit circle, the phase error stays at rounding level (~1e-6 rad over 128 steps);
//   * floor-term sums: round(z * 2^18) by the magic-number trick (FFMA, no F2I on the SFU-class pipe), the
//     rollout-valid mask folded into the scale; the warp REDUX results (uniform registers) of six steps leave
//     through three 16-byte shared-memory stores of one lane into per-warp slots -- no selects, no atomics; the
//     32 lanes' bias is removed when the slots are summed;
//   * six steps per loop iteration from two Philox calls (three steps = three Box-Muller pairs per call, common.cuh),
//     the calls of the next iteration issued one iteration ahead.
//   (bicycle: additionally |delta| <= u_max[1] <= pi/4, so tan(delta) needs no range reduction.)
// Anything outside those conditions (Q[2] != 0, large yaw increments, replayed noise, fp64) runs rollout_kernel.
#pragma once
#include "rollout_kernel.cuh"

namespace mppi {

constexpr float kLeanFixScale = 262144.0f;        // 2^18: |z| < 6.8 -> |q| < 2^21 (magic-number rounding needs < 2^22)
constexpr float kLeanMagic = 12582912.0f;         // 1.5 * 2^23
constexpr unsigned kLeanBias32 = 0x68000000u;     // 32 * float_as_int(kLeanMagic) = 32 * 0x4B400000 (mod 2^32)

// per-step constants: LeanStatic (host-folded) + what depends on x0 / goal / DynState
struct LeanConsts {
  float A0, A1, Ac, G0, G1, Gc;      // affine maps of the clip-unit controls (see LeanStatic)
  float um0, um0x2, um1, um1x2, bk;  // bicycle only
  float k0, k1;                      // std / (2 u_max)
  float std0, std1, ax2, ay2, th0;   // ax2, ay2 = 2 (x0 - goal) in cost units
};

// (half) yaw increment a and Simpson factor g = scale * dt * speed / 6 (Euler: scale * dt * speed) from the clip-unit
// controls s0, s1 in [0, 1]; only the Euler model carries theta
template <int MODEL>
__device__ __forceinline__ void lean_controls(const LeanConsts& lc, float s0, float s1, float& a, float& g, float& th) {
  if (MODEL == MPPI_MODEL_DIFF_DRIVE) {            // dd_dynamics, control/src/mppi:23-30
    a = fmaf(lc.A1, s1, fmaf(lc.A0, s0, lc.Ac));
    g = fmaf(lc.G1, s1, fmaf(lc.G0, s0, lc.Gc));
  } else if (MODEL == MPPI_MODEL_UNICYCLE_EULER) { // unicycle_dynamics + euler, control/src/mppi:33-36,57-58
    a = fmaf(lc.A1, s1, lc.Ac);                    // the FULL increment: Euler rotates once per step
    g = fmaf(lc.G0, s0, lc.Gc);
    th += a;
  } else {                                         // NEW bicycle: thdot = v tan(delta) / L
    const float v = fmaf(s0, lc.um0x2, -lc.um0), dl = fmaf(s1, lc.um1x2, -lc.um1);
    float sd, cd;
    Math<float>::sincos_poly_(dl, sd, cd);         // |delta| <= u_max[1] <= pi/4 is an admission condition of this kernel
    a = lc.bk * (v * __fdividef(sd, cd));
    g = lc.G0 * v;
  }
}

// Philox4x32-10 with the key schedule read from the kernel arguments (constant-bank operands of the XORs): the same
// function as philox4x32_10(ctr, key) in common.cuh, minus 20 key additions per call
__device__ __forceinline__ uint4 philox4x32_sched(uint4 c, const LeanStatic& ls) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < MPPI_PHILOX_ROUNDS; ++i) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ ls.pkx[i], lo1, hi0 ^ c.w ^ ls.pky[i], lo0);
  }
  return c;
}
__device__ __forceinline__ Normal6 lean_normal6(const LeanStatic& ls, unsigned long long kglobal, unsigned int call, unsigned int step) {
  return normal6_from_bits(philox4x32_sched(make_uint4((uint32_t)kglobal, (uint32_t)(kglobal >> 32), call, step), ls));
}

// NEW occupancy-grid term, same cell as grid_cost<float> in common.cuh (floor, outside = 100) with the
// floor done by the float->int conversion and the range test on unsigned integers; dx, dy, g_ox, g_oy in cost units and
// g_inv_res per cost unit
__device__ __forceinline__ float lean_grid_cost(const CostConsts<float>& cc, const signed char* __restrict__ cells, float dx, float dy) {
  const int ix = __float2int_rd((dx + cc.g_ox) * cc.g_inv_res);
  const int iy = __float2int_rd((dy + cc.g_oy) * cc.g_inv_res);
  int v = 100;
  if ((unsigned)ix < (unsigned)cc.gW && (unsigned)iy < (unsigned)cc.gH) v = cells[iy * cc.gW + ix];
  return cc.w_obs_100 * (float)v;
}

// sin/cos for |a| <= 1/16 (see the header for the error budget)
__device__ __forceinline__ void sincos_tiny(float a, float& s, float& c) {
  const float z = a * a;
  s = a * fmaf(z, -1.6666667e-1f, 1.0f);
  c = fmaf(z, fmaf(z, 4.1666668e-2f, -0.5f), 1.0f);
}

// per-step constants of BOTH launch shapes (per-tile CTAs and the SM-wide kernel): LeanStatic (host-folded) + what depends on
// x0 / goal / DynState.  Returns through `post` (two floats of shared memory, written by thread 0) the constants that are only
// needed AFTER the T-step loop, so that they occupy neither registers nor a stack frame during it.
__device__ __forceinline__ void lean_make_consts(const RolloutArgs& a, const double xs[3], const double gs[3], float std0_f, float std1_f,
                                                 float neg_inv_lam_ld, float* post, LeanConsts& lc, CostConsts<float>& cc) {
  const StaticParams& sp = a.sp;
  const LeanStatic& ls = a.lean;
  lc.A0 = ls.A0;
  lc.A1 = ls.A1;
  lc.Ac = ls.Ac;
  lc.G0 = ls.G0;
  lc.G1 = ls.G1;
  lc.Gc = ls.Gc;
  lc.um0 = ls.um0;
  lc.um0x2 = 2.0f * ls.um0;
  lc.um1 = ls.um1;
  lc.um1x2 = 2.0f * ls.um1;
  lc.bk = ls.bk;
  lc.k0 = std0_f * ls.inv2um0;
  lc.k1 = std1_f * ls.inv2um1;
  lc.std0 = std0_f;
  lc.std1 = std1_f;
  lc.ax2 = (float)(2.0 * (xs[0] - gs[0]) * (double)ls.sq);
  lc.ay2 = (float)(2.0 * (xs[1] - gs[1]) * (double)ls.sq);
  lc.th0 = (float)xs[2];
  cc.hqx = 1.f;
  cc.hqy = 1.f;
  cc.hqth = 0.f;
  cc.p1x = ls.p1x;      // P1 / (Q/2): the terminal cost takes the positions in cost units too
  cc.p1y = ls.p1y;
  cc.p1th = ls.p1th;
  cc.ax2 = lc.ax2;
  cc.ay2 = lc.ay2;
  cc.th0 = lc.th0;
  cc.gth2 = 0.f;        // staged in shared memory (post[0]) until the terminal cost needs it
  if (threadIdx.x == 0) {
    post[0] = (float)(2.0 * gs[2]);
    post[1] = neg_inv_lam_ld;
  }
  cc.g_inv_res = ls.g_inv_res;
  cc.g_ox = (float)((xs[0] - sp.g_x0) * (double)ls.sq);
  cc.g_oy = (float)((xs[1] - sp.g_y0) * (double)ls.sq);
  cc.w_obs_100 = ls.w_obs_100;
  cc.gW = sp.gW;
  cc.gH = sp.gH;
}

// the state a LEAN rollout carries through the T-step loop
struct LeanState {
  float dx, dy, th, acc, cth, sth;
};

// one model step + running cost (control/src/mppi:147-161) of BOTH launch shapes; z0, z1 = the step's standard normals,
// n = the step's nominal block (U0', U1', std0*g0, std1*g1); returns the running prefix cost to be stored in the cost tile
template <int MODEL, bool HAS_GRID>
__device__ __forceinline__ float lean_one_step(const LeanConsts& lc, const CostConsts<float>& cc, const signed char* __restrict__ cells,
                                               float qscale, const float4 n, float z0, float z1, int& q0, int& q1, LeanState& st) {
  // floor-term sums: exact fixed point (2^-18), one warp integer add (REDUX, warp-uniform result) per channel
#ifdef MPPI_EXP_NOREDUX   // measurement only (profiles/variants.py): what the floor-term sums cost -- NOT a product path
  q0 = q1 = 0;
#else
  q0 = __reduce_add_sync(0xffffffffu, __float_as_int(fmaf(z0, qscale, kLeanMagic)));
  q1 = __reduce_add_sync(0xffffffffu, __float_as_int(fmaf(z1, qscale, kLeanMagic)));
#endif
  // u_samp = clip(U[:,t] + eps), eps = std * z, in clip units: one FFMA.SAT each (:147-152; eps itself stays unclipped)
  const float s0 = __saturatef(fmaf(lc.k0, z0, n.x));
  const float s1 = __saturatef(fmaf(lc.k1, z1, n.y));
  float ah, g;
  lean_controls<MODEL>(lc, s0, s1, ah, g, st.th);
  float sa, ca;
  sincos_tiny(ah, sa, ca);
  if (MODEL == MPPI_MODEL_UNICYCLE_EULER) {           // euler, :57-58: position with the OLD heading
    st.dx = fmaf(g, st.cth, st.dx);
    st.dy = fmaf(g, st.sth, st.dy);
    const float cn = fmaf(st.cth, ca, -(st.sth * sa));
    st.sth = fmaf(st.sth, ca, st.cth * sa);
    st.cth = cn;
  } else {                                            // rk4, :39-54 (Simpson in x, y; see header)
    const float c2 = fmaf(st.cth, ca, -(st.sth * sa)), s2 = fmaf(st.sth, ca, st.cth * sa);
    const float h = g * fmaf(2.0f, ca, 4.0f);
    st.dx = fmaf(h, c2, st.dx);
    st.dy = fmaf(h, s2, st.dy);
    st.cth = fmaf(c2, ca, -(s2 * sa));
    st.sth = fmaf(s2, ca, c2 * sa);
  }
  // get_cost in delta form and cost units (:180-184; common.cuh running_cost), increments summed before they meet acc
  float c = n.w * z1;
  c = fmaf(n.z, z0, c);
  c = fmaf(st.dx, st.dx + lc.ax2, c);
  c = fmaf(st.dy, st.dy + lc.ay2, c);
  if (HAS_GRID) c += lean_grid_cost(cc, cells, st.dx, st.dy);
  st.acc += c;
  return st.acc;
}

// keep (cos, sin) on the unit circle: one first-order renormalisation per six steps
__device__ __forceinline__ void lean_renormalise(LeanState& st) {
  const float f = fmaf(fmaf(st.cth, st.cth, st.sth * st.sth), -0.5f, 1.5f);
  st.cth *= f;
  st.sth *= f;
}

#ifdef MPPI_EXP_TIMELINE   // measurement only (profiles/rollout_timeline.py) -- NOT compiled into the product library
#define RTS(slot)                                                                  \
  do {                                                                             \
    if (a.debug_ts && threadIdx.x == 0) {                                          \
      unsigned long long t_;                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                       \
      a.debug_ts[(size_t)blockIdx.x * 8 + (slot)] = t_;                            \
    }                                                                              \
  } while (0)
#else
#define RTS(slot) do { } while (0)
#endif

template <int MODEL, int MODE, bool HAS_GRID, int BLOCK>
__global__ void __launch_bounds__(BLOCK, 512 / BLOCK) rollout_lean_kernel(const __grid_constant__ RolloutArgs a) {
  typedef float R;
  typedef float4 Vec4;
  constexpr int NW = BLOCK / 32;
  constexpr int PS = CostTile<R, BLOCK>::PS;
  const StaticParams& sp = a.sp;
  const int T = sp.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nCTA = gridDim.x, cta = blockIdx.x;

  RTS(0);
#ifdef MPPI_EXP_TIMELINE
  if (a.debug_ts && threadIdx.x == 0) {
    unsigned int smid_;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
    a.debug_ts[(size_t)blockIdx.x * 8 + 7] = smid_;
  }
#endif
  // ---- shared memory carve-up ------------------------------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);                      // 8 B
  float* post = reinterpret_cast<float*>(smem_raw + 8);                       // 2 floats: constants only needed AFTER the T-step
                                                                              // loop live here, not in registers or a stack frame
  float4* nomL = reinterpret_cast<float4*>(smem_raw + 16);                    // [T] (U0, U1, std0*g0, std1*g1)
  size_t off = 16 + (size_t)T * sizeof(float4);
  Vec4* run = reinterpret_cast<Vec4*>(smem_raw + off);                        // running (m,S,N0,N1) / (m,L) per t
  off += (size_t)T * sizeof(Vec4);
  long long* ez64 = reinterpret_cast<long long*>(smem_raw + off);             // [T][2] CTA floor sums
  off += (size_t)T * 2 * sizeof(long long);
  int* ezw = reinterpret_cast<int*>(smem_raw + off);                          // [NW][T][2] per-warp floor sums of the tile
  off += (size_t)NW * T * 2 * sizeof(int);                                    //   (biased, see kLeanBias32; no atomics)
  int* ccount = reinterpret_cast<int*>(smem_raw + off);                       // [T] SCREEN counts
  off += (size_t)T * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  R* P = reinterpret_cast<R*>(smem_raw + off);                                // [T+1][PS], row 0 = zeros
  off += CostTile<R, BLOCK>::bytes(T);
  signed char* gcells = reinterpret_cast<signed char*>(smem_raw + off);       // grid copy (optional)

  // ---- prologue: TMA bulk copies of the nominal block (+ grid) into shared memory -------------
  const uint32_t nom_bytes = (uint32_t)(T * sizeof(float4));
  const bool grid_smem = HAS_GRID && sp.grid_in_smem;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, nom_bytes + (grid_smem ? (uint32_t)sp.grid_bytes_padded : 0u));
    tma_bulk_g2s(nomL, a.nom, nom_bytes, bar);
    if (grid_smem) tma_bulk_g2s(gcells, a.grid, (uint32_t)sp.grid_bytes_padded, bar);
  }
  for (int t = tid; t < T; t += BLOCK) {
    run[t] = make_float4(Math<R>::inf(), (MODE == MODE_SCREEN) ? Math<R>::inf() : 0.f, 0.f, 0.f);
    ez64[2 * t] = 0;
    ez64[2 * t + 1] = 0;
    ccount[t] = 0;
  }
  for (int k = tid; k < PS; k += BLOCK) P[k] = 0.f;
  // every global load of the prologue is issued here, before the barrier wait; the engine-lifetime constants
  // arrive pre-folded in the kernel arguments (LeanStatic)
  const DynState* __restrict__ ds = a.dyn;
  const unsigned int step = ds->step;
  const R neg_inv_lam_ld = ds->neg_inv_lam_f;
  const float std0_f = ds->noise_std_f[0], std1_f = ds->noise_std_f[1];
  double xs[3], gs[3];
  load_step_input(a.in, ds, xs, gs);
  LeanConsts lc;
  CostConsts<R> cc;     // grid / terminal constants in the layout lean_grid_cost / terminal_cost expect, in cost units
  lean_make_consts(a, xs, gs, std0_f, std1_f, neg_inv_lam_ld, post, lc, cc);
  const R margin = (float)screen_window(sp, xs, gs);   // SCREEN window of this step (common.cuh)
  const signed char* cells = grid_smem ? gcells : a.grid;
  const bool cost_to_go = sp.weighting == MPPI_WEIGHT_COST_TO_GO;
  float sth0, cth0;
  Math<R>::sincos_(lc.th0, sth0, cth0);
  int* ezrow = ezw + (size_t)warp * T * 2;   // this warp's [T][2] slots
  RTS(1);
  mbar_wait(bar, 0);
  __syncthreads();
  RTS(2);

  const bool multi_tile = a.ntiles > nCTA;
  // floor sums leave in the 2^-20 units of the generic kernel (kZFixScale), so the reduce kernels see one format
  constexpr double kToGeneric = kZFixScale / (double)kLeanFixScale;
  // floor sum i = 2 t + channel of the tile just finished: the warps' slots minus the 32 lanes' bias each
  auto tile_floor_sum = [&](int i) -> long long {
    long long sum = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) sum += (long long)(int)((unsigned)ezw[(size_t)w * T * 2 + i] - kLeanBias32);
    return sum;
  };
  for (int tile = cta; tile < a.ntiles; tile += nCTA) {
    const int k_local = tile * BLOCK + tid;
    const bool valid = k_local < sp.K;
    const unsigned long long kglobal = (unsigned long long)(sp.k_offset + k_local);
    const float qscale = valid ? kLeanFixScale : 0.f;     // invalid rollouts add exactly 0 to the floor sums
    LeanState st = {0.f, 0.f, lc.th0, 0.f, cth0, sth0};
    R* prow = P + PS + tid;                               // row 1 + t of this rollout's column
    // one model step + running cost + prefix store (control/src/mppi:147-161): lean_one_step, shared with the SM-wide kernel
    auto one_step = [&](const float4 n, float z0, float z1, int& q0, int& q1, int row) {
      prow[row * PS] = lean_one_step<MODEL, HAS_GRID>(lc, cc, cells, qscale, n, z0, z1, q0, q1, st);
    };
    Normal6 za = lean_normal6(a.lean, kglobal, 0u, step);
    Normal6 zb = lean_normal6(a.lean, kglobal, 1u, step);
    int t6 = 0;
    unsigned int call = 0u;
    for (; t6 + 6 <= T; t6 += 6) {   // six steps per iteration: two generator calls of three steps each
      const Normal6 z0 = za, z1 = zb;
      call += 2u;
      za = lean_normal6(a.lean, kglobal, call, step);   // one iteration ahead
      zb = lean_normal6(a.lean, kglobal, call + 1u, step);
      const float4* nl = nomL + t6;
      int4 qa, qb, qc;
      one_step(nl[0], z0.v[0], z0.v[1], qa.x, qa.y, 0);
      one_step(nl[1], z0.v[2], z0.v[3], qa.z, qa.w, 1);
      one_step(nl[2], z0.v[4], z0.v[5], qb.x, qb.y, 2);
      one_step(nl[3], z1.v[0], z1.v[1], qb.z, qb.w, 3);
      one_step(nl[4], z1.v[2], z1.v[3], qc.x, qc.y, 4);
      one_step(nl[5], z1.v[4], z1.v[5], qc.z, qc.w, 5);
      if (lane == 0) {   // the warp sums of the six steps: three 16-byte stores by one lane
        int4* ez4 = reinterpret_cast<int4*>(ezrow + 2 * t6);
        ez4[0] = qa;
        ez4[1] = qb;
        ez4[2] = qc;
      }
      prow += 6 * PS;
      lean_renormalise(st);
    }
    // the last T mod 6 steps: za / zb already hold their normals
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      if (t6 + j < T) {
        const Normal6& zz = (j < 3) ? za : zb;
        const int jj = (j < 3) ? j : j - 3;
        int2 q;
        one_step(nomL[t6 + j], zz.v[2 * jj], zz.v[2 * jj + 1], q.x, q.y, j);
        if (lane == 0) *reinterpret_cast<int2*>(ezrow + 2 * (t6 + j)) = q;
      }
    }
    RTS(3);
    // rk4 wraps theta into (-pi, pi] after every step (:52-53): theta_T is the angle of the carried pair
    R th = st.th, acc = st.acc;
    if (MODEL != MPPI_MODEL_UNICYCLE_EULER) th = atan2f(st.sth, st.cth);
    cc.gth2 = post[0];
    acc += terminal_cost<R>(cc, st.dx, st.dy, th);                           // :165-171
    if (!valid) acc = Math<R>::inf();
    // PDL: the reduce kernel may start getting resident now (it still waits for this grid to complete)
    if (tile + nCTA >= a.ntiles) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    RTS(4);
    prow[(T - 1 - t6) * PS] = acc;   // row T holds the rollout total Tot[k] (addressed from the running row pointer: the
                                     // tile-invariant form P + T*PS + tid gets hoisted and spilled)
    __syncthreads();
    RTS(5);

    // ---- transposed pass: lane l of warp w owns row t = 32*(w + NW*i) + l ----------------------
    const R neg_inv_lam = post[1];
    for (int tb = warp * 32; tb < T; tb += NW * 32) {
      const int t = tb + lane;
      if (t < T) {
        const bool direct = !multi_tile;
        const double f0 = direct ? (double)tile_floor_sum(2 * t) * kToGeneric : 0.0;
        const double f1 = direct ? (double)tile_floor_sum(2 * t + 1) * kToGeneric : 0.0;
        transposed_row<R, MODE, BLOCK>(a, t, tile, cta, nCTA, P, run, ccount, nullptr, cost_to_go, neg_inv_lam, margin, lc.std0,
                                       lc.std1, step, direct, f0, f1);
      }
    }
    RTS(6);
    if (multi_tile) {   // a persistent CTA folds its per-tile sums into 64-bit accumulators
      __syncthreads();
      for (int i = tid; i < 2 * T; i += BLOCK) ez64[i] += tile_floor_sum(i);
      __syncthreads();
    }
  }
  if (!multi_tile) return;   // one tile per CTA: every row's partial already went out from the transposed pass

  // ---- epilogue of a persistent CTA: one partial per (t, CTA) ---------------------------------------
  for (int t = tid; t < T; t += BLOCK) {
    const size_t idx = (size_t)t * nCTA + cta;
    if (MODE == MODE_SOFTMIN)
      reinterpret_cast<Vec4*>(a.part)[idx] = run[t];
    else
      a.cand_meta[idx] = make_float4(run[t].x, run[t].y, __int_as_float(ccount[t]), 0.f);
    a.epart[2 * idx] = (double)ez64[2 * t] * kToGeneric;
    a.epart[2 * idx + 1] = (double)ez64[2 * t + 1] * kToGeneric;
  }
}

// ---- SM-wide balanced variant (one CTA of 16 warps per SM) -----------------------------------------------------------
// The T-step loop is issue bound and a warp's rate is 1 / (warps on its scheduler), so the kernel ends when the
// busiest scheduler does.  BASELINE config 2 puts 65536 / 148 = 13.8 warps on an SM: with independent 64-thread CTAs
// the four schedulers hold 4, 4, 3, 3 warps and the two with three idle a quarter of the time (measured with the
// %globaltimer stamps of profiles/rollout_timeline.py: loop 12.2 us on the 3-warp schedulers, 16.7 us on the 4-warp
// ones).  Here one CTA owns the whole SM and the 14th/13th warp's work is CUT IN TIME: warps 0-11 (three per
// scheduler) roll six tiles of 64 rollouts as before; the seventh tile is rolled by warps 12, 13 (schedulers 0, 1) for
// the steps [0, split) and by warps 14, 15 (schedulers 2, 3) for [split, T), the rollout state (6 floats) handed over
// through shared memory under a named barrier.  Every scheduler now carries 3.5 warps' worth of steps.
// Each tile keeps its own cost tile / partial record, so the reduce kernels see 7 partials per CTA in the usual format.
constexpr int kSmFullTiles = 6;
constexpr int kSmTiles = kSmFullTiles + 1;
constexpr int kSmThreads = 64 * (kSmFullTiles + 2);   // 512

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__host__ __device__ inline size_t rollout_lean_sm_tile_bytes(int T) {   // per-tile shared memory of the SM-wide kernel
  size_t off = (size_t)T * sizeof(float4);          // run
  off += (size_t)2 * T * 2 * sizeof(int);           // per-warp floor sums
  off += (size_t)T * sizeof(int);                   // SCREEN counts
  off = (off + 15) & ~(size_t)15;
  off += (size_t)(T + 1) * (64 + 4) * sizeof(float);   // cost tile
  return (off + 15) & ~(size_t)15;
}

template <int MODEL, int MODE, bool HAS_GRID>
__global__ void __launch_bounds__(kSmThreads, 1) rollout_lean_sm_kernel(const __grid_constant__ RolloutArgs a) {
  typedef float R;
  typedef float4 Vec4;
  constexpr int BLOCK = 64;
  constexpr int PS = CostTile<R, BLOCK>::PS;
  const StaticParams& sp = a.sp;
  const int T = sp.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // role 0: a full rollout; role 1: steps [0, split) of the shared tile; role 2: steps [split, T) of the shared tile
  const int sub = (warp < 2 * kSmFullTiles) ? (warp >> 1) : kSmFullTiles;
  const int role = (warp < 2 * kSmFullTiles) ? 0 : ((warp < 2 * kSmFullTiles + 2) ? 1 : 2);
  const int wslot = warp & 1;
  const int col = wslot * 32 + lane;                       // this thread's rollout within its tile
  const int vtile = blockIdx.x * kSmTiles + sub;           // tile index == partial-record index
  const int nvt = a.ntiles;                                // tiles that hold rollouts = ceil(K / 64); the last CTA may own fewer than 7

#ifdef MPPI_EXP_TIMELINE
#define RTS2(slot)                                                                 \
  do {                                                                             \
    if (a.debug_ts && col == 0 && role != 1) {                                     \
      unsigned long long t_;                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                       \
      a.debug_ts[(size_t)vtile * 8 + (slot)] = t_;                                 \
    }                                                                              \
  } while (0)
  if (a.debug_ts && col == 0 && role != 1) {
    unsigned int smid_;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
    a.debug_ts[(size_t)vtile * 8 + 7] = smid_;
  }
#else
#define RTS2(slot) do { } while (0)
#endif
  RTS2(0);
  // ---- shared memory carve-up ------------------------------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float* post = reinterpret_cast<float*>(smem_raw + 8);
  float4* nomL = reinterpret_cast<float4*>(smem_raw + 16);
  size_t off = 16 + (size_t)T * sizeof(float4);
  float* hand = reinterpret_cast<float*>(smem_raw + off);                     // [6][64] hand-over of the shared tile
  off += (size_t)6 * BLOCK * sizeof(float);
  const size_t tile_bytes = rollout_lean_sm_tile_bytes(T);
  unsigned char* tbase = smem_raw + off + (size_t)sub * tile_bytes;
  Vec4* run = reinterpret_cast<Vec4*>(tbase);
  int* ezw = reinterpret_cast<int*>(tbase + (size_t)T * sizeof(Vec4));        // [2][T][2]
  int* ccount = ezw + 2 * T * 2;
  R* P = reinterpret_cast<R*>(tbase + ((((size_t)T * sizeof(Vec4) + (size_t)2 * T * 2 * sizeof(int) + (size_t)T * sizeof(int)) + 15) & ~(size_t)15));
  signed char* gcells = reinterpret_cast<signed char*>(smem_raw + off + (size_t)kSmTiles * tile_bytes);

  // ---- prologue ---------------------------------------------------------------------------------
  const uint32_t nom_bytes = (uint32_t)(T * sizeof(float4));
  const bool grid_smem = HAS_GRID && sp.grid_in_smem;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, nom_bytes + (grid_smem ? (uint32_t)sp.grid_bytes_padded : 0u));
    tma_bulk_g2s(nomL, a.nom, nom_bytes, bar);
    if (grid_smem) tma_bulk_g2s(gcells, a.grid, (uint32_t)sp.grid_bytes_padded, bar);
  }
  if (role != 1) {   // the tile's 64 finishing threads initialise its records (the shared tile: warps 14, 15)
    for (int t = col; t < T; t += BLOCK) {
      run[t] = make_float4(Math<R>::inf(), (MODE == MODE_SCREEN) ? Math<R>::inf() : 0.f, 0.f, 0.f);
      ccount[t] = 0;
    }
    for (int k = col; k < PS; k += BLOCK) P[k] = 0.f;
  }
  const DynState* __restrict__ ds = a.dyn;
  const unsigned int step = ds->step;
  const R neg_inv_lam_ld = ds->neg_inv_lam_f;
  const float std0_f = ds->noise_std_f[0], std1_f = ds->noise_std_f[1];
  double xs[3], gs[3];
  load_step_input(a.in, ds, xs, gs);
  LeanConsts lc;
  CostConsts<R> cc;
  lean_make_consts(a, xs, gs, std0_f, std1_f, neg_inv_lam_ld, post, lc, cc);
  const R margin = (float)screen_window(sp, xs, gs);   // SCREEN window of this step (common.cuh)
  const signed char* cells = grid_smem ? gcells : a.grid;
  const bool cost_to_go = sp.weighting == MPPI_WEIGHT_COST_TO_GO;
  int split = (a.lean.split / 6) * 6;                      // first step of the second part (a multiple of 6: whole iterations)
  split = max(6, min(split, ((T - 1) / 6) * 6));
  const int t_begin = (role == 2) ? split : 0;
  const int t_end = (role == 1) ? split : T;
  const int k_local = vtile * BLOCK + col;
  const bool valid = k_local < sp.K;
  const unsigned long long kglobal = (unsigned long long)(sp.k_offset + k_local);
  const float qscale = valid ? kLeanFixScale : 0.f;
  // the first normals of this warp's range, drawn before anything is waited for
  unsigned int call = (unsigned)t_begin / 3u;
  Normal6 za = lean_normal6(a.lean, kglobal, call, step);
  Normal6 zb = lean_normal6(a.lean, kglobal, call + 1u, step);
  LeanState st = {0.f, 0.f, lc.th0, 0.f, 0.f, 0.f};
  Math<R>::sincos_(lc.th0, st.sth, st.cth);
  int* ezrow = ezw + (size_t)wslot * T * 2;
  RTS2(1);
  mbar_wait(bar, 0);
  __syncthreads();
  RTS2(2);
  if (vtile >= nvt) {   // a tile past the last rollout (last CTA only): nothing to roll, no record to write
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    return;
  }
  if (role == 2) {   // take over the rollouts warps 12, 13 have advanced to `split`
    named_bar_sync(1, 128);
    st.dx = hand[0 * BLOCK + col];
    st.dy = hand[1 * BLOCK + col];
    st.cth = hand[2 * BLOCK + col];
    st.sth = hand[3 * BLOCK + col];
    st.acc = hand[4 * BLOCK + col];
    st.th = hand[5 * BLOCK + col];
  }
  R* prow = P + (size_t)(1 + t_begin) * PS + col;          // row 1 + t of this rollout's column

  auto one_step = [&](const float4 n, float z0, float z1, int& q0, int& q1, int row) {   // the same step as the per-tile kernel
    prow[row * PS] = lean_one_step<MODEL, HAS_GRID>(lc, cc, cells, qscale, n, z0, z1, q0, q1, st);
  };
  int t6 = t_begin;
  for (; t6 + 6 <= t_end; t6 += 6) {
    const Normal6 z0 = za, z1 = zb;
    call += 2u;
    za = lean_normal6(a.lean, kglobal, call, step);   // one iteration ahead
    zb = lean_normal6(a.lean, kglobal, call + 1u, step);
    const float4* nl = nomL + t6;
    int4 qa, qb, qc;
    one_step(nl[0], z0.v[0], z0.v[1], qa.x, qa.y, 0);
    one_step(nl[1], z0.v[2], z0.v[3], qa.z, qa.w, 1);
    one_step(nl[2], z0.v[4], z0.v[5], qb.x, qb.y, 2);
    one_step(nl[3], z1.v[0], z1.v[1], qb.z, qb.w, 3);
    one_step(nl[4], z1.v[2], z1.v[3], qc.x, qc.y, 4);
    one_step(nl[5], z1.v[4], z1.v[5], qc.z, qc.w, 5);
    if (lane == 0) {
      int4* ez4 = reinterpret_cast<int4*>(ezrow + 2 * t6);
      ez4[0] = qa;
      ez4[1] = qb;
      ez4[2] = qc;
    }
    prow += 6 * PS;
    lean_renormalise(st);
  }
  if (role == 1) {   // hand the state to warps 14, 15 and leave
    hand[0 * BLOCK + col] = st.dx;
    hand[1 * BLOCK + col] = st.dy;
    hand[2 * BLOCK + col] = st.cth;
    hand[3 * BLOCK + col] = st.sth;
    hand[4 * BLOCK + col] = st.acc;
    hand[5 * BLOCK + col] = st.th;
    __threadfence_block();
    named_bar_arrive(1, 128);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    return;
  }
  // the last T mod 6 steps: za / zb already hold their normals
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    if (t6 + j < T) {
      const Normal6& zz = (j < 3) ? za : zb;
      const int jj = (j < 3) ? j : j - 3;
      int2 q;
      one_step(nomL[t6 + j], zz.v[2 * jj], zz.v[2 * jj + 1], q.x, q.y, j);
      if (lane == 0) *reinterpret_cast<int2*>(ezrow + 2 * (t6 + j)) = q;
    }
  }
  RTS2(3);
  R th = st.th, acc = st.acc;
  if (MODEL != MPPI_MODEL_UNICYCLE_EULER) th = atan2f(st.sth, st.cth);
  cc.gth2 = post[0];
  acc += terminal_cost<R>(cc, st.dx, st.dy, th);
  if (!valid) acc = Math<R>::inf();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  RTS2(4);
  prow[(T - 1 - t6) * PS] = acc;                             // row T: the rollout total
  named_bar_sync(2 + sub, BLOCK);                            // the tile's two finishing warps
  RTS2(5);

  // ---- transposed pass of this tile: lane l of the tile's warp w owns row t = 32 w + l (+ 64 i) ----
  constexpr double kToGeneric = kZFixScale / (double)kLeanFixScale;
  const R neg_inv_lam = post[1];
  for (int tb = wslot * 32; tb < T; tb += BLOCK) {
    const int t = tb + lane;
    if (t < T) {
      const long long f0 = (long long)(int)((unsigned)ezw[2 * t] - kLeanBias32) + (long long)(int)((unsigned)ezw[2 * T + 2 * t] - kLeanBias32);
      const long long f1 = (long long)(int)((unsigned)ezw[2 * t + 1] - kLeanBias32) + (long long)(int)((unsigned)ezw[2 * T + 2 * t + 1] - kLeanBias32);
      transposed_row<R, MODE, BLOCK>(a, t, vtile, vtile, nvt, P, run, ccount, nullptr, cost_to_go, neg_inv_lam, margin, lc.std0, lc.std1,
                                     step, true, (double)f0 * kToGeneric, (double)f1 * kToGeneric);
    }
  }
  RTS2(6);
}

inline size_t rollout_lean_sm_smem_bytes(int T, int grid_bytes_padded_in_smem) {
  size_t off = 16 + (size_t)T * sizeof(float4) + (size_t)6 * 64 * sizeof(float);
  off += (size_t)kSmTiles * rollout_lean_sm_tile_bytes(T);
  off += (size_t)grid_bytes_padded_in_smem;
  return off + 128;
}

inline size_t rollout_lean_smem_bytes(int T, int block, int grid_bytes_padded_in_smem) {
  size_t off = 16 + (size_t)T * sizeof(float4);
  off += (size_t)T * sizeof(float4);               // run
  off += (size_t)T * 2 * sizeof(long long);
  off += (size_t)(block / 32) * T * 2 * sizeof(int);
  off += (size_t)T * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  off += (size_t)(T + 1) * (block + 4) * sizeof(float);
  off += (size_t)grid_bytes_padded_in_smem;
  return off + 128;
}

}  // namespace mppi
