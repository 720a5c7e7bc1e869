// rollout_kernel.cuh -- kernel 1 of the MPPI step: fused noise -> rollout -> cost -> per-t partials.
//
// Replaces the reference's hot loop 1, MPPI.get_cost2go (control/src/mppi:127-178) including the K*T
// Python calls of get_cost (:158-161,180-184), and the per-rollout half of update_action (:187-196).
//
// One thread integrates one rollout for all T steps (state in registers); a CTA owns tiles of BLOCK
// rollouts and is persistent over tiles.  The only per-(k,t) data that leaves the registers is the
// running prefix cost, kept in a shared-memory tile P[T][BLOCK+1] (row stride BLOCK+1 makes both the
// column-wise writes of the rollout loop and the row-wise reads below bank-conflict free).  After the
// T steps the tile is consumed TRANSPOSED: LANE l of a warp owns row t = 32*chunk + l and walks the
// BLOCK rollouts serially, with cost-to-go V[t,k] = Tot[k] - P[t-1,k] (= sum_{t'>=t} c[t',k],
// control/src/mppi:175) -- 32 time steps are reduced at once with no shuffles.  Nothing of size K*T
// is written to HBM.
//
//   MODE_SOFTMIN: per (CTA, t) online-softmin partial (m, S, N0, N1): m = min V, S = sum e,
//                 N = sum e*eps[t], e = exp(-(V-m)/lam) (control/src/mppi:189-196).  eps is
//                 regenerated from the Philox counters only for the handful of rollouts with e > 0.
//   MODE_SCREEN : per (CTA, t) the rollouts within `margin` of the running minimum (the support of
//                 the softmin) are listed for fp64 re-evaluation by the reduce kernel.
//
// The floor term of the weights (+1e-8 per rollout, :193) needs E[t] = sum_k eps[t,k]; it is
// accumulated exactly in fixed point (2^-20) with one REDUX (warp integer add) per channel and step;
// integer sums make the term independent of the summation order / sharding.
#pragma once
#include "common.cuh"
#include "reduce_kernels_args.h"

namespace mppi {


template <typename R>
__device__ __forceinline__ R load_eps_ext(const double* __restrict__ eps, int t, int c, int K, int k) {
  return R(eps[((size_t)t * 2 + c) * (size_t)K + k]);
}

// SCREEN slow path (rare): the list is full -> evict stale entries (listed against an older, higher
// running minimum), and if more than kMaxCand rollouts really are inside the window halve it until the
// old entries plus this tile's rollouts fit; then rebuild the list.  One lane, sequential.
template <typename R>
__device__ __forceinline__ int screen_tighten(uint2* list, int cnt_old, const R* tot, const R* pre, int block, int tile_base,
                                           R mnew, R& lim) {
  int total = 0;
  for (int it = 0; it < 32; ++it) {
    // first pass keeps the full window: entries listed against an older (higher) running minimum of a
    // persistent CTA are simply stale and get evicted; only if that is not enough the window is halved
    if (it > 0) lim = mnew + (lim - mnew) * R(0.5);
    total = 0;
    for (int i = 0; i < cnt_old; ++i) total += (R(__uint_as_float(list[i].y)) <= lim) ? 1 : 0;
    for (int k = 0; k < block; ++k) total += ((tot[k] - pre[k]) <= lim) ? 1 : 0;
    if (total <= kMaxCand) break;
  }
  if (total > kMaxCand) lim = -Math<R>::inf();   // > kMaxCand exact ties: force the fp64 redo
  int n = 0;
  for (int i = 0; i < cnt_old; ++i) {
    const uint2 e = list[i];
    if (R(__uint_as_float(e.y)) <= lim) list[n++] = e;
  }
  for (int k = 0; k < block; ++k) {
    const R v = tot[k] - pre[k];
    if (v <= lim && n < kMaxCand) list[n++] = make_uint2((unsigned)(tile_base + k), __float_as_uint((float)v));
  }
  return n;
}

// ---- cost tile geometry ---------------------------------------------------------------------------------
// P holds T+1 rows of BLOCK Reals: row 0 is all zeros ("prefix before step 0"), row 1+t the running cost after
// step t (row T = the rollout total).  Rows are padded by 16 bytes: the column-wise writes of the rollout loop
// stay bank-conflict free AND the transposed pass can read 16-byte vectors (lane l reads row 32c+l: with a row
// stride of BLOCK*sizeof(R)+16 bytes the 8 lanes of a quarter-warp phase hit 8 distinct 16-byte bank groups).
template <typename R> struct Vec16;
template <> struct Vec16<float> {
  typedef float4 V;
  static constexpr int N = 4;
  static __device__ __forceinline__ float min_diff(const float4& a, const float4& b) {
    return fminf(fminf(a.x - b.x, a.y - b.y), fminf(a.z - b.z, a.w - b.w));
  }
};
template <> struct Vec16<double> {
  typedef double2 V;
  static constexpr int N = 2;
  static __device__ __forceinline__ double min_diff(const double2& a, const double2& b) { return fmin(a.x - b.x, a.y - b.y); }
};
template <bool SMALL> struct MaskOf { typedef unsigned int type; };
template <> struct MaskOf<false> { typedef unsigned long long type; };
template <typename R, int BLOCK>
struct CostTile {
  static constexpr int PS = BLOCK + 16 / (int)sizeof(R);          // row stride in Reals
  static __host__ __device__ constexpr size_t bytes(int T) { return (size_t)(T + 1) * PS * sizeof(R); }
};

// One row t of the transposed pass over the BLOCK rollouts of a tile (executed by ONE lane): cost-to-go
// V[t,k] = Tot[k] - P[t-1,k], running minimum, then the online soft-min partial (MODE_SOFTMIN) or the
// candidate list of the fp32 screen (MODE_SCREEN).  Two vectorised sweeps: (1) the minimum, (2) a bit mask of
// the 16-byte groups that hold a rollout inside the window; only those groups are then visited element-wise
// (in increasing k, so the partials do not depend on how the sweep is organised).
template <typename R, int MODE, int BLOCK>
__device__ __forceinline__ void transposed_row(const RolloutArgs& a, int t, int tile, int cta, int nCTA, const R* P,
                                               typename Math<R>::Vec4* run, int* ccount, double* ed, bool cost_to_go,
                                               R neg_inv_lam, R margin, float std0, float std1, unsigned int step,
                                               bool direct_out = false, double fsum0 = 0.0, double fsum1 = 0.0) {
  // direct_out (a CTA that owns exactly one tile): the lane stores the (t, CTA) partial straight to global memory,
  // with the floor sums fsum0/1 of the row -- no shared-memory round trip, no epilogue pass
  typedef typename Math<R>::Vec4 Vec4;
  typedef typename Vec16<R>::V V16;
  constexpr int PS = CostTile<R, BLOCK>::PS;
  constexpr int N = Vec16<R>::N;
  constexpr int NG = BLOCK / N;                                   // 16-byte groups per row
  typedef typename MaskOf<(NG <= 32)>::type Mask;
  const StaticParams& sp = a.sp;
  const int T = sp.T;
  const R* tot = P + (size_t)T * PS;                              // row T: rollout totals
  const R* pre = P + (size_t)(cost_to_go ? t : 0) * PS;           // row t = prefix BEFORE step t (row 0 = zeros), :175
  const V16* tot4 = reinterpret_cast<const V16*>(tot);
  const V16* pre4 = reinterpret_cast<const V16*>(pre);
  const int tile_base = tile * BLOCK;
  const int nk = min(BLOCK, sp.K - tile_base);
  // the sweep: one 16-byte vector of totals (broadcast) and one of prefixes per group; the group minima stay in
  // registers (NG <= 64 values), so the window test below touches no memory
  R gmin[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) gmin[g] = Vec16<R>::min_diff(tot4[g], pre4[g]);
  R m = gmin[0];
#pragma unroll
  for (int g = 1; g < NG; ++g) m = Math<R>::min_(m, gmin[g]);
  if (sp.capture) {
    R* vc = reinterpret_cast<R*>(a.vcap) + (size_t)t * sp.K + tile_base;
    for (int k = 0; k < nk; ++k) vc[k] = tot[k] - pre[k];
  }
  if (sp.noise_external) {   // floor sums straight from the replayed noise
    const double* e0p = a.eps_ext + ((size_t)t * 2 + 0) * sp.K + tile_base;
    const double* e1p = a.eps_ext + ((size_t)t * 2 + 1) * sp.K + tile_base;
    double s0 = 0.0, s1 = 0.0;
    for (int k = 0; k < nk; ++k) {
      s0 += e0p[k];
      s1 += e1p[k];
    }
    ed[2 * t] += s0;
    ed[2 * t + 1] += s1;
  }
  Vec4 rr = run[t];
  const R mnew = Math<R>::min_(rr.x, m);
  // which groups hold a rollout with V <= lim?
  //   SOFTMIN: e^-80 ~ 2e-35 is below any rounding   SCREEN: the window [m, m + margin] of the running minimum
  const R lim = (MODE == MODE_SOFTMIN) ? mnew + R(80) / (-neg_inv_lam) * R(1.0001) : mnew + margin;
  Mask hits = 0;
#pragma unroll
  for (int g = 0; g < NG; ++g)
    if (gmin[g] <= lim) hits |= (Mask)1 << g;
  if (MODE == MODE_SOFTMIN) {
    // online softmin: weights relative to the running minimum of this CTA (:189-196)
    R S = R(0), N0 = R(0), N1 = R(0);
    while (hits) {
      const int g = (sizeof(Mask) == 4) ? __ffs((unsigned int)hits) - 1 : __ffsll((unsigned long long)hits) - 1;
      hits &= hits - 1;
      const V16 tv = tot4[g], pv = pre4[g];
      const R* tvp = reinterpret_cast<const R*>(&tv);
      const R* pvp = reinterpret_cast<const R*>(&pv);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int k = g * N + j;
        const R arg = ((tvp[j] - pvp[j]) - mnew) * neg_inv_lam;   // <= 0
        if (arg > R(-80)) {
          const R e = Math<R>::exp_(arg);
          R e0, e1;
          if (sp.noise_external) {
            e0 = load_eps_ext<R>(a.eps_ext, t, 0, sp.K, tile_base + k);
            e1 = load_eps_ext<R>(a.eps_ext, t, 1, sp.K, tile_base + k);
          } else {
            float f0, f1;
            philox_eps(sp.seed, (unsigned long long)(sp.k_offset + tile_base + k), t, step, std0, std1, f0, f1);
            e0 = R(f0);
            e1 = R(f1);
          }
          S += e;
          N0 = Math<R>::fma_(e, e0, N0);
          N1 = Math<R>::fma_(e, e1, N1);
        }
      }
    }
    const R sc = (rr.x == mnew) ? R(1) : Math<R>::exp_((rr.x - mnew) * neg_inv_lam);
    rr.y = Math<R>::fma_(rr.y, sc, S);
    rr.z = Math<R>::fma_(rr.z, sc, N0);
    rr.w = Math<R>::fma_(rr.w, sc, N1);
    rr.x = mnew;
    if (direct_out) {
      const size_t idx = (size_t)t * nCTA + cta;
      reinterpret_cast<Vec4*>(a.part)[idx] = rr;
      reinterpret_cast<double2*>(a.epart)[idx] = make_double2(fsum0, fsum1);
    } else {
      run[t] = rr;
    }
  } else {
    // screen: keep every rollout within the window [m, lim] of the CTA's running minimum m.
    // run[t] = (m, L): L is the tightest limit ever applied, so the list is guaranteed to hold
    // EVERY rollout of this CTA with V <= L.  Normally lim = m + margin; if more than kMaxCand
    // rollouts fall inside, the window is halved until they fit (the reduce kernel checks that L
    // still covers the window of the GLOBAL minimum, else the step is redone in fp64).
    R lim2 = lim;
    uint2* list = a.cand + ((size_t)t * nCTA + cta) * kMaxCand;
    const int cnt_old = ccount[t];
    int cnt = cnt_old;
    while (hits) {
      const int g = (sizeof(Mask) == 4) ? __ffs((unsigned int)hits) - 1 : __ffsll((unsigned long long)hits) - 1;
      hits &= hits - 1;
      const V16 tv = tot4[g], pv = pre4[g];
      const R* tvp = reinterpret_cast<const R*>(&tv);
      const R* pvp = reinterpret_cast<const R*>(&pv);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const int k = g * N + j;
        const R v = tvp[j] - pvp[j];
        if (v <= lim2) {
          if (cnt < kMaxCand) list[cnt] = make_uint2((unsigned)(tile_base + k), __float_as_uint((float)v));
          ++cnt;
        }
      }
    }
    if (cnt > kMaxCand) cnt = screen_tighten<R>(list, cnt_old, tot, pre, BLOCK, tile_base, mnew, lim2);
    rr.x = mnew;
    rr.y = Math<R>::min_(rr.y, lim2);
    if (direct_out) {
      const size_t idx = (size_t)t * nCTA + cta;
      a.cand_meta[idx] = make_float4((float)rr.x, (float)rr.y, __int_as_float(cnt), 0.f);
      reinterpret_cast<double2*>(a.epart)[idx] = make_double2(fsum0, fsum1);
    } else {
      run[t] = rr;
      ccount[t] = cnt;
    }
  }
}

template <typename R, int MODEL, int MODE, bool HAS_GRID, int BLOCK, bool FAST>
__global__ void __launch_bounds__(BLOCK) rollout_kernel(const __grid_constant__ RolloutArgs a) {
  typedef typename Math<R>::Vec4 Vec4;
  constexpr int NW = BLOCK / 32;
  constexpr int PS = CostTile<R, BLOCK>::PS;   // padded row stride of the cost tile
  const StaticParams& sp = a.sp;
  const int T = sp.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nCTA = gridDim.x, cta = blockIdx.x;

  // ---- shared memory carve-up ------------------------------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);                      // 16 B
  R* nomU0 = reinterpret_cast<R*>(smem_raw + 16);                             // 4*T Reals
  R* nomU1 = nomU0 + T;
  R* nomG0 = nomU1 + T;
  R* nomG1 = nomG0 + T;
  size_t off = 16 + (size_t)4 * T * sizeof(R);
  off = (off + 15) & ~(size_t)15;
  Vec4* run = reinterpret_cast<Vec4*>(smem_raw + off);                        // running (m,S,N0,N1) per t
  off += (size_t)T * sizeof(Vec4);
  long long* ez64 = reinterpret_cast<long long*>(smem_raw + off);             // [T][2]
  off += (size_t)T * 2 * sizeof(long long);
  double* ed = reinterpret_cast<double*>(smem_raw + off);                     // [T][2] (external noise)
  off += (size_t)T * 2 * sizeof(double);
  int* ez32 = reinterpret_cast<int*>(smem_raw + off);                         // [T][2] per-tile
  off += (size_t)T * 2 * sizeof(int);
  int* ccount = reinterpret_cast<int*>(smem_raw + off);                       // [T] SCREEN counts
  off += (size_t)T * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  R* P = reinterpret_cast<R*>(smem_raw + off);                                // [T+1][PS], row 0 = zeros
  off += CostTile<R, BLOCK>::bytes(T);
  off = (off + 15) & ~(size_t)15;
  signed char* gcells = reinterpret_cast<signed char*>(smem_raw + off);       // grid copy (optional)

  // ---- prologue: TMA bulk copies of the nominal block (+ grid) into shared memory -------------
  const uint32_t nom_bytes = (uint32_t)(4 * T * sizeof(R));
  const bool grid_smem = HAS_GRID && sp.grid_in_smem;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, nom_bytes + (grid_smem ? (uint32_t)sp.grid_bytes_padded : 0u));
    tma_bulk_g2s(nomU0, a.nom, nom_bytes, bar);
    if (grid_smem) tma_bulk_g2s(gcells, a.grid, (uint32_t)sp.grid_bytes_padded, bar);
  }
  for (int t = tid; t < T; t += BLOCK) {
    Vec4 v;
    v.x = Math<R>::inf();
    v.y = (MODE == MODE_SCREEN) ? Math<R>::inf() : R(0);
    v.z = R(0);
    v.w = R(0);
    run[t] = v;
    ez64[2 * t] = 0;
    ez64[2 * t + 1] = 0;
    ed[2 * t] = 0.0;
    ed[2 * t + 1] = 0.0;
    ez32[2 * t] = 0;
    ez32[2 * t + 1] = 0;
    ccount[t] = 0;
  }
  for (int k = tid; k < PS; k += BLOCK) P[k] = R(0);
  ModelConsts<R> mc;
  CostConsts<R> cc;
  make_consts<R>(sp, a.in, a.dyn, mc, cc);
  const R um0 = R(sp.u_max[0]), um1 = R(sp.u_max[1]);
  const float std0 = (float)a.dyn->noise_std[0], std1 = (float)a.dyn->noise_std[1];
  const unsigned int step = a.dyn->step;
  const R neg_inv_lam = R(-1.0 / a.dyn->lam);
  double xs_[3], gs_[3];
  load_step_input(a.in, a.dyn, xs_, gs_);
  const R margin = R((float)screen_window(sp, xs_, gs_));   // SCREEN window of this step (common.cuh)
  const signed char* cells = grid_smem ? gcells : a.grid;
  const bool cost_to_go = sp.weighting == MPPI_WEIGHT_COST_TO_GO;
  mbar_wait(bar, 0);
  __syncthreads();

  for (int tile = cta; tile < a.ntiles; tile += nCTA) {
    const int k_local = tile * BLOCK + tid;
    const bool valid = k_local < sp.K;
    const unsigned long long kglobal = (unsigned long long)(sp.k_offset + k_local);
    R dx = R(0), dy = R(0), th = cc.th0, acc = R(0);
    R cth, sth;
    Math<R>::sincos_(th, sth, cth);
    int eown0 = 0, eown1 = 0;   // fixed-point floor sums of the step this lane owns in the current 30-step chunk

    // ---- the T-step rollout (hot loop 1, control/src/mppi:136-163) ----------------------------
    // one model step + running cost + prefix store; z0, z1 = the two standard normals of the step (Philox noise);
    // `own` = the step's position in the current 30-step chunk of floor sums
    auto one_step = [&](int t, int own, R e0, R e1, float z0, float z1) {
      if (!sp.noise_external) {
        // floor-term sums: exact fixed point (2^-20, |z| < 8 so 32 lanes fit an int), one warp integer
        // add (REDUX) per channel; lane `own` keeps the warp sums of step t until the chunk is flushed
        int q0 = valid ? __float2int_rn(z0 * (float)kZFixScale) : 0;
        int q1 = valid ? __float2int_rn(z1 * (float)kZFixScale) : 0;
        q0 = __reduce_add_sync(0xffffffffu, q0);
        q1 = __reduce_add_sync(0xffffffffu, q1);
        if (lane == own) {
          eown0 = q0;
          eown1 = q1;
        }
      }
      // u_samp = clip(U[:,t] + eps)   control/src/mppi:147-152 (eps itself stays unclipped)
      const R u0 = clamp_<R>(nomU0[t] + e0, um0);
      const R u1 = clamp_<R>(nomU1[t] + e1, um1);
      model_step<R, MODEL, FAST>(mc, u0, u1, dx, dy, th, cth, sth);          // :154
#if defined(MPPI_USER_MODEL) && defined(MPPI_USER_COST)
      const R xa_[3] = {cc.x0 + dx, cc.y0 + dy, th}, ga_[3] = {cc.gx, cc.gy, cc.gth2 * R(0.5)};
      const R un_[2] = {nomU0[t], nomU1[t]}, ep_[2] = {e0, e1};
      R c = mppi_user_running_cost<R>(xa_, ga_, un_, ep_, t);               // the caller's cost functor (absolute state)
#else
      R c = running_cost<R>(cc, dx, dy, th, nomG0[t], nomG1[t], e0, e1);     // :160-161,180-184
#endif
      if (HAS_GRID) c += grid_cost<R>(cc, cells, dx, dy);
      acc += c;
      P[(t + 1) * PS + tid] = acc;
    };
    // every lane flushes the step it owns in the chunk of (at most 30) steps starting at `base`
    auto flush_chunk = [&](int base) {
      const int town = base + lane;
      if (lane < 30 && town < T) {
        atomicAdd(&ez32[2 * town], eown0);
        atomicAdd(&ez32[2 * town + 1], eown1);
      }
      eown0 = 0;
      eown1 = 0;
    };
    if (!sp.noise_external) {
      // Philox noise: SIX steps per iteration from two generator calls (three steps each), no branch inside a step
      // (FAST), the Philox + Box-Muller chains of the NEXT six steps issued alongside so that the integer / SFU work
      // overlaps the FP chain; (cos, sin) re-synchronised from theta at the end of every iteration.
      Normal6 za = philox_normal6(sp.seed, kglobal, 0u, step);
      Normal6 zb = philox_normal6(sp.seed, kglobal, 1u, step);
      int t6 = 0, own = 0, chunk = 0;
      unsigned int call = 0u;
      for (; t6 + 6 <= T; t6 += 6) {
        const Normal6 z0 = za, z1 = zb;
        call += 2u;
        za = philox_normal6(sp.seed, kglobal, call, step);        // one iteration ahead
        zb = philox_normal6(sp.seed, kglobal, call + 1u, step);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          one_step(t6 + j, own + j, R(eps_from_z(std0, z0.v[2 * j])), R(eps_from_z(std1, z0.v[2 * j + 1])), z0.v[2 * j], z0.v[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < 3; ++j)
          one_step(t6 + 3 + j, own + 3 + j, R(eps_from_z(std0, z1.v[2 * j])), R(eps_from_z(std1, z1.v[2 * j + 1])), z1.v[2 * j],
                   z1.v[2 * j + 1]);
        Math<R>::sincos_(th, sth, cth);
        own += 6;
        if (own == 30) {
          flush_chunk(chunk);
          chunk += 30;
          own = 0;
        }
      }
      // the last T mod 6 steps: za / zb already hold their normals
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        if (t6 + j < T) {
          const Normal6& zz = (j < 3) ? za : zb;
          const int jj = (j < 3) ? j : j - 3;
          one_step(t6 + j, own + j, R(eps_from_z(std0, zz.v[2 * jj])), R(eps_from_z(std1, zz.v[2 * jj + 1])), zz.v[2 * jj], zz.v[2 * jj + 1]);
        }
      }
      if (chunk < T) flush_chunk(chunk);
    } else {
      // replayed noise from HBM (mppi_set_noise): eps arrives in f64, the floor sums are formed in the transposed pass
      for (int t = 0; t < T; ++t) {
        const int kk = valid ? k_local : 0;
        const R e0 = load_eps_ext<R>(a.eps_ext, t, 0, sp.K, kk);
        const R e1 = load_eps_ext<R>(a.eps_ext, t, 1, sp.K, kk);
        one_step(t, 0, e0, e1, 0.f, 0.f);
        if ((t & 3) == 3) Math<R>::sincos_(th, sth, cth);
      }
    }
#if defined(MPPI_USER_MODEL) && defined(MPPI_USER_COST)
    {
      const R xa_[3] = {cc.x0 + dx, cc.y0 + dy, th}, ga_[3] = {cc.gx, cc.gy, cc.gth2 * R(0.5)};
      acc += mppi_user_terminal_cost<R>(xa_, ga_);
    }
#else
    acc += terminal_cost<R>(cc, dx, dy, th);                                 // :165-171
#endif
    if (!valid) acc = Math<R>::inf();
    P[T * PS + tid] = acc;   // row T holds the rollout total Tot[k]
    __syncthreads();

    // ---- transposed pass: lane l of warp w owns row t = 32*(w + NW*i) + l ----------------------
    for (int tb = warp * 32; tb < T; tb += NW * 32) {
      const int t = tb + lane;
      if (t < T)
        transposed_row<R, MODE, BLOCK>(a, t, tile, cta, nCTA, P, run, ccount, ed, cost_to_go, neg_inv_lam, margin, std0, std1, step);
    }
    __syncthreads();
    // fold the per-tile fixed-point sums into the CTA's 64-bit accumulators
    for (int i = tid; i < 2 * T; i += BLOCK) {
      ez64[i] += (long long)ez32[i];
      ez32[i] = 0;
    }
    __syncthreads();
  }

  // ---- epilogue: one partial per (t, CTA) ------------------------------------------------------
  // PDL: let the dependent reduce kernel start launching while the stragglers finish
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int t = tid; t < T; t += BLOCK) {
    const size_t idx = (size_t)t * nCTA + cta;
    if (MODE == MODE_SOFTMIN) {
      reinterpret_cast<Vec4*>(a.part)[idx] = run[t];
    } else {
      a.cand_meta[idx] = make_float4((float)run[t].x, (float)run[t].y, __int_as_float(ccount[t]), 0.f);
    }
    a.epart[2 * idx] = sp.noise_external ? ed[2 * t] : (double)ez64[2 * t];
    a.epart[2 * idx + 1] = sp.noise_external ? ed[2 * t + 1] : (double)ez64[2 * t + 1];
  }
}

// shared memory needed by one CTA of rollout_kernel
template <typename R>
__host__ __device__ inline size_t rollout_smem_bytes(int T, int block, int grid_bytes_padded_in_smem) {
  size_t off = 16 + (size_t)4 * T * sizeof(R);
  off = (off + 15) & ~(size_t)15;
  off += (size_t)T * 4 * sizeof(R);              // run
  off += (size_t)T * 2 * sizeof(long long);
  off += (size_t)T * 2 * sizeof(double);
  off += (size_t)T * 2 * sizeof(int);
  off += (size_t)T * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  off += (size_t)(T + 1) * (block + 16 / sizeof(R)) * sizeof(R);
  off = (off + 15) & ~(size_t)15;
  off += (size_t)grid_bytes_padded_in_smem;
  return off + 128;   // slack for the 128 B alignment of the dynamic segment
}

}  // namespace mppi
