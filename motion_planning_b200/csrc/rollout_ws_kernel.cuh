// rollout_ws_kernel.cuh -- warp-specialised variant of the fused rollout kernel (fp32, FAST path only).
//
// Why: at K = 65536 a B200 holds only 3.5 rollout warps per scheduler, each executing a ~150-instruction
// step whose critical path (rotate -> accumulate -> cost) is a long dependent chain; issue slots sit idle
// on fixed-latency stalls.  Everything in a step that does NOT depend on the state -- the Philox/Box-Muller
// noise, the clipped control, forward speed and yaw increment, the polynomial sin/cos of the half
// increment, the noise term of the cost -- is therefore moved to two PRODUCER warps per 32 rollouts, and the
// state-dependent recurrence stays in one CONSUMER warp.  Three warps of ~50 instructions per step each
// instead of one warp of ~150: the same work with 3x the thread-level parallelism.
//
//   CTA = 96 threads = 1 tile of 32 rollouts (lane <-> rollout):
//     warp 0   consumer : per step LDS (ca, sa, g, kth, cn) -> 2 plane rotations -> dx, dy, theta (wrap) ->
//                         running cost -> prefix tile; afterwards the transposed pass of rollout_kernel.cuh
//     warp 1,2 producers: producer p owns the step pairs g = p, p+2, ... (one Philox call each),
//                         Box-Muller, floor-sum REDUX, clip, speed/yaw, sincos polynomial, noise cost -> STS
//   hand-over through one shared-memory buffer per producer guarded by a full/empty mbarrier pair (the two
//   producers alternate, which double-buffers the consumer).
//
// Semantics, partial-record formats and Philox counters are identical to rollout_kernel<float,...,FAST>;
// the reduce kernels cannot tell the two apart.
#pragma once
#include "rollout_kernel.cuh"

namespace mppi {

constexpr int kWsGroup = 2;      // steps per hand-over = one Philox call (4 normals); keeps the buffers at 2.5 KB per tile
constexpr int kWsFields = 5;     // ca, sa, g, kth, cn

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

inline size_t rollout_ws_smem_bytes(int T, int grid_bytes_padded_in_smem) {
  size_t off = 64;                                              // 5 mbarriers
  off += (size_t)4 * T * sizeof(float);                         // nominal block
  off = (off + 15) & ~(size_t)15;
  off += (size_t)T * sizeof(float4);                            // run
  off += (size_t)T * 2 * sizeof(unsigned long long);            // ez64
  off += (size_t)T * sizeof(int);                               // ccount
  off = (off + 15) & ~(size_t)15;
  off += (size_t)T * (kWsTile + 1) * sizeof(float);             // prefix tile
  off = (off + 15) & ~(size_t)15;
  off += (size_t)2 * kWsGroup * kWsFields * 32 * sizeof(float); // hand-over buffers
  off = (off + 15) & ~(size_t)15;
  off += (size_t)grid_bytes_padded_in_smem;
  return off;
}

template <int MODEL, int MODE, bool HAS_GRID>
__global__ void __launch_bounds__(kWsThreads, 14) rollout_ws_kernel(const __grid_constant__ RolloutArgs a) {
  typedef float R;
  typedef float4 Vec4;
  constexpr int PS = kWsTile + 1;
  const StaticParams& sp = a.sp;
  const int T = sp.T;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nCTA = gridDim.x, cta = blockIdx.x;
  const int G = T / kWsGroup;                                   // T is even (checked by mppi_create)

  // ---- shared memory carve-up ------------------------------------------------------------------
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* bar_full = bar_tma + 1;                             // [2]
  uint64_t* bar_empty = bar_tma + 3;                            // [2]
  R* nomU0 = reinterpret_cast<R*>(smem_raw + 64);
  R* nomU1 = nomU0 + T;
  R* nomG0 = nomU1 + T;
  R* nomG1 = nomG0 + T;
  size_t off = 64 + (size_t)4 * T * sizeof(R);
  off = (off + 15) & ~(size_t)15;
  Vec4* run = reinterpret_cast<Vec4*>(smem_raw + off);
  off += (size_t)T * sizeof(Vec4);
  unsigned long long* ez64 = reinterpret_cast<unsigned long long*>(smem_raw + off);
  off += (size_t)T * 2 * sizeof(unsigned long long);
  int* ccount = reinterpret_cast<int*>(smem_raw + off);
  off += (size_t)T * sizeof(int);
  off = (off + 15) & ~(size_t)15;
  R* P = reinterpret_cast<R*>(smem_raw + off);
  off += (size_t)T * PS * sizeof(R);
  off = (off + 15) & ~(size_t)15;
  R* ring = reinterpret_cast<R*>(smem_raw + off);               // [2][kWsGroup][kWsFields][32]
  off += (size_t)2 * kWsGroup * kWsFields * 32 * sizeof(R);
  off = (off + 15) & ~(size_t)15;
  signed char* gcells = reinterpret_cast<signed char*>(smem_raw + off);

  // ---- prologue ---------------------------------------------------------------------------------
  const uint32_t nom_bytes = (uint32_t)(4 * T * sizeof(R));
  const bool grid_smem = HAS_GRID && sp.grid_in_smem;
  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], 1);
    mbar_init(&bar_empty[1], 1);
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar_tma, nom_bytes + (grid_smem ? (uint32_t)sp.grid_bytes_padded : 0u));
    tma_bulk_g2s(nomU0, a.nom, nom_bytes, bar_tma);
    if (grid_smem) tma_bulk_g2s(gcells, a.grid, (uint32_t)sp.grid_bytes_padded, bar_tma);
  }
  for (int t = tid; t < T; t += kWsThreads) {
    run[t] = make_float4(__int_as_float(0x7f800000), (MODE == MODE_SCREEN) ? __int_as_float(0x7f800000) : 0.f, 0.f, 0.f);
    ez64[2 * t] = 0ull;
    ez64[2 * t + 1] = 0ull;
    ccount[t] = 0;
  }
  ModelConsts<R> mc;
  CostConsts<R> cc;
  make_consts<R>(sp, a.dyn, mc, cc);
  const R um0 = R(sp.u_max[0]), um1 = R(sp.u_max[1]);
  const float std0 = (float)a.dyn->noise_std[0], std1 = (float)a.dyn->noise_std[1];
  const unsigned int step = a.dyn->step;
  const R neg_inv_lam = R(-1.0 / a.dyn->lam);
  const R margin = R(sp.margin);
  const signed char* cells = grid_smem ? gcells : a.grid;
  const bool cost_to_go = sp.weighting == MPPI_WEIGHT_COST_TO_GO;
  mbar_wait(bar_tma, 0);
  __syncthreads();

  unsigned int n_use = 0;   // hand-overs this warp has taken part in on ITS buffer (producers) / per buffer (consumer)
  unsigned int n_cons[2] = {0u, 0u};

  for (int tile = cta; tile < a.ntiles; tile += nCTA) {
    const int k_local = tile * kWsTile + lane;
    const bool valid = k_local < sp.K;
    if (warp > 0) {
      // =============================== PRODUCER p = warp - 1 =====================================
      const int p = warp - 1;
      R* buf = ring + (size_t)p * kWsGroup * kWsFields * 32;
      const unsigned long long kglobal = (unsigned long long)(sp.k_offset + k_local);
      int eown0 = 0, eown1 = 0;
      for (int g = p; g < G; g += 2) {
        const int t0 = g * kWsGroup;
        static_assert(kWsGroup == 2, "one Philox call = 4 normals = 2 steps x 2 channels per hand-over");
        const float4 za = philox_normal4(sp.seed, kglobal, (unsigned)g, step);
        const float zz[4] = {za.x, za.y, za.z, za.w};
        R out[kWsGroup][kWsFields];
#pragma unroll
        for (int s = 0; s < kWsGroup; ++s) {
          const int t = min(t0 + s, T - 1);                     // T is even, so t0 + s < T always; keeps loads in range
          const float z0 = zz[2 * s], z1 = zz[2 * s + 1];
          // floor-term sums: exact fixed point, one REDUX per channel; lane (t mod 32) keeps step t
          int q0 = valid ? __float2int_rn(z0 * (float)kZFixScale) : 0;
          int q1 = valid ? __float2int_rn(z1 * (float)kZFixScale) : 0;
          q0 = __reduce_add_sync(0xffffffffu, q0);
          q1 = __reduce_add_sync(0xffffffffu, q1);
          if (lane == (t & 31)) {
            eown0 = q0;
            eown1 = q1;
          }
          const R e0 = eps_from_z(std0, z0), e1 = eps_from_z(std1, z1);
          const R u0 = clamp_<R>(nomU0[t] + e0, um0);           // control/src/mppi:147-152
          const R u1 = clamp_<R>(nomU1[t] + e1, um1);
          R spd, w;
          speed_yaw<R, MODEL>(mc, u0, u1, spd, w);
          const R kth = mc.dt * w;
          Math<R>::sincos_poly_((MODEL == MPPI_MODEL_UNICYCLE_EULER) ? kth : R(0.5) * kth, out[s][1], out[s][0]);
          out[s][2] = (MODEL == MPPI_MODEL_UNICYCLE_EULER) ? mc.dt * spd : mc.dt * spd * R(1.0 / 6.0);
          out[s][3] = kth;
          out[s][4] = Math<R>::fma_(nomG1[t], e1, nomG0[t] * e0);   // lam * u.sig.eps, control/src/mppi:184
        }
        // everything above ran while the consumer may still be reading this buffer: wait only now
        mbar_wait(&bar_empty[p], (n_use & 1u) ^ 1u);            // passes at once the first time
#pragma unroll
        for (int s = 0; s < kWsGroup; ++s)
#pragma unroll
          for (int f = 0; f < kWsFields; ++f) buf[((size_t)s * kWsFields + f) * 32 + lane] = out[s][f];
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[p]);
        ++n_use;
        // flush the floor sums when this producer has finished its share of a 32-step chunk (or of the
        // tile): lane l owns step base + l, and producer p holds the steps of the groups with (g & 1) == p
        constexpr int kGroupsPerChunk = 32 / kWsGroup;
        if ((g & (kGroupsPerChunk - 1)) == kGroupsPerChunk - 2 + p || g + 2 >= G) {
          const int town = (t0 & ~31) + lane;
          if ((((lane / kWsGroup) & 1) == p) && town < T) {
            atomicAdd(&ez64[2 * town], (unsigned long long)(long long)eown0);
            atomicAdd(&ez64[2 * town + 1], (unsigned long long)(long long)eown1);
          }
        }
      }
    } else {
      // ==================================== CONSUMER ==============================================
      R dx = R(0), dy = R(0), th = cc.th0, acc = R(0), cth, sth;
      Math<R>::sincos_(th, sth, cth);
      for (int g = 0; g < G; ++g) {
        const int p = g & 1;
        const R* buf = ring + (size_t)p * kWsGroup * kWsFields * 32;
        mbar_wait(&bar_full[p], n_cons[p] & 1u);
        const int t0 = g * kWsGroup;
#pragma unroll
        for (int s = 0; s < kWsGroup; ++s) {
          const int t = t0 + s;
          {
            const R* o = buf + (size_t)s * kWsFields * 32 + lane;
            const R ca = o[0], sa = o[32], gfac = o[64], kth = o[96], cn = o[128];
            if (MODEL == MPPI_MODEL_UNICYCLE_EULER) {           // euler, control/src/mppi:57-58
              dx = fmaf(gfac, cth, dx);
              dy = fmaf(gfac, sth, dy);
              th = th + kth;
              const R cn2 = cth * ca - sth * sa;
              sth = fmaf(sth, ca, cth * sa);
              cth = cn2;
            } else {                                            // rk4, control/src/mppi:39-54 (see common.cuh)
              const R c2 = fmaf(cth, ca, -(sth * sa)), s2 = fmaf(sth, ca, cth * sa);
              const R c4 = fmaf(c2, ca, -(s2 * sa)), s4 = fmaf(s2, ca, c2 * sa);
              dx = fmaf(gfac, fmaf(4.0f, c2, cth + c4), dx);
              dy = fmaf(gfac, fmaf(4.0f, s2, sth + s4), dy);
              th = Math<R>::wrap_once_(th + kth);
              cth = c4;
              sth = s4;
            }
            R c = running_cost<R>(cc, dx, dy, th, R(0), R(0), R(0), R(0)) + cn;   // control/src/mppi:180-184
            if (HAS_GRID) c += grid_cost<R>(cc, cells, dx, dy);
            acc += c;
            P[t * PS + lane] = acc;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[p]);
        ++n_cons[p];
        if (g & 1) Math<R>::sincos_(th, sth, cth);              // re-synchronise (cos, sin) every 4 steps
      }
      acc += terminal_cost<R>(cc, dx, dy, th);                  // control/src/mppi:165-171
      if (!valid) acc = Math<R>::inf();
      P[(T - 1) * PS + lane] = acc;                             // row T-1 holds the rollout total
      __syncwarp();
      // transposed pass: lane l owns row t = 32 i + l (the producers already work on the next tile)
      for (int tb = 0; tb < T; tb += 32) {
        const int t = tb + lane;
        if (t < T)
          transposed_row<R, MODE, kWsTile>(a, t, tile, cta, nCTA, P, run, ccount, nullptr, cost_to_go, neg_inv_lam, margin, std0,
                                           std1, step);
      }
      __syncwarp();
    }
  }
  __syncthreads();   // all floor sums flushed, all rows final

  // ---- epilogue: one partial per (t, CTA) ------------------------------------------------------
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int t = tid; t < T; t += kWsThreads) {
    const size_t idx = (size_t)t * nCTA + cta;
    if (MODE == MODE_SOFTMIN)
      reinterpret_cast<Vec4*>(a.part)[idx] = run[t];
    else
      a.cand_meta[idx] = make_float4(run[t].x, run[t].y, __int_as_float(ccount[t]), 0.f);
    a.epart[2 * idx] = (double)(long long)ez64[2 * t];
    a.epart[2 * idx + 1] = (double)(long long)ez64[2 * t + 1];
  }
}

}  // namespace mppi
