// reduce_kernels_args.h -- kernel argument blocks (plain structs, shared by host and device TUs)
#pragma once
#include "common.cuh"
namespace mppi {

// engine-lifetime fp32 constants of the LEAN rollout kernel, folded on the host (no fp64 math in its prologue)
struct LeanStatic {
  // controls in clip units s = u / (2 u_max) + 1/2 in [0, 1]; positions in cost units d' = sq * d, sq = sqrt(Q/2)
  float A0, A1, Ac;          // half yaw increment a = A0 s0 + A1 s1 + Ac (Euler: the full increment)
  float G0, G1, Gc;          // Simpson factor sq * dt * speed / 6 (Euler: sq * dt * speed) = G0 s0 + G1 s1 + Gc; bicycle: G0 * v
  float um0, um1, bk;        // bicycle: v = 2 um0 s0 - um0, delta = 2 um1 s1 - um1, a = bk * v * tan(delta)
  float inv2um0, inv2um1;    // 1 / (2 u_max)
  float sq;                  // sqrt(Q[0] / 2) (== sqrt(Q[1] / 2): admission condition)
  float p1x, p1y, p1th;      // P1 / (Q/2) for x, y (terminal cost on cost-unit positions); P1[2]
  float g_inv_res, w_obs_100;   // cells per cost unit
  int split;                 // SM-wide balanced kernel: first step rolled by the second pair of warps of the shared tile
  uint32_t pkx[MPPI_PHILOX_ROUNDS], pky[MPPI_PHILOX_ROUNDS];   // Philox key schedule key + i * Weyl, per round
};

struct RolloutArgs {
  StaticParams sp;
  const DynState* dyn;
  const void* nom;             // nominal block, Real[4][T]: U0, U1, g0, g1 (g = lam * u.sig, per t)
  const signed char* grid;     // int8 cells (padded to 16 B)
  const double* eps_ext;       // (T,2,K) f64 when sp.noise_external
  void* part;                  // SOFTMIN: Vec4[T][nCTA] (m,S,N0,N1)
  double* epart;               // [T][nCTA][2] floor sums (fixed-point integer as double, or real if external)
  float4* cand_meta;           // SCREEN: [T][nCTA] (running min, applied limit, count as int bits, -)
  uint2* cand;                 // SCREEN: [T][nCTA][kMaxCand]  (k_local, float bits of V)
  void* vcap;                  // capture: Real[T][K]
  int ntiles;
  StepInput in;                // x0 / goal of this step
  LeanStatic lean;             // LEAN variant only
  unsigned long long* debug_ts;   // -DMPPI_EXP_TIMELINE builds only: [nCTA][8] globaltimer stamps + SM id (profiling aid)
};

struct FinalizeArgs {
  StaticParams sp;
  DynState* dyn;
  const double* gather;      // [world][T][6]
  double* Umaster;           // [2][T] latest_uvec
  double* Ulast;             // [2][T] update_action result before the shift
  float* nomF;               // [4][T]
  double* nomD;              // [4][T]
  double sg_a, sg_b;         // Gram basis on z=-h..h: p2 = z^2 - a, p3 = z^3 - b z
  double sg_inv_norm[4];     // 1 / sum_j p_i(z_j)^2
  int mode;                  // 0: full step, 1: update only (mppi_update_action)
  int closed_loop;           // 1: dyn->x0 <- x_next (device-resident loop of mppi_bench)
  // fused step (any world size): the finalizer block reads the MERGED row of every t -- the clipped update of all ranks,
  // formed by the row blocks themselves -- from this rank's flag-in-data buffer uint2 [2 parity][T][kRow2Words] (see
  // reduce_kernels.cuh); nullptr: the records are taken from `gather` and merged here (split-phase step with an external
  // exchange, mppi_update_action)
  const uint2* ll2_local;
  unsigned long long* debug_ts;   // optional 8 globaltimer stamps of the finalize phase (profiling aid): [1] all rows seen,
                                  // [3] filter coefficients done, [4] result published
  StepInput in;              // x0 / goal of this step (also used by the reduce kernel's fp64 re-evaluation)
  HostWire* host_res;        // mapped pinned host memory (nullptr: results are fetched from DynState)
  unsigned long long seq;    // sequence number of this step: its low 32 bits validate every word of host_res
};

struct RendezvousArgs {      // mppi_bench, world > 1 (reduce.cu: rendezvous_kernel)
  uint2* peers[kMaxFusedWorld];
  size_t rows_uint2;         // size of the row area of a buffer, in uint2: the rendezvous flags live behind it
  unsigned int epoch;        // > 0, +1 per rendezvous
  int world, rank;
};

struct ReduceArgs {
  StaticParams sp;
  FinalizeArgs fin;          // fin.dyn is THE DynState; the rest is used when fuse_finalize != 0
  int fused;                 // 1: rows leave through the flag-in-data buffers and block T of the grid (the finalizer) runs the
                             //    finalize phase as soon as the rows of all ranks have arrived; 0: rows go to `record` only
  int rank;
  uint2* ll2_local;          // this rank's buffer of merged rows (written by its row blocks, read by its finalizer block)
  uint2* ll_peers[kMaxFusedWorld];   // row-buffer base pointers of all ranks, IN the argument block (no dependent load on the
                             // serial tail): this rank's own buffer and, for world > 1, the CUDA IPC mappings of the peers'
                             // buffers (stores travel over NVLink)
  unsigned long long* debug_ts;   // optional [T][8] globaltimer stamps of the reduce phases (profiling aid)
  const void* part;          // SOFTMIN partials Vec4[T][nCTA]
  const double* epart;       // [T][nCTA][2]
  const float4* cand_meta;   // SCREEN
  const uint2* cand;
  const double* nomD;        // f64 nominal block [4][T]
  const signed char* grid;
  const double* eps_ext;
  double* record;            // [T][6]
  int nCTA;
};


}  // namespace mppi
