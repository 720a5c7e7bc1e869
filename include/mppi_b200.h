/*
 * mppi_b200.h -- C ABI of the B200-native MPPI rollout engine (libmppi_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of moribots/motion_planning: the MPPI
 * controller step implemented by the reference as the Python/NumPy class `MPPI` in
 * control/src/mppi:61-213 (there is no compiled FFI in the reference -- SURVEY.md section 0/8b --
 * so every entry point below names the reference *method* it replaces).
 *
 * Conventions
 *   - plain pointers and sizes only; all host arrays are C-order float64 like the reference's
 *     ndarrays (control/src/mppi uses float64 throughout); the caller owns every buffer passed
 *     in or out, the handle owns all device memory.
 *   - every function returns an mppi_status (0 = OK), mirroring the reference C++ library's
 *     0-success / non-zero-failure returns with out-pointers for results
 *     (control/include/control/TrajMPC.hpp:72,110-122,134,156); mppi_last_error() gives text.
 *   - a handle is bound to one CUDA device; calls may come from any host thread (rospy delivers
 *     Controller.pos_cb on a subscriber thread, control/src/mppi:299-303) but not concurrently.
 *   - there is NO CPU fallback: without a CUDA device mppi_create fails with MPPI_ERR_NO_DEVICE.
 */
#ifndef MPPI_B200_H_
#define MPPI_B200_H_

#ifdef __CUDACC_RTC__   /* run-time compilation of a user model (mppi_create_user): no host headers */
typedef signed char int8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define MPPI_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define MPPI_API __attribute__((visibility("default")))
#else
#define MPPI_API
#endif

typedef struct mppi_engine* mppi_handle;

typedef enum {
  MPPI_OK = 0,
  MPPI_ERR_INVALID = 1,      /* bad argument / bad params */
  MPPI_ERR_CUDA = 2,         /* a CUDA runtime call failed */
  MPPI_ERR_NO_DEVICE = 3,    /* no usable CUDA device (no CPU fallback exists) */
  MPPI_ERR_UNSUPPORTED = 4,  /* valid request this build cannot serve */
  MPPI_ERR_STATE = 5,        /* call sequence error (e.g. step_finish without step_local) */
  MPPI_ERR_NONFINITE = 6,    /* NaN/Inf reached the control output */
  MPPI_ERR_RETRY = 7         /* split-phase sharded step only: the fp32 screen of precision MIXED overflowed on some rank;
                                nothing was applied on ANY rank -- call mppi_step_local again with the same x0 (it now runs
                                the fp64 pipeline), exchange, mppi_step_finish.  Every rank gets this status for the same step. */
} mppi_status;

/* model = integrator-step functor of the reference ctor `MPPI(model=rk4, ...)`, control/src/mppi:62,66 */
typedef enum {
  MPPI_MODEL_DIFF_DRIVE = 0,     /* rk4 + dd_dynamics        control/src/mppi:23-30,39-54 */
  MPPI_MODEL_UNICYCLE_EULER = 1, /* euler + unicycle_dynamics control/src/mppi:33-36,57-58 */
  MPPI_MODEL_BICYCLE = 2,        /* NEW (BASELINE.json config 3): u=(v,delta), RK4 + wrap */
  MPPI_MODEL_USER = 3            /* caller-supplied ODE (and optionally cost), compiled at run time: mppi_create_user */
} mppi_model;

typedef enum {
  MPPI_WEIGHT_COST_TO_GO = 0,    /* reference: per-t softmin on cost-to-go, control/src/mppi:175,187-196 */
  MPPI_WEIGHT_TOTAL_COST = 1     /* north_star wording: one softmin on the K total rollout costs */
} mppi_weighting;

typedef enum {
  MPPI_PRECISION_F32 = 0,    /* fp32 rollouts + fp32 online softmin (fastest, conditioning-limited) */
  MPPI_PRECISION_F64 = 1,    /* everything in float64 like the reference (validation / strict parity) */
  MPPI_PRECISION_MIXED = 2   /* fp32 rollouts screen the softmin support, fp64 re-evaluates it */
} mppi_precision;

typedef struct {
  uint32_t struct_size;      /* = sizeof(mppi_params); ABI guard */
  uint32_t abi_version;      /* = MPPI_B200_ABI_VERSION */
  int32_t K;                 /* samples rolled on THIS device     `samples`, control/src/mppi:62,64 */
  int32_t T;                 /* horizon, even, >= 6               `horizon`, control/src/mppi:62,63 */
  int32_t model;             /* mppi_model */
  int32_t weighting;         /* mppi_weighting */
  int32_t precision;         /* mppi_precision */
  int32_t device;            /* CUDA device ordinal */
  double dt;                 /* <=0 -> 1/T                        control/src/mppi:67 */
  double q[3];               /* diag(Q)                           control/src/mppi:69 */
  double r[4];               /* R, 2x2 row-major                  control/src/mppi:71 */
  double p1[3];              /* diag(P1)                          control/src/mppi:73 */
  double sig[4];             /* sig 2x2 row-major (cost term lam*u.sig.eps)  control/src/mppi:88,184 */
  double noise_std[2];       /* std-dev of eps rows; reference = sig[0,0] for both  control/src/mppi:144-146 */
  double lambda;             /* lam                               control/src/mppi:89 */
  double u_max[2];           /* control clip                      control/src/mppi:18,151-152,198-206 */
  double wheel_radius;       /* control/src/mppi:19 */
  double wheel_base;         /* control/src/mppi:20 (bicycle: axle distance) */
  double eps_floor;          /* +1e-8 weight floor                control/src/mppi:193 */
  uint64_t seed;             /* Philox4x32-10 key */
  /* K-sharding over GPUs: this device rolls global ids [k_offset, k_offset+K) */
  int64_t k_offset;
  int64_t k_total;           /* <=0 -> K */
  int32_t world_size;        /* <=0 -> 1 */
  int32_t rank;
  void* stream;              /* optional cudaStream_t to launch on (NULL -> engine-owned stream + CUDA graph) */
  double refine_margin;      /* MIXED: cost window re-evaluated in fp64; <=0 -> default = 40 lambda + a head-room for the
                                fp32 error of the screen that scales with the step's cost magnitude */
} mppi_params;

#ifndef __CUDACC_RTC__   /* (device-side run-time compilation only needs the enums and structs above) */
/* Fill *p with the reference's hard-coded constants (control/src/mppi:18-20,62-73,88-89): K=10, T=100. */
MPPI_API mppi_status mppi_default_params(mppi_params* p);

/* MPPI.__init__ (control/src/mppi:62-77): allocate device state, U := zeros(2,T). */
MPPI_API mppi_status mppi_create(const mppi_params* p, mppi_handle* out);
MPPI_API mppi_status mppi_destroy(mppi_handle h);

/* ---- user-defined dynamics / cost functors -----------------------------------------
 * The reference takes its model as a constructor argument, `MPPI(model=rk4)` (control/src/mppi:62,66), called as
 * model(states, u, dt) (:154,213); its C++ library registers an ODE functor `ode(x, u, xdot_out)` with a generic RK4
 * (control/include/control/rk4.hpp:32,58).  Here the functor is CUDA source text, compiled for sm_100a at run time (NVRTC)
 * into its own instantiation of the rollout / reduce / finalize kernels:
 *
 *   kind 0     template <typename R> __device__ void mppi_user_ode(const R x[3], const R u[2], R xdot[3]);
 *              an arbitrary ODE of (x, y, theta), integrated by the generic integrator below; precision F64 or F32
 *   kind 1     template <typename R> __device__ void mppi_user_speed_yaw(const R u[2], R* speed, R* yaw_rate);
 *              a KINEMATIC functor: xdot = speed(u) cos(theta), ydot = speed(u) sin(theta), thetadot = yaw_rate(u) -- the family
 *              every built-in model belongs to (dd_dynamics, unicycle_dynamics: control/src/mppi:23-36).  It is dropped into the
 *              built-in kernels, so all three precisions work (the fp32 screen of MIXED needs speed_max / yaw_rate_max, bounds
 *              of |speed| and |yaw_rate| over the clipped controls) and a functor that restates a built-in model reproduces it
 *   optional   (has_cost != 0; precision F64 or F32)
 *              template <typename R> __device__ R mppi_user_running_cost(const R x[3], const R goal[3], const R u_nom[2],
 *                                                                       const R eps[2], int t);
 *              template <typename R> __device__ R mppi_user_terminal_cost(const R x[3], const R goal[3]);
 *
 * x = (x, y, theta) AFTER the step for the running cost (control/src/mppi:160-161), u_nom = the nominal control of step t,
 * eps = the sample's noise.  R is float or double (both are instantiated).  Without a cost functor the reference's quadratic
 * cost (Q, R, P1 of mppi_params, control/src/mppi:165-171,180-184) applies.  integrator: 0 = classic RK4 with the control
 * held over the step (control/src/mppi:39-50), 1 = explicit Euler (:57-58); wrap_theta != 0 wraps theta into (-pi, pi] after
 * every step (:52-53); a kind-1 functor is integrated like the built-in models: integrator 0 = rk4 WITH the wrap, 1 = Euler
 * WITHOUT it (wrap_theta must say the same).  params->model is ignored (set to MPPI_MODEL_USER).  Compilation errors:
 * MPPI_ERR_INVALID, text in mppi_last_error().  Zero-initialise the struct (`mppi_user_model um = {0};`): every field's zero is
 * its default. */
typedef struct {
  const char* source;        /* CUDA C++ text defining the functions above */
  int32_t integrator;        /* 0 RK4, 1 explicit Euler */
  int32_t wrap_theta;
  int32_t has_cost;
  int32_t kind;              /* 0 ODE functor, 1 kinematic functor */
  double speed_max;          /* kind 1 with precision MIXED: max |speed| over the clipped controls */
  double yaw_rate_max;       /* kind 1 with precision MIXED: max |yaw_rate| over the clipped controls */
} mppi_user_model;
MPPI_API mppi_status mppi_create_user(const mppi_params* p, const mppi_user_model* um, mppi_handle* out);
/* Compile a functor text without creating an engine (needs no GPU): MPPI_OK, or MPPI_ERR_INVALID with the compiler's log in
 * mppi_last_error(). */
MPPI_API mppi_status mppi_check_user_model(const mppi_user_model* um);

/* MPPI.initialize (control/src/mppi:79-83): latest_uvec := zeros(2,T). Noise stream is NOT rewound. */
MPPI_API mppi_status mppi_reset(mppi_handle h);

/* `goal` argument of get_path (control/src/mppi:85-87, set by Controller at :337,:352,:372). */
MPPI_API mppi_status mppi_set_goal(mppi_handle h, const double goal[3]);

/* `sig`, `lam` arguments of get_path (control/src/mppi:88-89). noise std := sig[0] as in the reference. */
MPPI_API mppi_status mppi_set_sampling(mppi_handle h, const double sig[4], double lambda);
/* NEW: per-channel noise std-dev and clip (bicycle has heterogeneous channels). */
MPPI_API mppi_status mppi_set_noise_std(mppi_handle h, const double noise_std[2]);

/* MPPI.get_path (control/src/mppi:85-102) = step(): K rollouts x T steps, per-t softmin update,
 * clip/SavGol/clip, perform_action, receding-horizon shift.  u_out = uvec[-1] = U[:,0] before the
 * shift (control/src/mppi:96-97,379); x_next = the returned predicted state.  Either may be NULL. */
MPPI_API mppi_status mppi_step(mppi_handle h, const double x0[3], double u_out[2], double x_next[3]);

/* latest_uvec accessors (control/src/mppi:81,92,100-101): U is (2,T) row-major. */
MPPI_API mppi_status mppi_get_nominal(mppi_handle h, double* U);
MPPI_API mppi_status mppi_set_nominal(mppi_handle h, const double* U);
/* the (2,T) sequence of the last step BEFORE the shift (what update_action returned, :92). */
MPPI_API mppi_status mppi_get_last_update(mppi_handle h, double* U);

/* NEW (BASELINE.json config 4): int8 occupancy grid, row-major idx = ix + iy*W,
 * values 0/50/100 (map/src/map/grid.cpp:126-144,251-266), origin = map_min (map/src/viz_grid.cpp:112-129);
 * running cost += w_obs * value/100, outside the map counts as 100. */
MPPI_API mppi_status mppi_set_grid(mppi_handle h, const int8_t* cells, int32_t W, int32_t H, double res,
                          double x_min, double y_min, double w_obs);
/* Patch a rectangle of the resident grid without reallocation or reconfiguration: rows y0..y0+h-1, columns x0..x0+w-1,
 * `cells` row-major (h, w).  This is how the incrementally revealed map of the reference's simulated sensor
 * (Grid::update_grid / fake_occupancy_grid, map/src/map/grid.cpp:155-199: the (2v+1)^2 neighbourhood of the robot's cell per
 * tick, global_planner/src/dsl.cpp:325-327) reaches the controller.  Asynchronous: staged in pinned memory and copied on the
 * engine's stream, i.e. ordered before the next step; the call does not wait for the device. */
MPPI_API mppi_status mppi_update_grid(mppi_handle h, const int8_t* cells, int32_t x0, int32_t y0, int32_t w, int32_t hgt);
MPPI_API mppi_status mppi_clear_grid(mppi_handle h);

/* ---- noise record / replay ("identical RNG seeds" protocol) ----------------------- */
/* Replay: use eps (T,2,K) float64 -- exactly what np.random.normal returned per t
 * (control/src/mppi:143-146) -- for every following step until mppi_use_philox. */
MPPI_API mppi_status mppi_set_noise(mppi_handle h, const double* eps);
/* Back to in-register Philox4x32-10; rewinds the step counter to 0. */
MPPI_API mppi_status mppi_use_philox(mppi_handle h, uint64_t seed);
/* Record: the eps (T,2,K) the LAST step used (regenerated from the same counters). */
MPPI_API mppi_status mppi_get_noise(mppi_handle h, double* eps);
/* Capture the cost-to-go of following steps (debug; materialises T*K values in HBM). */
MPPI_API mppi_status mppi_set_capture(mppi_handle h, int32_t on);
/* value_fcn (T,K) of the last step as get_cost2go returns it (control/src/mppi:175-178). */
MPPI_API mppi_status mppi_get_cost_to_go(mppi_handle h, double* V);

/* ---- the reference's finer-grained methods, as standalone device ops ------------------------- */
/* MPPI.get_cost2go(state, uvec, goal, lam, sig) with explicit noise (control/src/mppi:127-178). */
MPPI_API mppi_status mppi_cost_to_go(mppi_handle h, const double x0[3], const double* U, const double goal[3],
                            const double* eps, double* V);
/* MPPI.update_action(uvec, eps, value_fcn, sig, lam) (control/src/mppi:186-208). */
MPPI_API mppi_status mppi_update_action(mppi_handle h, const double* U_in, const double* eps, const double* V,
                               double* U_out);
/* MPPI.perform_action(state, uvec) (control/src/mppi:210-213): one model step with U[:,0]. */
MPPI_API mppi_status mppi_perform_action(mppi_handle h, const double x0[3], const double* U, double x_out[3]);
/* the `model` functor itself on n states: x (3,n), u (2,n) -> x_out (3,n) (control/src/mppi:39-54,154). */
MPPI_API mppi_status mppi_model_step(mppi_handle h, const double* x, const double* u, int32_t n, double* x_out);

/* ---- multi-GPU split-phase step (K sharded over ranks, one tiny exchange) --------- */
/* phase 1: rollouts + local merge; leaves this rank's record (T x 6 float64:
 * min V, sum e, sum e*eps0, sum e*eps1, sum eps0, sum eps1) in device memory. */
MPPI_API mppi_status mppi_step_local(mppi_handle h, const double x0[3]);
/* device pointers for the exchange: local record (T*6 f64) and gather buffer (world*T*6 f64). */
MPPI_API mppi_status mppi_exchange_buffers(mppi_handle h, void** record, size_t* record_bytes,
                                  void** gather, size_t* gather_bytes);
/* host-staged exchange for transports that cannot take device pointers (e.g. gloo): read this rank's
 * record (T*6 f64) after mppi_step_local / write all world_size records before mppi_step_finish. */
MPPI_API mppi_status mppi_read_record(mppi_handle h, double* record);
MPPI_API mppi_status mppi_write_gather(mppi_handle h, const double* all_records);
/* phase 2 (after the caller all-gathered records into the gather buffer on the same stream):
 * merge world_size records, update, filter, shift -- every rank ends with identical U. */
MPPI_API mppi_status mppi_step_finish(mppi_handle h, double u_out[2], double x_next[3]);

/* ---- fused peer-to-peer exchange over NVLink (ranks of one node) --------------------------------------
 * Instead of a collective call between mppi_step_local and mppi_step_finish, the ranks map each other's
 * row buffers (CUDA IPC): every block of rank r's reduce kernel stores the row of its time step straight into every peer's
 * buffer as flag-in-data words (8-byte stores carrying 4 bytes of payload and the step's flag: no fence, no arrival flag),
 * merges the peers' rows of the same time step, and the finalizer block of the same kernel finishes the update.  After
 * connecting, mppi_step / mppi_bench work for world_size > 1 and the whole sharded step is TWO kernel launches per rank.
 *   1. every rank: mppi_p2p_export(h, handle)      -- 64-byte cudaIpcMemHandle_t of its buffer
 *   2. all-gather the handles by any means (rank-major, world_size x 64 bytes)
 *   3. every rank: mppi_p2p_connect(h, all_handles)                                                  */
MPPI_API mppi_status mppi_p2p_export(mppi_handle h, void* handle64);
MPPI_API mppi_status mppi_p2p_connect(mppi_handle h, const void* all_handles);

/* ---- measurement ----------------------------------------------------------------------------- */
typedef struct {
  float step_ms;        /* mean device time of one whole step (all kernels), CUDA events on the launch stream */
  float rollout_ms;     /* mean device time of the fused rollout+cost kernel alone */
  float reduce_ms;      /* mean device time of the softmin/refine reduction kernel INCLUDING the finalize phase (its finalizer block) */
  float finalize_ms;    /* ~0: the update/filter/shift phase is fused into the reduce kernel (kept for ABI stability) */
  int32_t launches;     /* kernels launched inside the timed region */
  int32_t steps;
  int32_t refine_candidates;  /* MIXED: fp64 re-evaluations in the last step */
  int32_t refine_overflow;    /* MIXED: candidate-list overflows seen (forces a full-fp64 redo) */
  double refine_max_dev;      /* MIXED: max |V_fp32 - V_fp64| over re-evaluated rollouts, last step */
  double refine_head_room;    /* MIXED: head-room of the screening window of the last step; a step whose refine_max_dev exceeds
                                 half of it is redone in fp64 (counted in refine_overflow) */
} mppi_timing;

/* Device-resident closed loop on the model (x0 <- x_next on the device, as solve_path does,
 * control/src/mppi:117-119): `warmup` untimed + `steps` timed steps; inputs never leave HBM.  The two kernels of a step are
 * launched back to back by a host that runs ahead of the device (MPPI_B200_BENCH=graph replays a captured CUDA graph instead).
 * flush_l2 != 0 writes a > L2-sized buffer between timed steps (outside the timed intervals).
 * per_kernel != 0 additionally brackets every kernel with events (eager launches). */
MPPI_API mppi_status mppi_bench(mppi_handle h, const double x0[3], int32_t steps, int32_t warmup,
                       int32_t flush_l2, int32_t per_kernel, mppi_timing* out);
/* statistics of the most recent step (refine_* fields; *_ms only valid after mppi_bench). */
MPPI_API mppi_status mppi_last_stats(mppi_handle h, mppi_timing* out);
/* peak fp32 FMA rate of this device measured with a register-resident FFMA chain (TFLOP/s). */
MPPI_API mppi_status mppi_measure_fp32_peak(int32_t device, double* tflops, double* sm_clock_mhz);

/* bytes mppi_step moves per call: host->device (x0, goal) and device->host (result block). */
MPPI_API mppi_status mppi_io_bytes(mppi_handle h, size_t* h2d, size_t* d2h);
/* measurement aid: overwrite 256 MiB (> L2), then read 256 MiB of clean lines, and synchronise: the next step starts
 * cache-cold without inheriting the write-back of the flush's own dirty lines. */
MPPI_API mppi_status mppi_debug_flush_l2(mppi_handle h);
/* launch configuration of the rollout kernel (diagnostics): block, grid, tiles, dynamic smem, CTAs/SM, regs,
 * code path (0 general, 1 fast, 2 lean), reserved */
MPPI_API mppi_status mppi_launch_info(mppi_handle h, int32_t info[8]);

/* profiling aid: globaltimer stamps (ns) of the reduce-kernel phases of the last step, [T][8];
 * the first call arms the stamps. */
MPPI_API mppi_status mppi_debug_reduce_timestamps(mppi_handle h, unsigned long long* out);
/* Profiling aid, meaningful only in libraries built with -DMPPI_EXP_TIMELINE: %globaltimer stamps (ns) of the
 * rollout kernel's CTAs in the last step, out[n_ctas][8] = entry, loads issued, prologue done, loop done, total
 * stored, barrier passed, tile done, SM id.  The first call (out may be NULL) arms the stamps. */
MPPI_API mppi_status mppi_debug_rollout_timestamps(mppi_handle h, unsigned long long* out, size_t n_ctas);

/* Profiling aid: mean host-side duration (us) of the phases of mppi_step since the previous call of this function:
 * out[0] entry -> rollout kernel launched, [1] -> reduce kernel launched, [2] -> result seen in mapped host memory,
 * [3] -> return. */
MPPI_API mppi_status mppi_debug_host_timing(mppi_handle h, double out[4]);

MPPI_API const char* mppi_last_error(void);
MPPI_API const char* mppi_version(void);
MPPI_API int32_t mppi_device_count(void);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* MPPI_B200_H_ */
