/* TEST INFRASTRUCTURE ONLY -- C (OpenMP) restatement of the reference MPPI step.
 *
 * Follows /root/reference/control/src/mppi (Python/NumPy, float64), operation by operation:
 *   noise        np.random.seed / np.random.normal  control/src/mppi:15,143-146  (legacy MT19937 +
 *                polar Box-Muller with the cached second value; bit-exact with NumPy, checked in
 *                tests/test_oracle_port.py against tests/golden/ref_rng_kat.npz)
 *   dd_dynamics  control/src/mppi:23-30      rk4 + theta wrap  control/src/mppi:39-54
 *   get_cost2go  control/src/mppi:127-178    get_cost          control/src/mppi:180-184
 *   update_action control/src/mppi:186-208   savgol_filter(U, T-1, 3) (SciPy, third party) :202
 *   perform_action :210-213                  shift :100-101
 * Like the reference it materialises eps (T,2,K) and V (T,K); unlike the reference the K loop is
 * compiled and threaded (OpenMP over rollouts), which makes it the "strong CPU" baseline that
 * bench.py reports next to the GPU number (cpu_baseline.kind = "port").
 *
 * Used only by tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke(); never by the product.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ legacy NumPy RandomState */
typedef struct {
  uint32_t mt[624];
  int pos;
  int has_gauss;
  double gauss;
} rk_state;

void port_seed(rk_state* s, uint32_t seed) { /* init_genrand, as np.random.seed(int) */
  for (int i = 0; i < 624; ++i) {
    s->mt[i] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)(i + 1);
  }
  s->pos = 624;
  s->has_gauss = 0;
  s->gauss = 0.0;
}

static uint32_t rk_u32(rk_state* s) {
  if (s->pos == 624) {
    uint32_t* mt = s->mt;
    int i;
    for (i = 0; i < 624 - 397; ++i) {
      uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
      mt[i] = mt[i + 397] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    }
    for (; i < 623; ++i) {
      uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
      mt[i] = mt[i + (397 - 624)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    }
    uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ (-(int32_t)(y & 1) & 0x9908b0dfu);
    s->pos = 0;
  }
  uint32_t y = s->mt[s->pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

static double rk_double(rk_state* s) { /* 53-bit: (a >> 5, b >> 6) */
  uint32_t a = rk_u32(s) >> 5, b = rk_u32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

double port_gauss(rk_state* s) { /* legacy_gauss: polar method, returns f*x2 first and caches f*x1 */
  if (s->has_gauss) {
    s->has_gauss = 0;
    double g = s->gauss;
    s->gauss = 0.0;
    return g;
  }
  double x1, x2, r2;
  do {
    x1 = 2.0 * rk_double(s) - 1.0;
    x2 = 2.0 * rk_double(s) - 1.0;
    r2 = x1 * x1 + x2 * x2;
  } while (r2 >= 1.0 || r2 == 0.0);
  double f = sqrt(-2.0 * log(r2) / r2);
  s->gauss = f * x1;
  s->has_gauss = 1;
  return f * x2;
}

/* np.random.normal(0, scale, size=n): loc + scale * gauss */
void port_normal(rk_state* s, double scale, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = 0.0 + scale * port_gauss(s);
}

/* ------------------------------------------------------------------ parallel noise for timing runs
 * splitmix64 counter stream + Box-Muller per (k, t): statistically N(0, scale^2); used only when the
 * baseline is TIMED, so that the serial MT19937 does not cap the multi-core number. */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

typedef struct {
  int K, T;
  double dt;
  double q[3], R[4], p1[3], sig[4];
  double lam, u_max[2], r, L, eps_floor, noise_std;
} port_params;

void port_default_params(port_params* p, int K, int T) {
  memset(p, 0, sizeof(*p));
  p->K = K;
  p->T = T;
  p->dt = 1.0 / (double)T;                 /* control/src/mppi:67 */
  p->q[0] = p->q[1] = 1e3;                 /* :69 */
  p->R[0] = p->R[3] = 1.0;                 /* :71 */
  p->p1[0] = p->p1[1] = p->p1[2] = 1e3;    /* :73 */
  p->sig[0] = p->sig[3] = 0.9;             /* :88 */
  p->lam = 1e-3;                           /* :89 */
  p->u_max[0] = p->u_max[1] = 6.35492;     /* :18 */
  p->r = 0.033;                            /* :19 */
  p->L = 0.16;                             /* :20 */
  p->eps_floor = 1e-8;                     /* :193 */
  p->noise_std = 0.9;                      /* sig[0,0], :145 */
}

static inline void dd(const port_params* p, const double x[3], double u0, double u1, double o[3]) {
  o[0] = (p->r / 2.0) * cos(x[2]) * (u0 + u1); /* control/src/mppi:27 */
  o[1] = (p->r / 2.0) * sin(x[2]) * (u0 + u1); /* :28 */
  o[2] = (p->r / p->L) * (u1 - u0);            /* :29 */
}

static inline void rk4(const port_params* p, double x[3], double u0, double u1) {
  const double dt = p->dt;
  double k1[3], k2[3], k3[3], k4[3], xt[3];
  dd(p, x, u0, u1, k1);
  for (int i = 0; i < 3; ++i) { k1[i] *= dt; xt[i] = x[i] + k1[i] / 2; }   /* :45-46 */
  dd(p, xt, u0, u1, k2);
  for (int i = 0; i < 3; ++i) { k2[i] *= dt; xt[i] = x[i] + k2[i] / 2; }   /* :46-47 */
  dd(p, xt, u0, u1, k3);
  for (int i = 0; i < 3; ++i) { k3[i] *= dt; xt[i] = x[i] + k3[i]; }       /* :47-48 */
  dd(p, xt, u0, u1, k4);
  for (int i = 0; i < 3; ++i) { k4[i] *= dt; x[i] = x[i] + (1.0 / 6.0) * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]); } /* :50 */
  x[2] = x[2] - (ceil((x[2] + M_PI) / (2.0 * M_PI)) - 1.0) * 2.0 * M_PI;   /* :52-53 */
}

static inline double clipd(double v, double lim) { return v < -lim ? -lim : (v > lim ? lim : v); }

/* get_cost2go, control/src/mppi:127-178.  eps (T,2,K) in; V (T,K) out. */
void port_cost2go(const port_params* p, const double x0[3], const double* U, const double goal[3], const double* eps,
                  double* V) {
  const int K = p->K, T = p->T;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k) {
    double x[3] = {x0[0], x0[1], x0[2]};
    for (int t = 0; t < T; ++t) {
      const double e0 = eps[((size_t)t * 2 + 0) * K + k], e1 = eps[((size_t)t * 2 + 1) * K + k];
      const double n0 = U[t], n1 = U[T + t];
      const double u0 = clipd(n0 + e0, p->u_max[0]), u1 = clipd(n1 + e1, p->u_max[1]);   /* :147-152 */
      rk4(p, x, u0, u1);                                                                 /* :154 */
      const double d0 = x[0] - goal[0], d1 = x[1] - goal[1], d2 = x[2] - goal[2];
      /* get_cost :180-184 -- u is the NOMINAL control */
      const double quad = d0 * p->q[0] * d0 + d1 * p->q[1] * d1 + d2 * p->q[2] * d2;
      const double uRu = n0 * (p->R[0] * n0 + p->R[2] * n1) + n1 * (p->R[1] * n0 + p->R[3] * n1);
      const double use = (n0 * p->sig[0] + n1 * p->sig[2]) * e0 + (n0 * p->sig[1] + n1 * p->sig[3]) * e1;
      V[(size_t)t * K + k] = 0.5 * (quad + uRu) + p->lam * use;
    }
    const double d0 = x[0] - goal[0], d1 = x[1] - goal[1], d2 = x[2] - goal[2];          /* :165-171 */
    V[(size_t)(T - 1) * K + k] += d0 * p->p1[0] * d0 + d1 * p->p1[1] * d1 + d2 * p->p1[2] * d2;
    double acc = 0.0;                                                                     /* :175 */
    for (int t = T - 1; t >= 0; --t) {
      acc += V[(size_t)t * K + k];
      V[(size_t)t * K + k] = acc;
    }
  }
}

/* savgol_filter(u, T-1, 3, mode='interp') on one row of length T (SciPy; call site :202) */
static void savgol_row(int T, const double* u, double* out) {
  const int W = T - 1, h = W / 2;
  long double s2 = 0, s4 = 0;
  for (int j = 0; j < W; ++j) { long double z = j - h; s2 += z * z; s4 += z * z * z * z; }
  const long double a = s2 / W, b = s4 / s2;
  for (int fit = 0; fit < 2; ++fit) {
    long double c[4] = {0, 0, 0, 0}, n[4] = {0, 0, 0, 0};
    for (int j = 0; j < W; ++j) {
      long double z = j - h, pz[4] = {1.0L, z, z * z - a, z * z * z - b * z};
      for (int i = 0; i < 4; ++i) { c[i] += pz[i] * u[j + fit]; n[i] += pz[i] * pz[i]; }
    }
    const int lo = fit ? h + 1 : 0, hi = fit ? T - 1 : h;
    for (int t = lo; t <= hi; ++t) {
      long double z = t - fit - h;
      out[t] = (double)(c[0] / n[0] + c[1] / n[1] * z + c[2] / n[2] * (z * z - a) + c[3] / n[3] * (z * z * z - b * z));
    }
  }
}

/* update_action, control/src/mppi:186-208.  V is modified in place like the reference (:189). */
void port_update_action(const port_params* p, double* U, const double* eps, double* V) {
  const int K = p->K, T = p->T;
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; ++t) {
    double* v = V + (size_t)t * K;
    const double* e0 = eps + ((size_t)t * 2) * K;
    const double* e1 = e0 + K;
    double m = v[0];
    for (int k = 1; k < K; ++k) m = v[k] < m ? v[k] : m;
    double sw = 0.0, n0 = 0.0, n1 = 0.0;
    for (int k = 0; k < K; ++k) {
      v[k] -= m;                                        /* :189 */
      const double w = exp(-v[k] / p->lam) + p->eps_floor; /* :193 */
      sw += w;
      n0 += e0[k] * w;
      n1 += e1[k] * w;
    }
    U[t] += n0 / sw;                                    /* :195-196 */
    U[T + t] += n1 / sw;
  }
  double* tmp = (double*)malloc(sizeof(double) * T);
  for (int c = 0; c < 2; ++c) {
    double* row = U + (size_t)c * T;
    for (int t = 0; t < T; ++t) row[t] = clipd(row[t], p->u_max[c]);   /* :198-199 */
    savgol_row(T, row, tmp);                                           /* :202 */
    for (int t = 0; t < T; ++t) row[t] = clipd(tmp[t], p->u_max[c]);   /* :205-206 */
  }
  free(tmp);
}

/* One get_path, control/src/mppi:85-102.  noise_mode 0: eps given; 1: drawn from the legacy MT19937
 * state `rs` exactly as the reference does (serial); 2: parallel counter stream (timing only).
 * work_eps (T*2*K) and work_V (T*K) are caller-provided scratch.  U is updated in place (shifted). */
void port_step(const port_params* p, const double x0[3], const double goal[3], double* U, int noise_mode,
               const double* eps_in, rk_state* rs, uint64_t seed, double* work_eps, double* work_V, double u0_out[2],
               double x_next[3], double* U_new_out) {
  const int K = p->K, T = p->T;
  const double* eps = eps_in;
  if (noise_mode == 1) {
    for (int t = 0; t < T; ++t) port_normal(rs, p->sig[0], work_eps + (size_t)t * 2 * K, 2L * K); /* :143-146 */
    eps = work_eps;
  } else if (noise_mode == 2) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)T * K; ++i) {
      const uint64_t a = splitmix64(seed ^ (uint64_t)i * 0x2545f4914f6cdd1dull), b = splitmix64(a);
      const double u1 = ((a >> 11) + 0.5) / 9007199254740992.0, u2 = ((b >> 11) + 0.5) / 9007199254740992.0;
      const double rr = sqrt(-2.0 * log(u1));
      const long t = i / K, k = i % K;
      work_eps[((size_t)t * 2 + 0) * K + k] = p->noise_std * rr * cos(2.0 * M_PI * u2);
      work_eps[((size_t)t * 2 + 1) * K + k] = p->noise_std * rr * sin(2.0 * M_PI * u2);
    }
    eps = work_eps;
  }
  port_cost2go(p, x0, U, goal, eps, work_V);           /* :90 */
  port_update_action(p, U, eps, work_V);               /* :92 */
  double x[3] = {x0[0], x0[1], x0[2]};
  rk4(p, x, U[0], U[T]);                               /* perform_action :94,210-213 */
  if (U_new_out) memcpy(U_new_out, U, sizeof(double) * 2 * T);
  u0_out[0] = U[0];
  u0_out[1] = U[T];
  x_next[0] = x[0];
  x_next[1] = x[1];
  x_next[2] = x[2];
  for (int c = 0; c < 2; ++c) {                        /* shift :100-101 */
    memmove(U + (size_t)c * T, U + (size_t)c * T + 1, sizeof(double) * (T - 1));
    U[(size_t)c * T + T - 1] = 0.0;
  }
}

int port_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
size_t port_sizeof_rk_state(void) { return sizeof(rk_state); }
size_t port_sizeof_params(void) { return sizeof(port_params); }
