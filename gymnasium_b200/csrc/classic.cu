// classic.cu -- fused step + TimeLimit + autoreset kernels for the remaining classic-control families (sm_100a):
// MountainCar-v0, MountainCarContinuous-v0, Pendulum-v1, Acrobot-v1.  One thread per env, struct-of-arrays float64 state,
// same prologue/epilogue as cartpole.cu.  HBM-bound elementwise kernels (40-90 B per env-step).
//
// Each kernel reproduces the dtype every intermediate has in the reference (NumPy 2 / NEP 50 promotion), because the
// reference mixes float32 actions/states with Python floats:
//   MountainCarEnv.step/reset              gymnasium/envs/classic_control/mountain_car.py:131-170   (float64 state)
//   Continuous_MountainCarEnv.step/reset   continuous_mountain_car.py:147-196  (float64 on the first step after a reset,
//                                          float32 afterwards: ``self.state = np.array([...], dtype=np.float32)`` :185)
//   PendulumEnv.step/reset                 pendulum.py:127-165 (float64 state, float32 torque)
//   AcrobotEnv.step/reset/_dsdt, rk4, wrap, bound   acrobot.py:176-283, :398-470 (float64 RK4 over a 5-vector)
// sin/cos/fmod are libdevice's (the reference's are libm/NumPy's): observations agree to ~1e-7, flags exactly.
#include "common.cuh"

namespace b2e {
namespace {

struct ClassicArgs {
  int64_t n, env_offset;
  int32_t max_steps, mode, rng_mode, act_dtype;
  uint64_t philox_seed, call_counter;
  double p0, p1, p2, p3;  // per-family parameters (reset bounds, goal_velocity, g ...)
  double* __restrict__ state;   // [k][n]
  uint8_t* __restrict__ sflag;  // [n] family flag (MountainCarContinuous: state currently holds float32 values)
  int32_t* __restrict__ ctrl;
  uint64_t* __restrict__ rng;
  float* __restrict__ obs;
  double* __restrict__ reward;
  uint8_t* __restrict__ term;
  uint8_t* __restrict__ trunc;
  float* __restrict__ final_obs;
  const void* __restrict__ actions;
  const uint8_t* __restrict__ mask;
};

struct Draw {
  Pcg64 g;
  bool numpy;
  uint64_t seed, env, counter;
  uint32_t k;
  __device__ __forceinline__ double next() {
    if (numpy) return g.next_double();
    const uint4 r = philox_block(seed, env, counter, 48u + (k >> 1));
    const double u = (k & 1u) ? u53_to_double(r.z, r.w) : u53_to_double(r.x, r.y);
    ++k;
    return u;
  }
  __device__ __forceinline__ double uniform(double lo, double range) { return __dadd_rn(lo, __dmul_rn(range, next())); }
};
__device__ __forceinline__ Draw make_draw(const ClassicArgs& a, int64_t i) {
  Draw D;
  D.numpy = a.rng_mode == B2E_RNG_NUMPY;
  if (D.numpy) D.g = pcg64_load(a.rng, a.n, i);
  D.seed = a.philox_seed; D.env = (uint64_t)(a.env_offset + i); D.counter = a.call_counter; D.k = 0;
  return D;
}
__device__ __forceinline__ void finish_draw(const ClassicArgs& a, int64_t i, Draw& D) {
  if (D.numpy) pcg64_store_state(a.rng, i, D.g);
}
template <typename T>
__device__ __forceinline__ float load_faction(const void* __restrict__ p, int64_t i) {
  return (float)__ldg(reinterpret_cast<const T*>(p) + i);
}

// ---- generic driver: Env supplies kState, kObs, reset(), step(), obs() ----------------------------------------------------
template <class Env>
__global__ void __launch_bounds__(kBlock) classic_reset_kernel(const ClassicArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  if (a.mask != nullptr && a.mask[i] == 0) return;
  Draw D = make_draw(a, i);
  double s[Env::kState];
  uint8_t flag = 0;
  Env::reset(a, D, s, flag);
  finish_draw(a, i, D);
#pragma unroll
  for (int k = 0; k < Env::kState; ++k) a.state[k * a.n + i] = s[k];
  if (a.sflag) a.sflag[i] = flag;
  a.ctrl[i] = 0;
  Env::obs(s, a.obs + Env::kObs * i);
}

template <class Env, typename ActT>
__global__ void __launch_bounds__(kBlock) classic_step_kernel(const ClassicArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int32_t c = a.ctrl[i];
  double s[Env::kState];
#pragma unroll
  for (int k = 0; k < Env::kState; ++k) s[k] = a.state[k * a.n + i];
  uint8_t flag = a.sflag ? a.sflag[i] : 0;
  const ActT act = __ldg(reinterpret_cast<const ActT*>(a.actions) + i);
  double reward = 0.0;
  bool term = false, trunc = false;
  int32_t cn = 0;
  if (a.mode == B2E_AUTORESET_NEXT_STEP && ctrl_pending(c)) {
    Draw D = make_draw(a, i);
    Env::reset(a, D, s, flag);
    finish_draw(a, i, D);
  } else {
    term = Env::step(a, s, flag, act, reward);
    const int32_t elapsed = ctrl_elapsed(c) + 1;
    trunc = a.max_steps > 0 && elapsed >= a.max_steps;
    cn = elapsed;
    if (term || trunc) {
      if (a.mode == B2E_AUTORESET_NEXT_STEP) {
        cn |= kPending;
      } else if (a.mode == B2E_AUTORESET_SAME_STEP) {
        Env::obs(s, a.final_obs + Env::kObs * i);
        Draw D = make_draw(a, i);
        Env::reset(a, D, s, flag);
        finish_draw(a, i, D);
        cn = 0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < Env::kState; ++k) a.state[k * a.n + i] = s[k];
  if (a.sflag) a.sflag[i] = flag;
  a.ctrl[i] = cn;
  a.reward[i] = reward;
  a.term[i] = term;
  a.trunc[i] = trunc;
  Env::obs(s, a.obs + Env::kObs * i);
}

// ---- MountainCar-v0 ------------------------------------------------------------------------------------------------------
struct MountainCar {
  static constexpr int kState = 2, kObs = 2;
  // p0 = reset low, p1 = reset high, p2 = goal_velocity
  __device__ static void reset(const ClassicArgs& a, Draw& D, double* s, uint8_t&) {
    s[0] = D.uniform(a.p0, a.p1 - a.p0);  // mountain_car.py:163
    s[1] = 0.0;
  }
  __device__ static bool step(const ClassicArgs& a, double* s, uint8_t&, int64_t action, double& reward) {
    double position = s[0], velocity = s[1];
    const int act = min(max((int)action, 0), 2);
    // velocity += (action - 1) * self.force + math.cos(3 * position) * (-self.gravity)      (:137)
    const double inc = __dadd_rn(__dmul_rn((double)(act - 1), 0.001), __dmul_rn(cos(__dmul_rn(3.0, position)), -0.0025));
    velocity = __dadd_rn(velocity, inc);
    velocity = fmin(fmax(velocity, -0.07), 0.07);
    position = __dadd_rn(position, velocity);
    position = fmin(fmax(position, -1.2), 0.6);
    if (position == -1.2 && velocity < 0) velocity = 0.0;
    reward = -1.0;
    s[0] = position; s[1] = velocity;
    return position >= 0.5 && velocity >= a.p2;
  }
  __device__ static void obs(const double* s, float* o) { o[0] = (float)s[0]; o[1] = (float)s[1]; }
};

// ---- MountainCarContinuous-v0 ----------------------------------------------------------------------------------------
struct MountainCarContinuous {
  static constexpr int kState = 2, kObs = 2;
  __device__ static void reset(const ClassicArgs& a, Draw& D, double* s, uint8_t& flag) {
    s[0] = D.uniform(a.p0, a.p1 - a.p0);  // continuous_mountain_car.py:190 (float64 array until the first step)
    s[1] = 0.0;
    flag = 0;
  }
  __device__ static bool step(const ClassicArgs& a, double* s, uint8_t& flag, float action, double& reward) {
    // force = min(max(action[0], -1.0), 1.0): stays np.float32 inside the bounds, becomes a Python float outside
    const bool force_f32 = action >= -1.0f && action <= 1.0f;
    const double force_py = action < -1.0f ? -1.0 : 1.0;
    bool terminated;
    if (!flag) {  // position/velocity are np.float64 (first step after a reset)
      double position = s[0], velocity = s[1];
      const double term2 = __dmul_rn(0.0025, cos(__dmul_rn(3.0, position)));
      if (force_f32) {
        const float x = __fsub_rn(__fmul_rn(action, 0.0015f), (float)term2);  // f32 - weak Python float -> f32
        velocity = __dadd_rn(velocity, (double)x);                           // f64 += f32 -> f64
      } else {
        velocity = __dadd_rn(velocity, __dsub_rn(__dmul_rn(force_py, 0.0015), term2));
      }
      if (velocity > 0.07) velocity = 0.07;
      if (velocity < -0.07) velocity = -0.07;
      position = __dadd_rn(position, velocity);
      if (position > 0.6) position = 0.6;
      if (position < -1.2) position = -1.2;
      if (position == -1.2 && velocity < 0) velocity = 0.0;
      terminated = position >= 0.45 && velocity >= a.p2;
      s[0] = (double)(float)position;  // np.array([position, velocity], dtype=np.float32)  (:185)
      s[1] = (double)(float)velocity;
    } else {  // np.float32 scalars: every operation rounds to float32, Python floats are cast to float32 first
      float position = (float)s[0], velocity = (float)s[1];
      const float p3 = __fmul_rn(3.0f, position);
      const double term2 = __dmul_rn(0.0025, cos((double)p3));
      if (force_f32) {
        const float x = __fsub_rn(__fmul_rn(action, 0.0015f), (float)term2);
        velocity = __fadd_rn(velocity, x);
      } else {
        velocity = __fadd_rn(velocity, (float)__dsub_rn(__dmul_rn(force_py, 0.0015), term2));
      }
      if (velocity > 0.07f) velocity = 0.07f;
      if (velocity < -0.07f) velocity = -0.07f;
      position = __fadd_rn(position, velocity);
      if (position > 0.6f) position = 0.6f;
      if (position < -1.2f) position = -1.2f;
      if (position == -1.2f && velocity < 0.0f) velocity = 0.0f;
      terminated = position >= 0.45f && (double)velocity >= a.p2;
      s[0] = (double)position;
      s[1] = (double)velocity;
    }
    flag = 1;
    // reward = (100.0 if terminated else 0) - math.pow(action[0], 2) * 0.1            (:177-180)
    const double a64 = (double)action;
    reward = __dsub_rn(terminated ? 100.0 : 0.0, __dmul_rn(__dmul_rn(a64, a64), 0.1));
    return terminated;
  }
  __device__ static void obs(const double* s, float* o) { o[0] = (float)s[0]; o[1] = (float)s[1]; }
};

// ---- Pendulum-v1 ----------------------------------------------------------------------------------------------------------
struct Pendulum {
  static constexpr int kState = 2, kObs = 3;
  // p0 = x_init (high[0]), p1 = y_init (high[1]), p2 = g
  __device__ static void reset(const ClassicArgs& a, Draw& D, double* s, uint8_t&) {
    s[0] = D.uniform(-a.p0, a.p0 - -a.p0);  // np_random.uniform(low=-high, high=high)  (pendulum.py:159-160)
    s[1] = D.uniform(-a.p1, a.p1 - -a.p1);
  }
  __device__ static bool step(const ClassicArgs& a, double* s, uint8_t&, float action, double& reward) {
    const double th = s[0], thdot = s[1];
    const double kPi = 3.141592653589793;
    const float u = fminf(fmaxf(action, -2.0f), 2.0f);  // np.clip(u, -max_torque, max_torque)[0] -> np.float32
    // costs = angle_normalize(th) ** 2 + 0.1 * thdot**2 + 0.001 * (u**2)                     (:137)
    double m = fmod(__dadd_rn(th, kPi), 2.0 * kPi);  // np.remainder: fmod + sign fix-up
    if (m != 0.0) {
      if (m < 0.0) m = __dadd_rn(m, 2.0 * kPi);
    } else {
      m = 0.0;
    }
    const double an = __dsub_rn(m, kPi);
    const float u2 = __fmul_rn(0.001f, __fmul_rn(u, u));  // float32 ** 2 -> float32; weak 0.001 -> float32
    const double costs = __dadd_rn(__dadd_rn(__dmul_rn(an, an), __dmul_rn(0.1, __dmul_rn(thdot, thdot))), (double)u2);
    // newthdot = thdot + (3 * g / (2 * l) * np.sin(th) + 3.0 / (m * l**2) * u) * dt            (:139)
    const double k1 = __ddiv_rn(__dmul_rn(3.0, a.p2), 2.0);
    const float tu = __fmul_rn(3.0f, u);
    double newthdot = __dadd_rn(thdot, __dmul_rn(__dadd_rn(__dmul_rn(k1, sin(th)), (double)tu), 0.05));
    newthdot = fmin(fmax(newthdot, -8.0), 8.0);
    s[0] = __dadd_rn(th, __dmul_rn(newthdot, 0.05));
    s[1] = newthdot;
    reward = -costs;
    return false;
  }
  __device__ static void obs(const double* s, float* o) {
    o[0] = (float)cos(s[0]); o[1] = (float)sin(s[0]); o[2] = (float)s[1];
  }
};

// ---- Acrobot-v1 -----------------------------------------------------------------------------------------------------------
struct Acrobot {
  static constexpr int kState = 4, kObs = 6;
  __device__ static void reset(const ClassicArgs& a, Draw& D, double* s, uint8_t&) {
    for (int k = 0; k < 4; ++k) s[k] = (double)(float)D.uniform(a.p0, a.p1 - a.p0);  // .astype(np.float32)  (acrobot.py:186-188)
  }
  __device__ static void dsdt(const double* s, double torque, double* d) {  // acrobot.py:237-283 ("book")
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 0.5, I1 = 1.0, I2 = 1.0, g = 9.8, pi = 3.141592653589793;
    const double theta1 = s[0], theta2 = s[1], dtheta1 = s[2], dtheta2 = s[3];
    const double c2 = cos(theta2), s2 = sin(theta2);
    // d1 = m1*lc1**2 + m2*(l1**2 + lc2**2 + 2*l1*lc2*cos(theta2)) + I1 + I2
    const double d1 = __dadd_rn(__dadd_rn(__dadd_rn(m1 * lc1 * lc1, __dmul_rn(m2, __dadd_rn(l1 * l1 + lc2 * lc2, __dmul_rn(2 * l1 * lc2, c2)))), I1), I2);
    // d2 = m2*(lc2**2 + l1*lc2*cos(theta2)) + I2
    const double d2 = __dadd_rn(__dmul_rn(m2, __dadd_rn(lc2 * lc2, __dmul_rn(l1 * lc2, c2))), I2);
    // phi2 = m2*lc2*g*cos(theta1 + theta2 - pi/2.0)
    const double phi2 = __dmul_rn(m2 * lc2 * g, cos(__dsub_rn(__dadd_rn(theta1, theta2), pi / 2.0)));
    // phi1 = -m2*l1*lc2*dtheta2**2*sin(theta2) - 2*m2*l1*lc2*dtheta2*dtheta1*sin(theta2) + (m1*lc1 + m2*l1)*g*cos(theta1 - pi/2) + phi2
    const double t1 = __dmul_rn(__dmul_rn(-m2 * l1 * lc2, __dmul_rn(dtheta2, dtheta2)), s2);
    const double t2 = __dmul_rn(__dmul_rn(__dmul_rn(2 * m2 * l1 * lc2, dtheta2), dtheta1), s2);
    const double t3 = __dmul_rn((m1 * lc1 + m2 * l1) * g, cos(__dsub_rn(theta1, pi / 2)));
    const double phi1 = __dadd_rn(__dadd_rn(__dsub_rn(t1, t2), t3), phi2);
    // ddtheta2 = (a + d2/d1*phi1 - m2*l1*lc2*dtheta1**2*sin(theta2) - phi2) / (m2*lc2**2 + I2 - d2**2/d1)
    const double num = __dsub_rn(__dsub_rn(__dadd_rn(torque, __dmul_rn(__ddiv_rn(d2, d1), phi1)),
                                           __dmul_rn(__dmul_rn(m2 * l1 * lc2, __dmul_rn(dtheta1, dtheta1)), s2)), phi2);
    const double den = __dsub_rn(m2 * lc2 * lc2 + I2, __ddiv_rn(__dmul_rn(d2, d2), d1));
    const double ddtheta2 = __ddiv_rn(num, den);
    const double ddtheta1 = __ddiv_rn(-__dadd_rn(__dmul_rn(d2, ddtheta2), phi1), d1);
    d[0] = dtheta1; d[1] = dtheta2; d[2] = ddtheta1; d[3] = ddtheta2;
  }
  __device__ static double wrap(double x, double m, double M) {  // acrobot.py:398-418
    const double diff = __dsub_rn(M, m);
    while (x > M) x = __dsub_rn(x, diff);
    while (x < m) x = __dadd_rn(x, diff);
    return x;
  }
  __device__ static bool step(const ClassicArgs& a, double* s, uint8_t&, int64_t action, double& reward) {
    const double pi = 3.141592653589793;
    const int act = min(max((int)action, 0), 2);
    const double torque = act == 0 ? -1.0 : act == 1 ? 0.0 : 1.0;  // AVAIL_TORQUE (:160)
    // rk4(self._dsdt, s_augmented, [0, 0.2])  (acrobot.py:421-470): dt = 0.2, dt2 = 0.1
    const double dt = 0.2, dt2 = dt / 2.0;
    double k1[4], k2[4], k3[4], k4[4], y[4];
    dsdt(s, torque, k1);
    for (int k = 0; k < 4; ++k) y[k] = __dadd_rn(s[k], __dmul_rn(dt2, k1[k]));
    dsdt(y, torque, k2);
    for (int k = 0; k < 4; ++k) y[k] = __dadd_rn(s[k], __dmul_rn(dt2, k2[k]));
    dsdt(y, torque, k3);
    for (int k = 0; k < 4; ++k) y[k] = __dadd_rn(s[k], __dmul_rn(dt, k3[k]));
    dsdt(y, torque, k4);
    double ns[4];
    for (int k = 0; k < 4; ++k)  // y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
      ns[k] = __dadd_rn(s[k], __dmul_rn(dt / 6.0, __dadd_rn(__dadd_rn(__dadd_rn(k1[k], __dmul_rn(2.0, k2[k])), __dmul_rn(2.0, k3[k])), k4[k])));
    ns[0] = wrap(ns[0], -pi, pi);
    ns[1] = wrap(ns[1], -pi, pi);
    ns[2] = fmin(fmax(ns[2], -4 * pi), 4 * pi);  // bound(): min(max(x, m), M)
    ns[3] = fmin(fmax(ns[3], -9 * pi), 9 * pi);
    for (int k = 0; k < 4; ++k) s[k] = ns[k];
    const bool terminated = __dsub_rn(-cos(s[0]), cos(__dadd_rn(s[1], s[0]))) > 1.0;  // :228-231
    reward = terminated ? 0.0 : -1.0;
    return terminated;
  }
  __device__ static void obs(const double* s, float* o) {
    o[0] = (float)cos(s[0]); o[1] = (float)sin(s[0]); o[2] = (float)cos(s[1]); o[3] = (float)sin(s[1]);
    o[4] = (float)s[2]; o[5] = (float)s[3];
  }
};

template <class Env, typename ActT>
int launch_step(const ClassicArgs& a, cudaStream_t st) {
  classic_step_kernel<Env, ActT><<<grid_for(a.n), kBlock, 0, st>>>(a);
  return cuda_status(cudaGetLastError(), "b2e_classic_step");
}
template <class Env>
int dispatch(const ClassicArgs& a, bool discrete, cudaStream_t st) {
  if (discrete) {
    switch (a.act_dtype) {
      case B2E_ACT_I64: return launch_step<Env, int64_t>(a, st);
      case B2E_ACT_I32: return launch_step<Env, int32_t>(a, st);
      case B2E_ACT_U8: return launch_step<Env, uint8_t>(a, st);
    }
  }
  set_error("b2e_classic_step: action_dtype %d does not fit this family", a.act_dtype);
  return B2E_EINVAL;
}
template <class Env>
int dispatch_float(const ClassicArgs& a, cudaStream_t st) {
  if (a.act_dtype == B2E_ACT_F32) {
    classic_step_kernel<Env, float><<<grid_for(a.n), kBlock, 0, st>>>(a);
    return cuda_status(cudaGetLastError(), "b2e_classic_step");
  }
  set_error("b2e_classic_step: action_dtype %d must be float32 for this family", a.act_dtype);
  return B2E_EINVAL;
}

int fill(const b2e_batch* b, const b2e_classic_cfg* cfg, ClassicArgs& a, const char* fn) {
  if (int e = check_batch(b, fn)) return e;
  if (!cfg || !cfg->state || !cfg->ctrl || cfg->family < 0 || cfg->family > 3 ||
      (b->rng_mode == B2E_RNG_NUMPY && !cfg->rng) || (cfg->family == B2E_CLASSIC_MOUNTAINCAR_CONTINUOUS && !cfg->sflag)) {
    set_error("%s: bad configuration", fn);
    return B2E_EINVAL;
  }
  a = ClassicArgs{};
  a.n = b->n; a.env_offset = b->env_offset; a.max_steps = b->max_episode_steps; a.mode = b->autoreset_mode;
  a.rng_mode = b->rng_mode; a.act_dtype = b->action_dtype; a.philox_seed = b->philox_seed; a.call_counter = b->call_counter;
  a.p0 = cfg->p[0]; a.p1 = cfg->p[1]; a.p2 = cfg->p[2]; a.p3 = cfg->p[3];
  a.state = cfg->state; a.sflag = cfg->sflag; a.ctrl = cfg->ctrl; a.rng = cfg->rng;
  return 0;
}

}  // namespace
}  // namespace b2e

using namespace b2e;

extern "C" int b2e_classic_reset(const b2e_batch* b, const b2e_classic_cfg* cfg, const uint8_t* mask, float* obs,
                                 void* stream) {
  ClassicArgs a;
  if (int e = fill(b, cfg, a, "b2e_classic_reset")) return e;
  if (!obs) {
    set_error("b2e_classic_reset: obs is NULL");
    return B2E_EINVAL;
  }
  if (b->n == 0) return 0;
  a.mask = mask; a.obs = obs;
  cudaStream_t st = (cudaStream_t)stream;
  switch (cfg->family) {
    case B2E_CLASSIC_MOUNTAINCAR: classic_reset_kernel<MountainCar><<<grid_for(a.n), kBlock, 0, st>>>(a); break;
    case B2E_CLASSIC_MOUNTAINCAR_CONTINUOUS: classic_reset_kernel<MountainCarContinuous><<<grid_for(a.n), kBlock, 0, st>>>(a); break;
    case B2E_CLASSIC_PENDULUM: classic_reset_kernel<Pendulum><<<grid_for(a.n), kBlock, 0, st>>>(a); break;
    default: classic_reset_kernel<Acrobot><<<grid_for(a.n), kBlock, 0, st>>>(a); break;
  }
  return cuda_status(cudaGetLastError(), "b2e_classic_reset");
}

extern "C" int b2e_classic_step(const b2e_batch* b, const b2e_classic_cfg* cfg, const void* actions, float* obs,
                                double* reward, uint8_t* terminated, uint8_t* truncated, float* final_obs, void* stream) {
  ClassicArgs a;
  if (int e = fill(b, cfg, a, "b2e_classic_step")) return e;
  if (!actions || !obs || !reward || !terminated || !truncated ||
      (b->autoreset_mode == B2E_AUTORESET_SAME_STEP && !final_obs)) {
    set_error("b2e_classic_step: null pointer");
    return B2E_EINVAL;
  }
  if (b->n == 0) return 0;
  a.actions = actions; a.obs = obs; a.reward = reward; a.term = terminated; a.trunc = truncated; a.final_obs = final_obs;
  cudaStream_t st = (cudaStream_t)stream;
  switch (cfg->family) {
    case B2E_CLASSIC_MOUNTAINCAR: return dispatch<MountainCar>(a, true, st);
    case B2E_CLASSIC_MOUNTAINCAR_CONTINUOUS: return dispatch_float<MountainCarContinuous>(a, st);
    case B2E_CLASSIC_PENDULUM: return dispatch_float<Pendulum>(a, st);
    default: return dispatch<Acrobot>(a, true, st);
  }
}
