/* oracle/lunar_lander.c -- CPU oracle for LunarLander-v3: a plain-C restatement of the Box2D 2.3.x subset that
 * gymnasium/envs/box2d/lunar_lander.py drives, plus the environment logic itself.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into or called by gymnasium_b200/.
 *
 * PARITY UNPINNED.  The arithmetic of this env lives in the third-party wheel `box2d ==2.3.10` (pybox2d, SWIG over
 * Box2D C++ 2.3.x; pyproject.toml:38-43 of the reference).  That wheel is not vendored in /root/reference, not
 * installed in this image and not installable (no network, no swig), and the reference holds no golden vectors for
 * this env.  This file therefore restates Box2D's published algorithm from its documented structure
 * (b2World::Step -> b2ContactManager::Collide -> b2Island::Solve -> b2ContactSolver / b2RevoluteJoint, b2CollideEdge-
 * AndPolygon, b2PolygonShape::ComputeMass, b2DynamicTree fat AABBs) and anchors on the reference's own call sites:
 *   reset      gymnasium/envs/box2d/lunar_lander.py:321-447
 *   step       lunar_lander.py:471-665 (world.Step(1/50, 180, 60) at :619)
 *   contacts   lunar_lander.py:58-76 (ContactDetector Begin/EndContact)
 * The only behavioural pin the reference offers for this boundary is the heuristic landing test
 * (tests/envs/test_env_implementation.py:12-16 with the policy at lunar_lander.py:791-842); tests/test_oracle_lunar.py
 * runs it against this file.
 * b2World::SolveTOI (continuous collision of the three dynamic bodies against the static terrain: b2TimeOfImpact with its
 * b2Distance / GJK core and b2SeparationFunction root finder, the TOI sub-step island with its own position and velocity
 * solve) IS restated -- world.Step runs it after the discrete solve on every call (lunar_lander.py:619).
 *
 * All Box2D arithmetic is float32 with one rounding per operation (compile with -ffp-contract=off); the Python-side
 * glue (dispersion, impulses, observation scaling, shaping reward) is float64 as in the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------------
 * numpy Generator(PCG64(SeedSequence(seed))) -- gymnasium/utils/seeding.py:39-41 (see oracle/np_rng.py for the pinned
 * Python restatement; this is the same algorithm in C) */
typedef unsigned __int128 u128;
typedef struct { u128 state, inc; int seeded; int has32; uint32_t word; /* pcg64_next32's one-word buffer */ } pcg64_t;

static void pcg64_seed(pcg64_t* g, uint64_t seed) {
  const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
  const uint32_t MIX_L = 0xca01f9ddu, MIX_R = 0x4973f715u;
  uint32_t hc = INIT_A, pool[4], words[4] = {(uint32_t)seed, (uint32_t)(seed >> 32), 0, 0};
  for (int i = 0; i < 4; ++i) { uint32_t v = words[i] ^ hc; hc *= MULT_A; v *= hc; v ^= v >> 16; pool[i] = v; }
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < 4; ++d)
      if (s != d) {
        uint32_t v = pool[s] ^ hc; hc *= MULT_A; v *= hc; v ^= v >> 16;
        uint32_t r = MIX_L * pool[d] - MIX_R * v; r ^= r >> 16; pool[d] = r;
      }
  uint32_t out[8], hb = INIT_B;
  for (int i = 0; i < 8; ++i) { uint32_t v = pool[i & 3] ^ hb; hb *= MULT_B; v *= hb; v ^= v >> 16; out[i] = v; }
  uint64_t w[4];
  for (int k = 0; k < 4; ++k) w[k] = out[2 * k] | ((uint64_t)out[2 * k + 1] << 32);
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
  u128 initstate = ((u128)w[0] << 64) | w[1], initseq = ((u128)w[2] << 64) | w[3];
  g->inc = (initseq << 1) | 1;
  g->state = 0;
  g->state = g->state * mult + g->inc;
  g->state += initstate;
  g->state = g->state * mult + g->inc;
  g->seeded = 1;
  g->has32 = 0; g->word = 0; /* a fresh Generator starts with an empty 32-bit buffer */
}
static double pcg64_double(pcg64_t* g) {
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
  g->state = g->state * mult + g->inc;
  uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state, x = hi ^ lo;
  unsigned rot = (unsigned)(hi >> 58);
  x = (x >> rot) | (x << ((64 - rot) & 63));
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
static double pcg64_uniform(pcg64_t* g, double lo, double hi) { return lo + (hi - lo) * pcg64_double(g); }
static uint64_t pcg64_next64(pcg64_t* g) {
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
  g->state = g->state * mult + g->inc;
  uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state, x = hi ^ lo;
  unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((64 - rot) & 63));
}
/* pcg64_next32 (numpy/random/src/pcg64/pcg64.h): low half of a fresh 64-bit draw first, the high half on the next call */
static uint32_t pcg64_next32(pcg64_t* g) {
  if (g->has32) { g->has32 = 0; return g->word; }
  uint64_t x = pcg64_next64(g);
  g->has32 = 1; g->word = (uint32_t)(x >> 32);
  return (uint32_t)x;
}
/* Generator.integers(low, high) for a range below 2^32: buffered_bounded_lemire_uint32 (numpy/random/src/distributions/
 * distributions.c), pinned against numpy in oracle/np_rng.py (bounded_uint32) */
static int64_t pcg64_integers(pcg64_t* g, int64_t low, int64_t high) {
  const uint32_t rng = (uint32_t)(high - low - 1), rng_excl = rng + 1u;
  uint64_t m = (uint64_t)pcg64_next32(g) * rng_excl;
  uint32_t leftover = (uint32_t)m;
  if (leftover < rng_excl) {
    const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
    while (leftover < threshold) { m = (uint64_t)pcg64_next32(g) * rng_excl; leftover = (uint32_t)m; }
  }
  return low + (int64_t)(m >> 32);
}

/* ------------------------------------------------------------------------------------------------------------------
 * b2Math */
typedef struct { float x, y; } v2;
typedef struct { float s, c; } rot_t;
typedef struct { v2 p; rot_t q; } xf_t;
typedef struct { v2 lo, hi; } aabb_t;

static inline v2 V(float x, float y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vneg(v2 a) { return V(-a.x, -a.y); }
static inline v2 vscale(float s, v2 a) { return V(s * a.x, s * a.y); }
static inline float vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline float vcross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }
static inline v2 vcross_vs(v2 a, float s) { return V(s * a.y, -s * a.x); } /* b2Cross(vec, s) */
static inline v2 vcross_sv(float s, v2 a) { return V(-s * a.y, s * a.x); } /* b2Cross(s, vec) */
static inline float vlen(v2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
static inline v2 vmin(v2 a, v2 b) { return V(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y); }
static inline v2 vmax(v2 a, v2 b) { return V(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y); }
static inline float fclamp(float a, float lo, float hi) { return fmaxf(lo, fminf(a, hi)); }
/* sin/cos in double from a fixed sequence of IEEE operations (Cody-Waite reduction by pi/2 in two pieces + the classic
 * fdlibm kernel polynomials, no FMA): the CUDA engine executes the very same sequence, so both sides agree bit for bit
 * without depending on any libm.  Accurate to ~1e-16, i.e. its float rounding equals a correctly rounded sinf/cosf
 * (what Box2D gets from glibc) except on ~1e-9 of inputs. */
static inline void det_sincos(double x, double* sn, double* cs) {
  const double fn = rint(x * 6.36619772367581382433e-01);
  const double r = x - fn * 1.57079632673412561417e+00;
  const double w = fn * 6.07710050650619224932e-11;
  const double y = r - w;
  const double z = y * y;
  const double ps = 8.33333333332248946124e-03 +
                    z * (-1.98412698298579493134e-04 +
                         z * (2.75573137070700676789e-06 + z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10)));
  const double sk = y + (z * y) * (-1.66666666666666324348e-01 + z * ps);
  const double pc = z * (4.16666666666666019037e-02 +
                         z * (-1.38888888888741095749e-03 +
                              z * (2.48015872894767294178e-05 +
                                   z * (-2.75573143513906633035e-07 +
                                        z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11)))));
  const double ck = 1.0 - (0.5 * z - z * pc);
  switch ((int)fn & 3) {
    case 0: *sn = sk; *cs = ck; break;
    case 1: *sn = ck; *cs = -sk; break;
    case 2: *sn = -sk; *cs = -ck; break;
    default: *sn = -ck; *cs = sk; break;
  }
}
/* b2Rot::Set(angle): sinf/cosf */
static inline rot_t rot_set(float a) {
  double s, c;
  det_sincos((double)a, &s, &c);
  rot_t r = {(float)s, (float)c};
  return r;
}
/* tanh in double from a fixed sequence of IEEE operations (the CUDA engine runs the same one): exp by Cody-Waite reduction
 * with ln 2 in two pieces + a degree-13 Taylor polynomial, tanh = (e^2|x| - 1) / (e^2|x| + 1); accurate to ~2e-16 for the
 * |x| <= 2 the wind function feeds it, i.e. its float rounding equals libm's except on ~1e-8 of inputs. */
static inline double det_tanh(double x) {
  const double ax = fabs(x);
  double r;
  if (ax < 1e-3) {
    r = ax * (1.0 - (ax * ax) * (1.0 / 3.0));
  } else if (ax > 20.0) {
    r = 1.0;
  } else {
    const double t = 2.0 * ax;
    const double k = rint(t * 1.44269504088896338700e+00);
    const double y = (t - k * 6.93147180369123816490e-01) - k * 1.90821492927058770002e-10;
    double p = 1.0 / 6227020800.0;
    p = p * y + 1.0 / 479001600.0;
    p = p * y + 1.0 / 39916800.0;
    p = p * y + 1.0 / 3628800.0;
    p = p * y + 1.0 / 362880.0;
    p = p * y + 1.0 / 40320.0;
    p = p * y + 1.0 / 5040.0;
    p = p * y + 1.0 / 720.0;
    p = p * y + 1.0 / 120.0;
    p = p * y + 1.0 / 24.0;
    p = p * y + 1.0 / 6.0;
    p = p * y + 0.5;
    p = p * y + 1.0;
    p = p * y + 1.0;
    const double e = p * (double)(1ull << (int)k); /* 0 <= k <= 58 */
    r = (e - 1.0) / (e + 1.0);
  }
  return x < 0 ? -r : r;
}
static inline v2 rmul(rot_t q, v2 v) { return V(q.c * v.x - q.s * v.y, q.s * v.x + q.c * v.y); }
static inline v2 rmulT(rot_t q, v2 v) { return V(q.c * v.x + q.s * v.y, -q.s * v.x + q.c * v.y); }
static inline v2 xmul(xf_t T, v2 v) { return V((T.q.c * v.x - T.q.s * v.y) + T.p.x, (T.q.s * v.x + T.q.c * v.y) + T.p.y); }
static inline v2 xmulT(xf_t T, v2 v) {
  float px = v.x - T.p.x, py = v.y - T.p.y;
  return V(T.q.c * px + T.q.s * py, -T.q.s * px + T.q.c * py);
}
/* b2Vec2::Normalize */
static inline float vnormalize(v2* a) {
  float len = vlen(*a);
  if (len < 1.19209290e-7f) return 0.0f;
  float inv = 1.0f / len;
  a->x *= inv; a->y *= inv;
  return len;
}

/* b2Settings.h */
#define B2_PI 3.14159265359f
#define LINEAR_SLOP 0.005f
#define ANGULAR_SLOP (2.0f / 180.0f * B2_PI)
#define POLYGON_RADIUS (2.0f * LINEAR_SLOP)
#define AABB_EXTENSION 0.1f
#define AABB_MULTIPLIER 2.0f
#define VELOCITY_THRESHOLD 1.0f
#define MAX_LINEAR_CORRECTION 0.2f
#define MAX_ANGULAR_CORRECTION (8.0f / 180.0f * B2_PI)
#define MAX_TRANSLATION 2.0f
#define MAX_ROTATION (0.5f * B2_PI)
#define BAUMGARTE 0.2f
#define TIME_TO_SLEEP 0.5f
#define LINEAR_SLEEP_TOL 0.01f
#define ANGULAR_SLEEP_TOL (2.0f / 180.0f * B2_PI)

/* ------------------------------------------------------------------------------------------------------------------
 * scene: body 0 = moon (static, 11 edge fixtures), 1 = lander, 2 = legs[0] (i=-1), 3 = legs[1] (i=+1) */
#define NBODY 4
#define NEDGE 11  /* fixture 0: base edge (0,0)-(W,0); 1..10: terrain */
#define NDYN 3
#define NPAIR (NDYN * NEDGE)
#define MAXV 8

typedef struct {
  int count;
  v2 v[MAXV], n[MAXV], centroid;
} poly_t;

typedef struct {
  int dynamic;
  xf_t xf;
  v2 local_center, c0, c;
  float a0, a;
  v2 vel;
  float w;
  v2 force;
  float torque, mass, inv_mass, I, inv_I, sleep_time;
  int awake;
  float alpha0;               /* b2Sweep::alpha0 (continuous collision) */
} body_t;

typedef struct { uint8_t indexA, indexB, typeA, typeB; } cid_t; /* b2ContactFeature */
static inline uint32_t cid_key(cid_t c) { return c.indexA | (c.indexB << 8) | (c.typeA << 16) | ((uint32_t)c.typeB << 24); }

typedef struct { v2 local_point; float normal_impulse, tangent_impulse; cid_t id; } mpoint_t;
typedef struct { mpoint_t points[2]; v2 local_normal, local_point; int type /*1 faceA, 2 faceB*/, count; } manifold_t;

typedef struct {
  int exists, touching;
  int enabled, toi_flag, toi_count; /* b2Contact e_enabledFlag / e_toiFlag / m_toiCount */
  float toi;                        /* m_toi: cached time of impact of this step */
  long seq;          /* creation order: lists are walked most-recent-first like Box2D's intrusive lists */
  float friction;
  manifold_t m;
} contact_t;

typedef struct {
  int bodyB;                 /* bodyA is the lander (1) */
  v2 local_anchor_a, local_anchor_b;
  float reference_angle, lower, upper, max_motor_torque, motor_speed;
  float imp_x, imp_y, imp_z, motor_impulse;
  int limit_state;           /* 0 inactive, 1 lower, 2 upper, 3 equal */
  /* solver temp */
  v2 rA, rB, lcA, lcB;
  float m[3][3];             /* m_mass columns ex, ey, ez: m[col][row] */
  float motor_mass, mA, mB, iA, iB;
} joint_t;

typedef struct {
  body_t b[NBODY];
  poly_t poly[NDYN];          /* lander, leg0, leg1 (body-local) */
  v2 edge_v1[NEDGE], edge_v2[NEDGE];
  float edge_friction[NEDGE], poly_friction[NDYN];
  aabb_t edge_fat[NEDGE], poly_fat[NDYN];
  contact_t contact[NPAIR];   /* index = dyn * NEDGE + edge */
  long seq;
  joint_t joint[2];           /* joint[k] connects lander and legs[k] */
  float inv_dt0;
  int new_fixture;
  /* env */
  int game_over, leg_contact[2], has_prev_shaping;
  long toi_calls, toi_events;  /* statistics: b2TimeOfImpact evaluations / solid TOI events (sub-steps) since reset */
  long toi_clear, toi_clear_wrong; /* evaluations the CUDA engine's "clearly separated" test would skip / of those, how
                                      many b2TimeOfImpact did NOT answer alpha = 1 (must stay 0: the skip is exact) */
  double prev_shaping, helipad_y;
  float gravity;
  /* LunarLander(continuous, enable_wind, wind_power, turbulence_power): configuration (kept across resets) and the wind
   * pattern offsets drawn at every reset (lunar_lander.py:401-403) */
  int continuous, enable_wind;
  double wind_power, turbulence_power;
  int64_t wind_idx, torque_idx;
  pcg64_t rng;
} lander_t;

/* --- b2PolygonShape::Set (hull) + ComputeCentroid ------------------------------------------------------------- */
static v2 compute_centroid(const v2* vs, int count) {
  v2 c = V(0, 0), pRef = V(0, 0);
  float area = 0.0f;
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < count; ++i) {
    v2 p1 = pRef, p2 = vs[i], p3 = i + 1 < count ? vs[i + 1] : vs[0];
    v2 e1 = vsub(p2, p1), e2 = vsub(p3, p1);
    float D = vcross(e1, e2), tri = 0.5f * D;
    area += tri;
    c = vadd(c, vscale(tri * inv3, vadd(vadd(p1, p2), p3)));
  }
  return vscale(1.0f / area, c);
}
static void poly_set(poly_t* p, const v2* in, int n) {
  v2 ps[MAXV];
  int cnt = 0;
  for (int i = 0; i < n; ++i) { /* weld close points */
    int unique = 1;
    for (int j = 0; j < cnt; ++j) {
      v2 d = vsub(in[i], ps[j]);
      if (vdot(d, d) < (0.5f * LINEAR_SLOP) * (0.5f * LINEAR_SLOP)) { unique = 0; break; }
    }
    if (unique) ps[cnt++] = in[i];
  }
  int i0 = 0;
  float x0 = ps[0].x;
  for (int i = 1; i < cnt; ++i) {
    float x = ps[i].x;
    if (x > x0 || (x == x0 && ps[i].y < ps[i0].y)) { i0 = i; x0 = x; }
  }
  int hull[MAXV], m = 0, ih = i0;
  for (;;) {
    hull[m] = ih;
    int ie = 0;
    for (int j = 1; j < cnt; ++j) {
      if (ie == ih) { ie = j; continue; }
      v2 r = vsub(ps[ie], ps[hull[m]]), v = vsub(ps[j], ps[hull[m]]);
      float c = vcross(r, v);
      if (c < 0.0f) ie = j;
      if (c == 0.0f && vdot(v, v) > vdot(r, r)) ie = j;
    }
    ++m;
    ih = ie;
    if (ie == i0) break;
  }
  p->count = m;
  for (int i = 0; i < m; ++i) p->v[i] = ps[hull[i]];
  for (int i = 0; i < m; ++i) {
    int i2 = i + 1 < m ? i + 1 : 0;
    v2 e = vsub(p->v[i2], p->v[i]);
    p->n[i] = vcross_vs(e, 1.0f);
    vnormalize(&p->n[i]);
  }
  p->centroid = compute_centroid(p->v, m);
}
static void poly_set_box(poly_t* p, float hx, float hy) {
  p->count = 4;
  p->v[0] = V(-hx, -hy); p->v[1] = V(hx, -hy); p->v[2] = V(hx, hy); p->v[3] = V(-hx, hy);
  p->n[0] = V(0, -1); p->n[1] = V(1, 0); p->n[2] = V(0, 1); p->n[3] = V(-1, 0);
  p->centroid = V(0, 0);
}
/* b2PolygonShape::ComputeMass */
static void poly_mass(const poly_t* p, float density, float* mass, v2* center_out, float* I_out) {
  v2 center = V(0, 0), s = V(0, 0);
  float area = 0.0f, I = 0.0f;
  for (int i = 0; i < p->count; ++i) s = vadd(s, p->v[i]);
  s = vscale(1.0f / (float)p->count, s);
  const float inv3 = 1.0f / 3.0f;
  for (int i = 0; i < p->count; ++i) {
    v2 e1 = vsub(p->v[i], s), e2 = vsub(i + 1 < p->count ? p->v[i + 1] : p->v[0], s);
    float D = vcross(e1, e2), tri = 0.5f * D;
    area += tri;
    center = vadd(center, vscale(tri * inv3, vadd(e1, e2)));
    float ex1 = e1.x, ey1 = e1.y, ex2 = e2.x, ey2 = e2.y;
    float intx2 = ex1 * ex1 + ex2 * ex1 + ex2 * ex2, inty2 = ey1 * ey1 + ey2 * ey1 + ey2 * ey2;
    I += (0.25f * inv3 * D) * (intx2 + inty2);
  }
  *mass = density * area;
  center = vscale(1.0f / area, center);
  *center_out = vadd(center, s);
  *I_out = density * I;
  *I_out += *mass * (vdot(*center_out, *center_out) - vdot(center, center));
}

/* b2Body::ResetMassData for a one-fixture body, then the sweep set-up b2Body does at creation */
static void body_init_dynamic(body_t* b, const poly_t* p, float density, v2 pos, float angle) {
  memset(b, 0, sizeof(*b));
  b->dynamic = 1;
  b->awake = 1;
  b->xf.p = pos;
  b->xf.q = rot_set(angle);
  b->c0 = b->c = pos;
  b->a0 = b->a = angle;
  float mass, I;
  v2 center;
  poly_mass(p, density, &mass, &center, &I);
  b->mass = mass;
  b->inv_mass = 1.0f / mass;
  v2 lc = vscale(b->inv_mass, vscale(mass, center)); /* localCenter = sum(mass_i * center_i) * invMass */
  I -= mass * vdot(lc, lc);
  b->I = I;
  b->inv_I = 1.0f / I;
  b->local_center = lc;
  b->c0 = b->c = xmul(b->xf, lc);
}
static void sync_transform(body_t* b) {
  b->xf.q = rot_set(b->a);
  b->xf.p = vsub(b->c, rmul(b->xf.q, b->local_center));
}
static void set_awake(body_t* b, int flag) {
  if (flag) {
    if (!b->awake) { b->awake = 1; b->sleep_time = 0.0f; }
  } else {
    b->awake = 0; b->sleep_time = 0.0f; b->vel = V(0, 0); b->w = 0.0f; b->force = V(0, 0); b->torque = 0.0f;
  }
}

/* --- AABBs / broad phase ---------------------------------------------------------------------------------------- */
static aabb_t poly_aabb(const poly_t* p, xf_t xf) {
  v2 lo = xmul(xf, p->v[0]), hi = lo;
  for (int i = 1; i < p->count; ++i) { v2 v = xmul(xf, p->v[i]); lo = vmin(lo, v); hi = vmax(hi, v); }
  aabb_t r = {V(lo.x - POLYGON_RADIUS, lo.y - POLYGON_RADIUS), V(hi.x + POLYGON_RADIUS, hi.y + POLYGON_RADIUS)};
  return r;
}
static aabb_t fatten(aabb_t a) {
  aabb_t r = {V(a.lo.x - AABB_EXTENSION, a.lo.y - AABB_EXTENSION), V(a.hi.x + AABB_EXTENSION, a.hi.y + AABB_EXTENSION)};
  return r;
}
static int aabb_contains(aabb_t a, aabb_t b) {
  return a.lo.x <= b.lo.x && a.lo.y <= b.lo.y && b.hi.x <= a.hi.x && b.hi.y <= a.hi.y;
}
static int aabb_overlap(aabb_t a, aabb_t b) {
  v2 d1 = vsub(b.lo, a.hi), d2 = vsub(a.lo, b.hi);
  if (d1.x > 0.0f || d1.y > 0.0f) return 0;
  if (d2.x > 0.0f || d2.y > 0.0f) return 0;
  return 1;
}

/* --- b2CollideEdgeAndPolygon (b2EPCollider, plain edge without ghost vertices) ---------------------------------- */
typedef struct { v2 v; cid_t id; } clipv_t;
static int clip_segment(clipv_t out[2], const clipv_t in[2], v2 normal, float offset, int vertexIndexA) {
  int n = 0;
  float d0 = vdot(normal, in[0].v) - offset, d1 = vdot(normal, in[1].v) - offset;
  if (d0 <= 0.0f) out[n++] = in[0];
  if (d1 <= 0.0f) out[n++] = in[1];
  if (d0 * d1 < 0.0f) {
    float interp = d0 / (d0 - d1);
    out[n].v = vadd(in[0].v, vscale(interp, vsub(in[1].v, in[0].v)));
    out[n].id.indexA = (uint8_t)vertexIndexA;
    out[n].id.indexB = in[0].id.indexB;
    out[n].id.typeA = 0; /* e_vertex */
    out[n].id.typeB = 1; /* e_face */
    ++n;
  }
  return n;
}
static void collide_edge_polygon(manifold_t* mf, v2 ev1, v2 ev2, const poly_t* pb, xf_t xfB) {
  /* xfA is the identity (the moon sits at the origin): m_xf = b2MulT(xfA, xfB) = xfB */
  const xf_t xf = xfB;
  const v2 centroidB = xmul(xf, pb->centroid);
  v2 edge1 = vsub(ev2, ev1);
  vnormalize(&edge1);
  const v2 normal1 = V(edge1.y, -edge1.x);
  const float offset1 = vdot(normal1, vsub(centroidB, ev1));
  const int front = offset1 >= 0.0f;
  v2 normal, lower, upper;
  if (front) { normal = normal1; lower = vneg(normal1); upper = vneg(normal1); }
  else { normal = vneg(normal1); lower = normal1; upper = normal1; }
  v2 pv[MAXV], pn[MAXV];
  const int count = pb->count;
  for (int i = 0; i < count; ++i) { pv[i] = xmul(xf, pb->v[i]); pn[i] = rmul(xf.q, pb->n[i]); }
  const float radius = 2.0f * POLYGON_RADIUS;
  mf->count = 0;
  /* ComputeEdgeSeparation */
  float edge_sep = 3.402823466e+38f;
  for (int i = 0; i < count; ++i) { float s = vdot(normal, vsub(pv[i], ev1)); if (s < edge_sep) edge_sep = s; }
  if (edge_sep > radius) return;
  /* ComputePolygonSeparation */
  int ptype = 0 /*unknown*/, pindex = -1;
  float psep = -3.402823466e+38f;
  const v2 perp = V(-normal.y, normal.x);
  for (int i = 0; i < count; ++i) {
    v2 n = vneg(pn[i]);
    float s1 = vdot(n, vsub(pv[i], ev1)), s2 = vdot(n, vsub(pv[i], ev2)), s = s1 < s2 ? s1 : s2;
    if (s > radius) { ptype = 2; pindex = i; psep = s; break; }
    if (vdot(n, perp) >= 0.0f) { if (vdot(vsub(n, upper), normal) < -ANGULAR_SLOP) continue; }
    else { if (vdot(vsub(n, lower), normal) < -ANGULAR_SLOP) continue; }
    if (s > psep) { ptype = 2; pindex = i; psep = s; }
  }
  if (ptype != 0 && psep > radius) return;
  const float k_rel = 0.98f, k_abs = 0.001f;
  int primary_edge; /* 1: edgeA axis, 0: polygon axis */
  if (ptype == 0) primary_edge = 1;
  else if (psep > k_rel * edge_sep + k_abs) primary_edge = 0;
  else primary_edge = 1;
  clipv_t ie[2];
  int rf_i1, rf_i2;
  v2 rf_v1, rf_v2, rf_normal;
  if (primary_edge) {
    mf->type = 1;
    int best = 0;
    float bestv = vdot(normal, pn[0]);
    for (int i = 1; i < count; ++i) { float v = vdot(normal, pn[i]); if (v < bestv) { bestv = v; best = i; } }
    int i1 = best, i2 = i1 + 1 < count ? i1 + 1 : 0;
    ie[0].v = pv[i1]; ie[0].id.indexA = 0; ie[0].id.indexB = (uint8_t)i1; ie[0].id.typeA = 1; ie[0].id.typeB = 0;
    ie[1].v = pv[i2]; ie[1].id.indexA = 0; ie[1].id.indexB = (uint8_t)i2; ie[1].id.typeA = 1; ie[1].id.typeB = 0;
    if (front) { rf_i1 = 0; rf_i2 = 1; rf_v1 = ev1; rf_v2 = ev2; rf_normal = normal1; }
    else { rf_i1 = 1; rf_i2 = 0; rf_v1 = ev2; rf_v2 = ev1; rf_normal = vneg(normal1); }
  } else {
    mf->type = 2;
    ie[0].v = ev1; ie[0].id.indexA = 0; ie[0].id.indexB = (uint8_t)pindex; ie[0].id.typeA = 0; ie[0].id.typeB = 1;
    ie[1].v = ev2; ie[1].id.indexA = 0; ie[1].id.indexB = (uint8_t)pindex; ie[1].id.typeA = 0; ie[1].id.typeB = 1;
    rf_i1 = pindex; rf_i2 = rf_i1 + 1 < count ? rf_i1 + 1 : 0;
    rf_v1 = pv[rf_i1]; rf_v2 = pv[rf_i2]; rf_normal = pn[rf_i1];
  }
  const v2 side1 = V(rf_normal.y, -rf_normal.x), side2 = vneg(side1);
  const float off1 = vdot(side1, rf_v1), off2 = vdot(side2, rf_v2);
  clipv_t c1[2], c2[2];
  if (clip_segment(c1, ie, side1, off1, rf_i1) < 2) return;
  if (clip_segment(c2, c1, side2, off2, rf_i2) < 2) return;
  if (primary_edge) { mf->local_normal = rf_normal; mf->local_point = rf_v1; }
  else { mf->local_normal = pb->n[rf_i1]; mf->local_point = pb->v[rf_i1]; }
  int pc = 0;
  for (int i = 0; i < 2; ++i) {
    float sep = vdot(rf_normal, vsub(c2[i].v, rf_v1));
    if (sep <= radius) {
      mpoint_t* cp = &mf->points[pc];
      if (primary_edge) { cp->local_point = xmulT(xf, c2[i].v); cp->id = c2[i].id; }
      else {
        cp->local_point = c2[i].v;
        cp->id.typeA = c2[i].id.typeB; cp->id.typeB = c2[i].id.typeA;
        cp->id.indexA = c2[i].id.indexB; cp->id.indexB = c2[i].id.indexA;
      }
      ++pc;
    }
  }
  mf->count = pc;
}

/* --- contact listener (lunar_lander.py:58-76) --------------------------------------------------------------------- */
static void begin_contact(lander_t* L, int dyn) {
  if (dyn == 0) L->game_over = 1;
  else L->leg_contact[dyn - 1] = 1;
}
static void end_contact(lander_t* L, int dyn) {
  if (dyn >= 1) L->leg_contact[dyn - 1] = 0;
}

/* order contact indices most-recent-first */
static int contacts_sorted(const lander_t* L, int* idx, int dyn_filter /* -1 = all */) {
  int n = 0;
  for (int k = 0; k < NPAIR; ++k)
    if (L->contact[k].exists && (dyn_filter < 0 || k / NEDGE == dyn_filter)) idx[n++] = k;
  for (int i = 1; i < n; ++i) { /* insertion sort, descending seq */
    int k = idx[i], j = i - 1;
    while (j >= 0 && L->contact[idx[j]].seq < L->contact[k].seq) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = k;
  }
  return n;
}

/* b2Contact::Update: re-enable, evaluate the manifold at the bodies' current transforms, carry the impulses of matching
 * feature ids, wake on a change of the touching state, Begin/EndContact */
static void contact_update(lander_t* L, int k) {
  contact_t* c = &L->contact[k];
  int dyn = k / NEDGE, e = k % NEDGE;
  body_t* bB = &L->b[1 + dyn];
  manifold_t old = c->m;
  c->enabled = 1;
  int was = c->touching;
  collide_edge_polygon(&c->m, L->edge_v1[e], L->edge_v2[e], &L->poly[dyn], bB->xf);
  int touching = c->m.count > 0;
  for (int i = 0; i < c->m.count; ++i) {
    mpoint_t* mp2 = &c->m.points[i];
    mp2->normal_impulse = 0.0f; mp2->tangent_impulse = 0.0f;
    for (int j = 0; j < old.count; ++j)
      if (cid_key(old.points[j].id) == cid_key(mp2->id)) {
        mp2->normal_impulse = old.points[j].normal_impulse;
        mp2->tangent_impulse = old.points[j].tangent_impulse;
        break;
      }
  }
  if (touching != was) set_awake(bB, 1);
  c->touching = touching;
  if (!was && touching) begin_contact(L, dyn);
  if (was && !touching) end_contact(L, dyn);
}

/* b2ContactManager::Collide */
static void collide(lander_t* L) {
  int idx[NPAIR];
  int n = contacts_sorted(L, idx, -1);
  for (int t = 0; t < n; ++t) {
    contact_t* c = &L->contact[idx[t]];
    int dyn = idx[t] / NEDGE, e = idx[t] % NEDGE;
    body_t* bB = &L->b[1 + dyn];
    if (!bB->awake) continue; /* activeA (static) false, activeB = awake */
    if (!aabb_overlap(L->edge_fat[e], L->poly_fat[dyn])) {
      if (c->touching) end_contact(L, dyn);
      c->exists = 0;
      continue;
    }
    contact_update(L, idx[t]);
  }
}

/* b2BroadPhase::UpdatePairs for the moved dynamic proxies -> b2ContactManager::AddPair (sorted by proxy ids) */
static void find_new_contacts(lander_t* L, const int moved[NDYN]) {
  for (int e = 0; e < NEDGE; ++e)       /* pairs sort by (edge proxy id, polygon proxy id) */
    for (int d = 0; d < NDYN; ++d) {
      if (!moved[d]) continue;
      contact_t* c = &L->contact[d * NEDGE + e];
      if (c->exists) continue;
      if (!aabb_overlap(L->edge_fat[e], L->poly_fat[d])) continue;
      memset(c, 0, sizeof(*c));
      c->exists = 1;
      c->enabled = 1;      /* b2Contact constructor: m_flags = e_enabledFlag, m_toiCount = 0 */
      c->seq = ++L->seq;
      c->friction = sqrtf(L->edge_friction[e] * L->poly_friction[d]); /* b2MixFriction */
      set_awake(&L->b[1 + d], 1);
    }
}

/* --- b2RevoluteJoint -------------------------------------------------------------------------------------------- */
typedef struct { v2 c; float a; } pos_t;
typedef struct { v2 v; float w; } velo_t;

static void solve33(float m[3][3], float bx, float by, float bz, float* x, float* y, float* z) {
  /* b2Mat33::Solve33: columns ex=m[0], ey=m[1], ez=m[2] */
  float exx = m[0][0], exy = m[0][1], exz = m[0][2], eyx = m[1][0], eyy = m[1][1], eyz = m[1][2];
  float ezx = m[2][0], ezy = m[2][1], ezz = m[2][2];
  float cx = eyy * ezz - eyz * ezy, cy = eyz * ezx - eyx * ezz, cz = eyx * ezy - eyy * ezx; /* cross(ey, ez) */
  float det = exx * cx + exy * cy + exz * cz;
  if (det != 0.0f) det = 1.0f / det;
  *x = det * (bx * cx + by * cy + bz * cz);
  float dx = by * ezz - bz * ezy, dy = bz * ezx - bx * ezz, dz = bx * ezy - by * ezx; /* cross(b, ez) */
  *y = det * (exx * dx + exy * dy + exz * dz);
  float fx = eyy * bz - eyz * by, fy = eyz * bx - eyx * bz, fz = eyx * by - eyy * bx; /* cross(ey, b) */
  *z = det * (exx * fx + exy * fy + exz * fz);
}
static v2 solve22(float m[3][3], v2 b) {
  float a11 = m[0][0], a12 = m[1][0], a21 = m[0][1], a22 = m[1][1];
  float det = a11 * a22 - a12 * a21;
  if (det != 0.0f) det = 1.0f / det;
  return V(det * (a22 * b.x - a12 * b.y), det * (a11 * b.y - a21 * b.x));
}

static void joint_init_velocity(joint_t* j, const lander_t* L, pos_t* P, velo_t* Vv, int ia, int ib, float dt_ratio) {
  const body_t *A = &L->b[1], *B = &L->b[j->bodyB];
  j->lcA = A->local_center; j->lcB = B->local_center;
  j->mA = A->inv_mass; j->mB = B->inv_mass; j->iA = A->inv_I; j->iB = B->inv_I;
  float aA = P[ia].a, aB = P[ib].a;
  v2 vA = Vv[ia].v, vB = Vv[ib].v;
  float wA = Vv[ia].w, wB = Vv[ib].w;
  rot_t qA = rot_set(aA), qB = rot_set(aB);
  j->rA = rmul(qA, vsub(j->local_anchor_a, j->lcA));
  j->rB = rmul(qB, vsub(j->local_anchor_b, j->lcB));
  float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
  j->m[0][0] = mA + mB + j->rA.y * j->rA.y * iA + j->rB.y * j->rB.y * iB;
  j->m[1][0] = -j->rA.y * j->rA.x * iA - j->rB.y * j->rB.x * iB;
  j->m[2][0] = -j->rA.y * iA - j->rB.y * iB;
  j->m[0][1] = j->m[1][0];
  j->m[1][1] = mA + mB + j->rA.x * j->rA.x * iA + j->rB.x * j->rB.x * iB;
  j->m[2][1] = j->rA.x * iA + j->rB.x * iB;
  j->m[0][2] = j->m[2][0];
  j->m[1][2] = j->m[2][1];
  j->m[2][2] = iA + iB;
  j->motor_mass = iA + iB;
  if (j->motor_mass > 0.0f) j->motor_mass = 1.0f / j->motor_mass;
  /* enableMotor = enableLimit = true, fixedRotation = false */
  float angle = aB - aA - j->reference_angle;
  if (fabsf(j->upper - j->lower) < 2.0f * ANGULAR_SLOP) j->limit_state = 3;
  else if (angle <= j->lower) { if (j->limit_state != 1) j->imp_z = 0.0f; j->limit_state = 1; }
  else if (angle >= j->upper) { if (j->limit_state != 2) j->imp_z = 0.0f; j->limit_state = 2; }
  else { j->limit_state = 0; j->imp_z = 0.0f; }
  /* warm starting */
  j->imp_x *= dt_ratio; j->imp_y *= dt_ratio; j->imp_z *= dt_ratio; j->motor_impulse *= dt_ratio;
  v2 Pi = V(j->imp_x, j->imp_y);
  vA = vsub(vA, vscale(mA, Pi));
  wA -= iA * (vcross(j->rA, Pi) + j->motor_impulse + j->imp_z);
  vB = vadd(vB, vscale(mB, Pi));
  wB += iB * (vcross(j->rB, Pi) + j->motor_impulse + j->imp_z);
  Vv[ia].v = vA; Vv[ia].w = wA; Vv[ib].v = vB; Vv[ib].w = wB;
}

static void joint_solve_velocity(joint_t* j, velo_t* Vv, int ia, int ib, float dt) {
  v2 vA = Vv[ia].v, vB = Vv[ib].v;
  float wA = Vv[ia].w, wB = Vv[ib].w;
  float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
  if (j->limit_state != 3) { /* motor */
    float Cdot = wB - wA - j->motor_speed;
    float impulse = -j->motor_mass * Cdot;
    float old = j->motor_impulse, maxi = dt * j->max_motor_torque;
    j->motor_impulse = fclamp(old + impulse, -maxi, maxi);
    impulse = j->motor_impulse - old;
    wA -= iA * impulse;
    wB += iB * impulse;
  }
  if (j->limit_state != 0) {
    v2 Cdot1 = vsub(vsub(vadd(vB, vcross_sv(wB, j->rB)), vA), vcross_sv(wA, j->rA));
    float Cdot2 = wB - wA;
    float ix, iy, iz;
    solve33(j->m, Cdot1.x, Cdot1.y, Cdot2, &ix, &iy, &iz);
    ix = -ix; iy = -iy; iz = -iz;
    if (j->limit_state == 3) { j->imp_x += ix; j->imp_y += iy; j->imp_z += iz; }
    else if (j->limit_state == 1) {
      float newi = j->imp_z + iz;
      if (newi < 0.0f) {
        v2 rhs = vadd(vneg(Cdot1), vscale(j->imp_z, V(j->m[2][0], j->m[2][1])));
        v2 red = solve22(j->m, rhs);
        ix = red.x; iy = red.y; iz = -j->imp_z;
        j->imp_x += red.x; j->imp_y += red.y; j->imp_z = 0.0f;
      } else { j->imp_x += ix; j->imp_y += iy; j->imp_z += iz; }
    } else {
      float newi = j->imp_z + iz;
      if (newi > 0.0f) {
        v2 rhs = vadd(vneg(Cdot1), vscale(j->imp_z, V(j->m[2][0], j->m[2][1])));
        v2 red = solve22(j->m, rhs);
        ix = red.x; iy = red.y; iz = -j->imp_z;
        j->imp_x += red.x; j->imp_y += red.y; j->imp_z = 0.0f;
      } else { j->imp_x += ix; j->imp_y += iy; j->imp_z += iz; }
    }
    v2 Pi = V(ix, iy);
    vA = vsub(vA, vscale(mA, Pi));
    wA -= iA * (vcross(j->rA, Pi) + iz);
    vB = vadd(vB, vscale(mB, Pi));
    wB += iB * (vcross(j->rB, Pi) + iz);
  } else {
    v2 Cdot = vsub(vsub(vadd(vB, vcross_sv(wB, j->rB)), vA), vcross_sv(wA, j->rA));
    v2 imp = solve22(j->m, vneg(Cdot));
    j->imp_x += imp.x; j->imp_y += imp.y;
    vA = vsub(vA, vscale(mA, imp));
    wA -= iA * vcross(j->rA, imp);
    vB = vadd(vB, vscale(mB, imp));
    wB += iB * vcross(j->rB, imp);
  }
  Vv[ia].v = vA; Vv[ia].w = wA; Vv[ib].v = vB; Vv[ib].w = wB;
}

static int joint_solve_position(joint_t* j, pos_t* P, int ia, int ib) {
  v2 cA = P[ia].c, cB = P[ib].c;
  float aA = P[ia].a, aB = P[ib].a;
  float angular_error = 0.0f, position_error;
  if (j->limit_state != 0) {
    float angle = aB - aA - j->reference_angle, limit_impulse = 0.0f;
    if (j->limit_state == 3) {
      float C = fclamp(angle - j->lower, -MAX_ANGULAR_CORRECTION, MAX_ANGULAR_CORRECTION);
      limit_impulse = -j->motor_mass * C; angular_error = fabsf(C);
    } else if (j->limit_state == 1) {
      float C = angle - j->lower;
      angular_error = -C;
      C = fclamp(C + ANGULAR_SLOP, -MAX_ANGULAR_CORRECTION, 0.0f);
      limit_impulse = -j->motor_mass * C;
    } else {
      float C = angle - j->upper;
      angular_error = C;
      C = fclamp(C - ANGULAR_SLOP, 0.0f, MAX_ANGULAR_CORRECTION);
      limit_impulse = -j->motor_mass * C;
    }
    aA -= j->iA * limit_impulse;
    aB += j->iB * limit_impulse;
  }
  {
    rot_t qA = rot_set(aA), qB = rot_set(aB);
    v2 rA = rmul(qA, vsub(j->local_anchor_a, j->lcA)), rB = rmul(qB, vsub(j->local_anchor_b, j->lcB));
    v2 C = vsub(vsub(vadd(cB, rB), cA), rA);
    position_error = vlen(C);
    float mA = j->mA, mB = j->mB, iA = j->iA, iB = j->iB;
    float k11 = mA + mB + iA * rA.y * rA.y + iB * rB.y * rB.y;
    float k12 = -iA * rA.x * rA.y - iB * rB.x * rB.y;
    float k22 = mA + mB + iA * rA.x * rA.x + iB * rB.x * rB.x;
    float det = k11 * k22 - k12 * k12; /* b2Mat22::Solve: a11*a22 - a12*a21 */
    if (det != 0.0f) det = 1.0f / det;
    v2 imp = V(-(det * (k22 * C.x - k12 * C.y)), -(det * (k11 * C.y - k12 * C.x)));
    cA = vsub(cA, vscale(mA, imp));
    aA -= iA * vcross(rA, imp);
    cB = vadd(cB, vscale(mB, imp));
    aB += iB * vcross(rB, imp);
  }
  P[ia].c = cA; P[ia].a = aA; P[ib].c = cB; P[ib].a = aB;
  return position_error <= LINEAR_SLOP && angular_error <= ANGULAR_SLOP;
}

/* --- b2ContactSolver ------------------------------------------------------------------------------------------- */
typedef struct {
  v2 rA, rB;
  float normal_impulse, tangent_impulse, normal_mass, tangent_mass, velocity_bias;
} vcp_t;
typedef struct {
  contact_t* c;
  int indexB, count, pos_count, type;
  vcp_t p[2];
  v2 normal, local_normal, local_point, local_points[2], lcB;
  float K[2][2], NM[2][2]; /* columns: K[col][row] */
  float friction, mB, iB;
} vc_t;

static void world_manifold(const manifold_t* m, xf_t xfB, v2* normal, v2 pts[2]) {
  /* xfA = identity; radiusA = radiusB = b2_polygonRadius */
  if (m->type == 1) {
    *normal = m->local_normal; /* b2Mul(identity.q, n) */
    v2 plane = m->local_point;
    for (int i = 0; i < m->count; ++i) {
      v2 clip = xmul(xfB, m->points[i].local_point);
      v2 cA = vadd(clip, vscale(POLYGON_RADIUS - vdot(vsub(clip, plane), *normal), *normal));
      v2 cB = vsub(clip, vscale(POLYGON_RADIUS, *normal));
      pts[i] = vscale(0.5f, vadd(cA, cB));
    }
  } else {
    *normal = rmul(xfB.q, m->local_normal);
    v2 plane = xmul(xfB, m->local_point);
    for (int i = 0; i < m->count; ++i) {
      v2 clip = m->points[i].local_point; /* b2Mul(identity, p) */
      v2 cB = vadd(clip, vscale(POLYGON_RADIUS - vdot(vsub(clip, plane), *normal), *normal));
      v2 cA = vsub(clip, vscale(POLYGON_RADIUS, *normal));
      pts[i] = vscale(0.5f, vadd(cA, cB));
    }
    *normal = vneg(*normal);
  }
}

/* ------------------------------------------------------------------------------------------------------------------
 * b2World::Solve for the single island {lander, legs} (+ moon as a static participant) */
static void solve_island(lander_t* L, float h, float dt_ratio, int vel_iters, int pos_iters) {
  /* island body order follows the DFS from the most recently created body: legs[1] (body 3), lander, legs[0];
   * the moon (static) is appended when a touching contact reaches it -- it has zero inverse mass so only the index
   * differs.  Local indices here: 0 = moon, 1..3 = bodies 1..3. */
  pos_t P[NBODY];
  velo_t Vv[NBODY];
  if (!L->b[1].awake && !L->b[2].awake && !L->b[3].awake) return;
  for (int i = 1; i < NBODY; ++i) set_awake(&L->b[i], 1); /* "Make sure the body is awake" during island build */
  P[0].c = V(0, 0); P[0].a = 0; Vv[0].v = V(0, 0); Vv[0].w = 0;
  const v2 gravity = V(0.0f, L->gravity);
  for (int i = 1; i < NBODY; ++i) {
    body_t* b = &L->b[i];
    v2 v = b->vel;
    float w = b->w;
    b->c0 = b->c; b->a0 = b->a;
    v = vadd(v, vscale(h, vadd(vscale(1.0f, gravity), vscale(b->inv_mass, b->force))));
    w += h * b->inv_I * b->torque;
    v = vscale(1.0f / (1.0f + h * 0.0f), v);
    w *= 1.0f / (1.0f + h * 0.0f);
    P[i].c = b->c; P[i].a = b->a; Vv[i].v = v; Vv[i].w = w;
  }
  /* island contact order: contacts of legs[1], then lander, then legs[0] (each list most-recent-first); only touching */
  vc_t vcs[NPAIR];
  int nvc = 0;
  const int dyn_order[3] = {2, 0, 1};
  for (int t = 0; t < 3; ++t) {
    int idx[NPAIR], n = contacts_sorted(L, idx, dyn_order[t]);
    for (int k = 0; k < n; ++k) {
      contact_t* c = &L->contact[idx[k]];
      if (!c->touching) continue;
      vc_t* vc = &vcs[nvc++];
      memset(vc, 0, sizeof(*vc));
      vc->c = c;
      vc->indexB = 1 + dyn_order[t];
      vc->friction = c->friction;
      vc->count = vc->pos_count = c->m.count;
      vc->type = c->m.type;
      const body_t* B = &L->b[vc->indexB];
      vc->mB = B->inv_mass; vc->iB = B->inv_I; vc->lcB = B->local_center;
      vc->local_normal = c->m.local_normal; vc->local_point = c->m.local_point;
      for (int j = 0; j < vc->count; ++j) {
        vc->p[j].normal_impulse = dt_ratio * c->m.points[j].normal_impulse;
        vc->p[j].tangent_impulse = dt_ratio * c->m.points[j].tangent_impulse;
        vc->local_points[j] = c->m.points[j].local_point;
      }
    }
  }
  /* InitializeVelocityConstraints */
  for (int k = 0; k < nvc; ++k) {
    vc_t* vc = &vcs[k];
    int ib = vc->indexB;
    float mB = vc->mB, iB = vc->iB;
    v2 cB = P[ib].c, vB = Vv[ib].v;
    float aB = P[ib].a, wB = Vv[ib].w;
    xf_t xfB;
    xfB.q = rot_set(aB);
    xfB.p = vsub(cB, rmul(xfB.q, vc->lcB));
    v2 pts[2];
    world_manifold(&vc->c->m, xfB, &vc->normal, pts);
    for (int j = 0; j < vc->count; ++j) {
      vcp_t* p = &vc->p[j];
      p->rA = pts[j];               /* cA = 0 */
      p->rB = vsub(pts[j], cB);
      float rnB = vcross(p->rB, vc->normal);
      float kN = mB + iB * rnB * rnB;   /* mA = iA = 0: 0 + mB + 0*rnA*rnA + iB*rnB*rnB */
      p->normal_mass = kN > 0.0f ? 1.0f / kN : 0.0f;
      v2 tangent = vcross_vs(vc->normal, 1.0f);
      float rtB = vcross(p->rB, tangent);
      float kT = mB + iB * rtB * rtB;
      p->tangent_mass = kT > 0.0f ? 1.0f / kT : 0.0f;
      p->velocity_bias = 0.0f; /* restitution 0: -0 * vRel */
      (void)vB; (void)wB;
    }
    if (vc->count == 2) {
      float rn1B = vcross(vc->p[0].rB, vc->normal), rn2B = vcross(vc->p[1].rB, vc->normal);
      float k11 = mB + iB * rn1B * rn1B, k22 = mB + iB * rn2B * rn2B, k12 = mB + iB * rn1B * rn2B;
      if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
        vc->K[0][0] = k11; vc->K[0][1] = k12; vc->K[1][0] = k12; vc->K[1][1] = k22;
        float a = k11, b = k12, c = k12, d = k22, det = a * d - b * c;
        if (det != 0.0f) det = 1.0f / det;
        vc->NM[0][0] = det * d; vc->NM[1][0] = -det * b; vc->NM[0][1] = -det * c; vc->NM[1][1] = det * a;
      } else vc->count = 1;
    }
  }
  /* WarmStart */
  for (int k = 0; k < nvc; ++k) {
    vc_t* vc = &vcs[k];
    int ib = vc->indexB;
    v2 vB = Vv[ib].v;
    float wB = Vv[ib].w;
    v2 tangent = vcross_vs(vc->normal, 1.0f);
    for (int j = 0; j < vc->count; ++j) {
      v2 Pi = vadd(vscale(vc->p[j].normal_impulse, vc->normal), vscale(vc->p[j].tangent_impulse, tangent));
      wB += vc->iB * vcross(vc->p[j].rB, Pi);
      vB = vadd(vB, vscale(vc->mB, Pi));
    }
    Vv[ib].v = vB; Vv[ib].w = wB;
  }
  /* joints: island order = joint of legs[1] first, then legs[0] */
  joint_init_velocity(&L->joint[1], L, P, Vv, 1, 3, dt_ratio);
  joint_init_velocity(&L->joint[0], L, P, Vv, 1, 2, dt_ratio);
  for (int it = 0; it < vel_iters; ++it) {
    joint_solve_velocity(&L->joint[1], Vv, 1, 3, h);
    joint_solve_velocity(&L->joint[0], Vv, 1, 2, h);
    for (int k = 0; k < nvc; ++k) {
      vc_t* vc = &vcs[k];
      int ib = vc->indexB;
      float mB = vc->mB, iB = vc->iB;
      v2 vB = Vv[ib].v;
      float wB = Vv[ib].w;
      v2 normal = vc->normal, tangent = vcross_vs(normal, 1.0f);
      for (int j = 0; j < vc->count; ++j) { /* friction first */
        vcp_t* p = &vc->p[j];
        v2 dv = vadd(vB, vcross_sv(wB, p->rB)); /* - vA - cross(wA, rA) with vA = wA = 0 */
        float vt = vdot(dv, tangent) - 0.0f;
        float lambda = p->tangent_mass * (-vt);
        float maxf = vc->friction * p->normal_impulse;
        float newi = fclamp(p->tangent_impulse + lambda, -maxf, maxf);
        lambda = newi - p->tangent_impulse;
        p->tangent_impulse = newi;
        v2 Pi = vscale(lambda, tangent);
        vB = vadd(vB, vscale(mB, Pi));
        wB += iB * vcross(p->rB, Pi);
      }
      if (vc->count == 1) {
        vcp_t* p = &vc->p[0];
        v2 dv = vadd(vB, vcross_sv(wB, p->rB));
        float vn = vdot(dv, normal);
        float lambda = -p->normal_mass * (vn - p->velocity_bias);
        float newi = fmaxf(p->normal_impulse + lambda, 0.0f);
        lambda = newi - p->normal_impulse;
        p->normal_impulse = newi;
        v2 Pi = vscale(lambda, normal);
        vB = vadd(vB, vscale(mB, Pi));
        wB += iB * vcross(p->rB, Pi);
      } else {
        vcp_t *c1 = &vc->p[0], *c2 = &vc->p[1];
        float ax = c1->normal_impulse, ay = c2->normal_impulse;
        v2 dv1 = vadd(vB, vcross_sv(wB, c1->rB)), dv2 = vadd(vB, vcross_sv(wB, c2->rB));
        float vn1 = vdot(dv1, normal), vn2 = vdot(dv2, normal);
        float bx = vn1 - c1->velocity_bias, by = vn2 - c2->velocity_bias;
        /* b -= K * a */
        bx -= vc->K[0][0] * ax + vc->K[1][0] * ay;
        by -= vc->K[0][1] * ax + vc->K[1][1] * ay;
        float xx, xy;
        int solved = 0;
        /* case 1 */
        xx = -(vc->NM[0][0] * bx + vc->NM[1][0] * by);
        xy = -(vc->NM[0][1] * bx + vc->NM[1][1] * by);
        if (xx >= 0.0f && xy >= 0.0f) solved = 1;
        if (!solved) { /* case 2 */
          xx = -c1->normal_mass * bx; xy = 0.0f;
          vn2 = vc->K[0][1] * xx + by;
          if (xx >= 0.0f && vn2 >= 0.0f) solved = 1;
        }
        if (!solved) { /* case 3 */
          xx = 0.0f; xy = -c2->normal_mass * by;
          vn1 = vc->K[1][0] * xy + bx;
          if (xy >= 0.0f && vn1 >= 0.0f) solved = 1;
        }
        if (!solved) { /* case 4 */
          xx = 0.0f; xy = 0.0f;
          if (bx >= 0.0f && by >= 0.0f) solved = 1;
        }
        if (solved) {
          float dx = xx - ax, dy = xy - ay;
          v2 P1 = vscale(dx, normal), P2 = vscale(dy, normal);
          vB = vadd(vB, vscale(mB, vadd(P1, P2)));
          wB += iB * (vcross(c1->rB, P1) + vcross(c2->rB, P2));
          c1->normal_impulse = xx; c2->normal_impulse = xy;
        }
      }
      Vv[ib].v = vB; Vv[ib].w = wB;
    }
  }
  /* StoreImpulses */
  for (int k = 0; k < nvc; ++k)
    for (int j = 0; j < vcs[k].count; ++j) {
      vcs[k].c->m.points[j].normal_impulse = vcs[k].p[j].normal_impulse;
      vcs[k].c->m.points[j].tangent_impulse = vcs[k].p[j].tangent_impulse;
    }
  /* integrate positions */
  for (int i = 1; i < NBODY; ++i) {
    v2 c = P[i].c, v = Vv[i].v;
    float a = P[i].a, w = Vv[i].w;
    v2 tr = vscale(h, v);
    if (vdot(tr, tr) > MAX_TRANSLATION * MAX_TRANSLATION) { float ratio = MAX_TRANSLATION / vlen(tr); v = vscale(ratio, v); }
    float rotn = h * w;
    if (rotn * rotn > MAX_ROTATION * MAX_ROTATION) { float ratio = MAX_ROTATION / fabsf(rotn); w *= ratio; }
    c = vadd(c, vscale(h, v));
    a += h * w;
    P[i].c = c; P[i].a = a; Vv[i].v = v; Vv[i].w = w;
  }
  /* position iterations */
  int position_solved = 0;
  for (int it = 0; it < pos_iters; ++it) {
    float min_sep = 0.0f;
    for (int k = 0; k < nvc; ++k) {
      vc_t* vc = &vcs[k];
      int ib = vc->indexB;
      float mB = vc->mB, iB = vc->iB;
      v2 cB = P[ib].c;
      float aB = P[ib].a;
      for (int j = 0; j < vc->pos_count; ++j) {
        xf_t xfB;
        xfB.q = rot_set(aB);
        xfB.p = vsub(cB, rmul(xfB.q, vc->lcB));
        v2 normal, point;
        float sep;
        if (vc->type == 1) {
          normal = vc->local_normal;
          v2 plane = vc->local_point, clip = xmul(xfB, vc->local_points[j]);
          sep = vdot(vsub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
          point = clip;
        } else {
          normal = rmul(xfB.q, vc->local_normal);
          v2 plane = xmul(xfB, vc->local_point), clip = vc->local_points[j];
          sep = vdot(vsub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
          point = clip;
          normal = vneg(normal);
        }
        v2 rB = vsub(point, cB);
        min_sep = fminf(min_sep, sep);
        float C = fclamp(BAUMGARTE * (sep + LINEAR_SLOP), -MAX_LINEAR_CORRECTION, 0.0f);
        float rnB = vcross(rB, normal);
        float K = mB + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        v2 Pi = vscale(impulse, normal);
        cB = vadd(cB, vscale(mB, Pi));
        aB += iB * vcross(rB, Pi);
      }
      P[ib].c = cB; P[ib].a = aB;
    }
    int contacts_ok = min_sep >= -3.0f * LINEAR_SLOP;
    int j1 = joint_solve_position(&L->joint[1], P, 1, 3);
    int j0 = joint_solve_position(&L->joint[0], P, 1, 2);
    if (contacts_ok && j1 && j0) { position_solved = 1; break; }
  }
  for (int i = 1; i < NBODY; ++i) {
    body_t* b = &L->b[i];
    b->c = P[i].c; b->a = P[i].a; b->vel = Vv[i].v; b->w = Vv[i].w;
    sync_transform(b);
  }
  /* sleep */
  float min_sleep = 3.402823466e+38f;
  const float lin2 = LINEAR_SLEEP_TOL * LINEAR_SLEEP_TOL, ang2 = ANGULAR_SLEEP_TOL * ANGULAR_SLEEP_TOL;
  for (int i = 1; i < NBODY; ++i) {
    body_t* b = &L->b[i];
    if (b->w * b->w > ang2 || vdot(b->vel, b->vel) > lin2) { b->sleep_time = 0.0f; min_sleep = 0.0f; }
    else { b->sleep_time += h; min_sleep = fminf(min_sleep, b->sleep_time); }
  }
  if (min_sleep >= TIME_TO_SLEEP && position_solved)
    for (int i = 1; i < NBODY; ++i) set_awake(&L->b[i], 0);
}

/* ------------------------------------------------------------------------------------------------------------------
 * Continuous collision: b2Distance (GJK), b2TimeOfImpact, b2World::SolveTOI / b2Island::SolveTOI.
 * Proxy A is always a terrain edge (2 vertices, body = the static moon whose transform is the identity: b2Mul / b2MulT by
 * it are exact no-ops and are written as such), proxy B a polygon of a dynamic body.  Radii: b2_polygonRadius each. */
#define B2_EPSILON 1.192092896e-07f
#define B2_MAXFLOAT 3.402823466e+38f
#define MAX_SUB_STEPS 8
#define TOI_BAUMGARTE 0.75f
#define MAX_POLY_VERTS 8

typedef struct { v2 lc, c0, c; float a0, a, alpha0; } sweep_t;
static xf_t sweep_xf(const sweep_t* s, float beta) { /* b2Sweep::GetTransform */
  xf_t xf;
  xf.p = vadd(vscale(1.0f - beta, s->c0), vscale(beta, s->c));
  float angle = (1.0f - beta) * s->a0 + beta * s->a;
  xf.q = rot_set(angle);
  xf.p = vsub(xf.p, rmul(xf.q, s->lc));
  return xf;
}
static void sweep_advance(sweep_t* s, float alpha) { /* b2Sweep::Advance */
  float beta = (alpha - s->alpha0) / (1.0f - s->alpha0);
  s->c0 = vadd(s->c0, vscale(beta, vsub(s->c, s->c0)));
  s->a0 += beta * (s->a - s->a0);
  s->alpha0 = alpha;
}
static sweep_t body_sweep(const body_t* b) {
  sweep_t s = {b->local_center, b->c0, b->c, b->a0, b->a, b->alpha0};
  return s;
}
static void body_set_sweep(body_t* b, const sweep_t* s) { b->c0 = s->c0; b->c = s->c; b->a0 = s->a0; b->a = s->a; b->alpha0 = s->alpha0; }
static void body_advance(body_t* b, float alpha) { /* b2Body::Advance */
  sweep_t s = body_sweep(b);
  sweep_advance(&s, alpha);
  s.c = s.c0;
  s.a = s.a0;
  body_set_sweep(b, &s);
  sync_transform(b);
}

typedef struct { const v2* v; int count; } proxy_t;
static int proxy_support(const proxy_t* p, v2 d) { /* b2DistanceProxy::GetSupport */
  int best = 0;
  float bestv = vdot(p->v[0], d);
  for (int i = 1; i < p->count; ++i) {
    float v = vdot(p->v[i], d);
    if (v > bestv) { best = i; bestv = v; }
  }
  return best;
}

typedef struct { v2 wA, wB, w; float a; int iA, iB; } sv_t;        /* b2SimplexVertex */
typedef struct { float metric; int count; int iA[3], iB[3]; } scache_t; /* b2SimplexCache */
typedef struct { sv_t v[3]; int count; } simplex_t;

static float simplex_metric(const simplex_t* s) {
  switch (s->count) {
    case 1: return 0.0f;
    case 2: return vlen(vsub(s->v[0].w, s->v[1].w));
    case 3: return vcross(vsub(s->v[1].w, s->v[0].w), vsub(s->v[2].w, s->v[0].w));
    default: return 0.0f;
  }
}
static void simplex_solve2(simplex_t* s) {
  v2 w1 = s->v[0].w, w2 = s->v[1].w, e12 = vsub(w2, w1);
  float d12_2 = -vdot(w1, e12);
  if (d12_2 <= 0.0f) { s->v[0].a = 1.0f; s->count = 1; return; }
  float d12_1 = vdot(w2, e12);
  if (d12_1 <= 0.0f) { s->v[1].a = 1.0f; s->count = 1; s->v[0] = s->v[1]; return; }
  float inv = 1.0f / (d12_1 + d12_2);
  s->v[0].a = d12_1 * inv;
  s->v[1].a = d12_2 * inv;
  s->count = 2;
}
static void simplex_solve3(simplex_t* s) {
  v2 w1 = s->v[0].w, w2 = s->v[1].w, w3 = s->v[2].w;
  v2 e12 = vsub(w2, w1);
  float w1e12 = vdot(w1, e12), w2e12 = vdot(w2, e12), d12_1 = w2e12, d12_2 = -w1e12;
  v2 e13 = vsub(w3, w1);
  float w1e13 = vdot(w1, e13), w3e13 = vdot(w3, e13), d13_1 = w3e13, d13_2 = -w1e13;
  v2 e23 = vsub(w3, w2);
  float w2e23 = vdot(w2, e23), w3e23 = vdot(w3, e23), d23_1 = w3e23, d23_2 = -w2e23;
  float n123 = vcross(e12, e13);
  float d123_1 = n123 * vcross(w2, w3), d123_2 = n123 * vcross(w3, w1), d123_3 = n123 * vcross(w1, w2);
  if (d12_2 <= 0.0f && d13_2 <= 0.0f) { s->v[0].a = 1.0f; s->count = 1; return; }
  if (d12_1 > 0.0f && d12_2 > 0.0f && d123_3 <= 0.0f) {
    float inv = 1.0f / (d12_1 + d12_2);
    s->v[0].a = d12_1 * inv; s->v[1].a = d12_2 * inv; s->count = 2; return;
  }
  if (d13_1 > 0.0f && d13_2 > 0.0f && d123_2 <= 0.0f) {
    float inv = 1.0f / (d13_1 + d13_2);
    s->v[0].a = d13_1 * inv; s->v[2].a = d13_2 * inv; s->count = 2; s->v[1] = s->v[2]; return;
  }
  if (d12_1 <= 0.0f && d23_2 <= 0.0f) { s->v[1].a = 1.0f; s->count = 1; s->v[0] = s->v[1]; return; }
  if (d13_1 <= 0.0f && d23_1 <= 0.0f) { s->v[2].a = 1.0f; s->count = 1; s->v[0] = s->v[2]; return; }
  if (d23_1 > 0.0f && d23_2 > 0.0f && d123_1 <= 0.0f) {
    float inv = 1.0f / (d23_1 + d23_2);
    s->v[1].a = d23_1 * inv; s->v[2].a = d23_2 * inv; s->count = 2; s->v[0] = s->v[2]; return;
  }
  float inv = 1.0f / (d123_1 + d123_2 + d123_3);
  s->v[0].a = d123_1 * inv; s->v[1].a = d123_2 * inv; s->v[2].a = d123_3 * inv; s->count = 3;
}

/* b2Distance with useRadii = false; returns the distance, updates the cache */
static float gjk_distance(scache_t* cache, const proxy_t* pA, const proxy_t* pB, xf_t xfB) {
  simplex_t s;
  s.count = cache->count;
  for (int i = 0; i < s.count; ++i) { /* b2Simplex::ReadCache */
    sv_t* v = &s.v[i];
    v->iA = cache->iA[i]; v->iB = cache->iB[i];
    v->wA = pA->v[v->iA];
    v->wB = xmul(xfB, pB->v[v->iB]);
    v->w = vsub(v->wB, v->wA);
    v->a = 0.0f;
  }
  if (s.count > 1) {
    float metric1 = cache->metric, metric2 = simplex_metric(&s);
    if (metric2 < 0.5f * metric1 || 2.0f * metric1 < metric2 || metric2 < B2_EPSILON) s.count = 0;
  }
  if (s.count == 0) {
    sv_t* v = &s.v[0];
    v->iA = 0; v->iB = 0;
    v->wA = pA->v[0];
    v->wB = xmul(xfB, pB->v[0]);
    v->w = vsub(v->wB, v->wA);
    v->a = 1.0f;
    s.count = 1;
  }
  int saveA[3], saveB[3], iter = 0;
  while (iter < 20) {
    int save_count = s.count;
    for (int i = 0; i < save_count; ++i) { saveA[i] = s.v[i].iA; saveB[i] = s.v[i].iB; }
    if (s.count == 2) simplex_solve2(&s);
    else if (s.count == 3) simplex_solve3(&s);
    if (s.count == 3) break;
    v2 d; /* b2Simplex::GetSearchDirection */
    if (s.count == 1) d = vneg(s.v[0].w);
    else {
      v2 e12 = vsub(s.v[1].w, s.v[0].w);
      float sgn = vcross(e12, vneg(s.v[0].w));
      d = sgn > 0.0f ? vcross_sv(1.0f, e12) : vcross_vs(e12, 1.0f);
    }
    if (vdot(d, d) < B2_EPSILON * B2_EPSILON) break;
    sv_t* v = &s.v[s.count];
    v->iA = proxy_support(pA, vneg(d));               /* b2MulT(identity, -d) */
    v->wA = pA->v[v->iA];
    v->iB = proxy_support(pB, rmulT(xfB.q, d));
    v->wB = xmul(xfB, pB->v[v->iB]);
    v->w = vsub(v->wB, v->wA);
    ++iter;
    int duplicate = 0;
    for (int i = 0; i < save_count; ++i)
      if (v->iA == saveA[i] && v->iB == saveB[i]) { duplicate = 1; break; }
    if (duplicate) break;
    ++s.count;
  }
  v2 a, b; /* GetWitnessPoints */
  if (s.count == 1) { a = s.v[0].wA; b = s.v[0].wB; }
  else if (s.count == 2) {
    a = vadd(vscale(s.v[0].a, s.v[0].wA), vscale(s.v[1].a, s.v[1].wA));
    b = vadd(vscale(s.v[0].a, s.v[0].wB), vscale(s.v[1].a, s.v[1].wB));
  } else {
    a = vadd(vadd(vscale(s.v[0].a, s.v[0].wA), vscale(s.v[1].a, s.v[1].wA)), vscale(s.v[2].a, s.v[2].wA));
    b = a;
  }
  cache->metric = simplex_metric(&s); /* WriteCache */
  cache->count = s.count;
  for (int i = 0; i < s.count; ++i) { cache->iA[i] = s.v[i].iA; cache->iB[i] = s.v[i].iB; }
  return vlen(vsub(a, b));
}

/* b2SeparationFunction (sweep A is the identity for all t) */
typedef struct { const proxy_t *pA, *pB; sweep_t sB; int type /*0 points, 1 faceA, 2 faceB*/; v2 local_point, axis; } sepfn_t;
static void sep_init(sepfn_t* f, const scache_t* cache, const proxy_t* pA, const proxy_t* pB, const sweep_t* sB, float t1) {
  f->pA = pA; f->pB = pB; f->sB = *sB;
  xf_t xfB = sweep_xf(sB, t1);
  if (cache->count == 1) {
    f->type = 0;
    v2 pointA = pA->v[cache->iA[0]], pointB = xmul(xfB, pB->v[cache->iB[0]]);
    f->axis = vsub(pointB, pointA);
    vnormalize(&f->axis);
  } else if (cache->iA[0] == cache->iA[1]) {
    f->type = 2; /* two points on B, one on A */
    v2 b1 = pB->v[cache->iB[0]], b2 = pB->v[cache->iB[1]];
    f->axis = vcross_vs(vsub(b2, b1), 1.0f);
    vnormalize(&f->axis);
    v2 normal = rmul(xfB.q, f->axis);
    f->local_point = vscale(0.5f, vadd(b1, b2));
    v2 pointB = xmul(xfB, f->local_point), pointA = pA->v[cache->iA[0]];
    float s = vdot(vsub(pointA, pointB), normal);
    if (s < 0.0f) f->axis = vneg(f->axis);
  } else {
    f->type = 1; /* two points on A */
    v2 a1 = pA->v[cache->iA[0]], a2 = pA->v[cache->iA[1]];
    f->axis = vcross_vs(vsub(a2, a1), 1.0f);
    vnormalize(&f->axis);
    v2 normal = f->axis;
    f->local_point = vscale(0.5f, vadd(a1, a2));
    v2 pointA = f->local_point, pointB = xmul(xfB, pB->v[cache->iB[0]]);
    float s = vdot(vsub(pointB, pointA), normal);
    if (s < 0.0f) f->axis = vneg(f->axis);
  }
}
static float sep_find_min(const sepfn_t* f, int* iA, int* iB, float t) {
  xf_t xfB = sweep_xf(&f->sB, t);
  if (f->type == 0) {
    *iA = proxy_support(f->pA, f->axis);
    *iB = proxy_support(f->pB, rmulT(xfB.q, vneg(f->axis)));
    v2 pointA = f->pA->v[*iA], pointB = xmul(xfB, f->pB->v[*iB]);
    return vdot(vsub(pointB, pointA), f->axis);
  } else if (f->type == 1) {
    v2 normal = f->axis, pointA = f->local_point;
    *iA = -1;
    *iB = proxy_support(f->pB, rmulT(xfB.q, vneg(normal)));
    v2 pointB = xmul(xfB, f->pB->v[*iB]);
    return vdot(vsub(pointB, pointA), normal);
  } else {
    v2 normal = rmul(xfB.q, f->axis), pointB = xmul(xfB, f->local_point);
    *iB = -1;
    *iA = proxy_support(f->pA, vneg(normal));
    v2 pointA = f->pA->v[*iA];
    return vdot(vsub(pointA, pointB), normal);
  }
}
static float sep_evaluate(const sepfn_t* f, int iA, int iB, float t) {
  xf_t xfB = sweep_xf(&f->sB, t);
  if (f->type == 0) {
    v2 pointA = f->pA->v[iA], pointB = xmul(xfB, f->pB->v[iB]);
    return vdot(vsub(pointB, pointA), f->axis);
  } else if (f->type == 1) {
    v2 normal = f->axis, pointA = f->local_point, pointB = xmul(xfB, f->pB->v[iB]);
    return vdot(vsub(pointB, pointA), normal);
  } else {
    v2 normal = rmul(xfB.q, f->axis), pointB = xmul(xfB, f->local_point), pointA = f->pA->v[iA];
    return vdot(vsub(pointA, pointB), normal);
  }
}

/* b2TimeOfImpact with tMax = 1: returns the state (1 failed, 2 overlapped, 3 touching, 4 separated) and *t_out */
static int time_of_impact(const proxy_t* pA, const proxy_t* pB, sweep_t sB, float* t_out) {
  const float tMax = 1.0f;
  { /* b2Sweep::Normalize */
    const float two_pi = 2.0f * B2_PI;
    float d = two_pi * floorf(sB.a0 / two_pi);
    sB.a0 -= d;
    sB.a -= d;
  }
  const float total_radius = POLYGON_RADIUS + POLYGON_RADIUS;
  const float target = fmaxf(LINEAR_SLOP, total_radius - 3.0f * LINEAR_SLOP), tolerance = 0.25f * LINEAR_SLOP;
  float t1 = 0.0f;
  int iter = 0, state = 0;
  *t_out = tMax;
  scache_t cache;
  cache.count = 0;
  for (;;) {
    xf_t xfB = sweep_xf(&sB, t1);
    float distance = gjk_distance(&cache, pA, pB, xfB);
    if (distance <= 0.0f) { state = 2; *t_out = 0.0f; break; }
    if (distance < target + tolerance) { state = 3; *t_out = t1; break; }
    sepfn_t fcn;
    sep_init(&fcn, &cache, pA, pB, &sB, t1);
    int done = 0, push_back = 0;
    float t2 = tMax;
    for (;;) {
      int iA, iB;
      float s2 = sep_find_min(&fcn, &iA, &iB, t2);
      if (s2 > target + tolerance) { state = 4; *t_out = tMax; done = 1; break; }
      if (s2 > target - tolerance) { t1 = t2; break; }
      float s1 = sep_evaluate(&fcn, iA, iB, t1);
      if (s1 < target - tolerance) { state = 1; *t_out = t1; done = 1; break; }
      if (s1 <= target + tolerance) { state = 3; *t_out = t1; done = 1; break; }
      int root_iters = 0;
      float a1 = t1, a2 = t2;
      for (;;) {
        float t;
        if (root_iters & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
        else t = 0.5f * (a1 + a2);
        ++root_iters;
        float s = sep_evaluate(&fcn, iA, iB, t);
        if (fabsf(s - target) < tolerance) { t2 = t; break; }
        if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
        if (root_iters == 50) break;
      }
      ++push_back;
      if (push_back == MAX_POLY_VERTS) break;
    }
    ++iter;
    if (done) break;
    if (iter == 20) { state = 1; *t_out = t1; break; }
  }
  return state;
}

/* b2Body::SynchronizeFixtures + b2DynamicTree::MoveProxy for dynamic body d; returns 1 if the fat AABB moved */
static int sync_fixtures(lander_t* L, int d) {
  body_t* b = &L->b[1 + d];
  xf_t xf1;
  xf1.q = rot_set(b->a0);
  xf1.p = vsub(b->c0, rmul(xf1.q, b->local_center));
  aabb_t a1 = poly_aabb(&L->poly[d], xf1), a2 = poly_aabb(&L->poly[d], b->xf);
  aabb_t comb = {vmin(a1.lo, a2.lo), vmax(a1.hi, a2.hi)};
  v2 disp = vsub(b->xf.p, xf1.p);
  if (aabb_contains(L->poly_fat[d], comb)) return 0;
  aabb_t fb = fatten(comb);
  v2 dd = vscale(AABB_MULTIPLIER, disp);
  if (dd.x < 0.0f) fb.lo.x += dd.x; else fb.hi.x += dd.x;
  if (dd.y < 0.0f) fb.lo.y += dd.y; else fb.hi.y += dd.y;
  L->poly_fat[d] = fb;
  return 1;
}

/* b2Island::SolveTOI for the island {moon, body 1 + dyn} with the contacts idx[0..n) (idx[0] is the TOI contact): TOI
 * position solve (<= 20 iterations, Baumgarte 0.75), leap of faith, 180 velocity iterations without warm starting and
 * without joints, position integration over the rest of the step */
static void solve_toi_island(lander_t* L, int dyn, const int* idx, int n, float h, int vel_iters) {
  body_t* B = &L->b[1 + dyn];
  const float mB = B->inv_mass, iB = B->inv_I;
  const v2 lcB = B->local_center;
  v2 cB = B->c, vB = B->vel;
  float aB = B->a, wB = B->w;
  vc_t vcs[NPAIR];
  for (int k = 0; k < n; ++k) { /* b2ContactSolver constructor, warmStarting = false */
    vc_t* vc = &vcs[k];
    contact_t* c = &L->contact[idx[k]];
    memset(vc, 0, sizeof(*vc));
    vc->c = c;
    vc->friction = c->friction;
    vc->count = vc->pos_count = c->m.count;
    vc->type = c->m.type;
    vc->local_normal = c->m.local_normal; vc->local_point = c->m.local_point;
    for (int j = 0; j < vc->count; ++j) vc->local_points[j] = c->m.points[j].local_point;
  }
  for (int it = 0; it < 20; ++it) { /* SolveTOIPositionConstraints */
    float min_sep = 0.0f;
    for (int k = 0; k < n; ++k) {
      vc_t* vc = &vcs[k];
      for (int j = 0; j < vc->pos_count; ++j) {
        xf_t xfB;
        xfB.q = rot_set(aB);
        xfB.p = vsub(cB, rmul(xfB.q, lcB));
        v2 normal, point;
        float sep;
        if (vc->type == 1) {
          normal = vc->local_normal;
          v2 plane = vc->local_point, clip = xmul(xfB, vc->local_points[j]);
          sep = vdot(vsub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
          point = clip;
        } else {
          normal = rmul(xfB.q, vc->local_normal);
          v2 plane = xmul(xfB, vc->local_point), clip = vc->local_points[j];
          sep = vdot(vsub(clip, plane), normal) - POLYGON_RADIUS - POLYGON_RADIUS;
          point = clip;
          normal = vneg(normal);
        }
        v2 rB = vsub(point, cB);
        min_sep = fminf(min_sep, sep);
        float C = fclamp(TOI_BAUMGARTE * (sep + LINEAR_SLOP), -MAX_LINEAR_CORRECTION, 0.0f);
        float rnB = vcross(rB, normal);
        float K = mB + iB * rnB * rnB;
        float impulse = K > 0.0f ? -C / K : 0.0f;
        v2 Pi = vscale(impulse, normal);
        cB = vadd(cB, vscale(mB, Pi));
        aB += iB * vcross(rB, Pi);
      }
    }
    if (min_sep >= -1.5f * LINEAR_SLOP) break;
  }
  B->c0 = cB; B->a0 = aB; /* leap of faith to the new safe state (the moon's sweep does not change) */
  for (int k = 0; k < n; ++k) { /* InitializeVelocityConstraints */
    vc_t* vc = &vcs[k];
    xf_t xfB;
    xfB.q = rot_set(aB);
    xfB.p = vsub(cB, rmul(xfB.q, lcB));
    v2 pts[2];
    manifold_t m = vc->c->m;
    world_manifold(&m, xfB, &vc->normal, pts);
    for (int j = 0; j < vc->count; ++j) {
      vcp_t* p = &vc->p[j];
      p->rB = vsub(pts[j], cB);
      float rnB = vcross(p->rB, vc->normal);
      float kN = mB + iB * rnB * rnB;
      p->normal_mass = kN > 0.0f ? 1.0f / kN : 0.0f;
      v2 tangent = vcross_vs(vc->normal, 1.0f);
      float rtB = vcross(p->rB, tangent);
      float kT = mB + iB * rtB * rtB;
      p->tangent_mass = kT > 0.0f ? 1.0f / kT : 0.0f;
      p->velocity_bias = 0.0f;
    }
    if (vc->count == 2) {
      float rn1B = vcross(vc->p[0].rB, vc->normal), rn2B = vcross(vc->p[1].rB, vc->normal);
      float k11 = mB + iB * rn1B * rn1B, k22 = mB + iB * rn2B * rn2B, k12 = mB + iB * rn1B * rn2B;
      if (k11 * k11 < 1000.0f * (k11 * k22 - k12 * k12)) {
        vc->K[0][0] = k11; vc->K[0][1] = k12; vc->K[1][0] = k12; vc->K[1][1] = k22;
        float a = k11, b = k12, c = k12, d = k22, det = a * d - b * c;
        if (det != 0.0f) det = 1.0f / det;
        vc->NM[0][0] = det * d; vc->NM[1][0] = -det * b; vc->NM[0][1] = -det * c; vc->NM[1][1] = det * a;
      } else vc->count = 1;
    }
  }
  for (int it = 0; it < vel_iters; ++it) /* SolveVelocityConstraints: contacts only */
    for (int k = 0; k < n; ++k) {
      vc_t* vc = &vcs[k];
      v2 normal = vc->normal, tangent = vcross_vs(normal, 1.0f);
      for (int j = 0; j < vc->count; ++j) {
        vcp_t* p = &vc->p[j];
        v2 dv = vadd(vB, vcross_sv(wB, p->rB));
        float vt = vdot(dv, tangent) - 0.0f;
        float lambda = p->tangent_mass * (-vt);
        float maxf = vc->friction * p->normal_impulse;
        float newi = fclamp(p->tangent_impulse + lambda, -maxf, maxf);
        lambda = newi - p->tangent_impulse;
        p->tangent_impulse = newi;
        v2 Pi = vscale(lambda, tangent);
        vB = vadd(vB, vscale(mB, Pi));
        wB += iB * vcross(p->rB, Pi);
      }
      if (vc->count == 1) {
        vcp_t* p = &vc->p[0];
        v2 dv = vadd(vB, vcross_sv(wB, p->rB));
        float vn = vdot(dv, normal);
        float lambda = -p->normal_mass * (vn - p->velocity_bias);
        float newi = fmaxf(p->normal_impulse + lambda, 0.0f);
        lambda = newi - p->normal_impulse;
        p->normal_impulse = newi;
        v2 Pi = vscale(lambda, normal);
        vB = vadd(vB, vscale(mB, Pi));
        wB += iB * vcross(p->rB, Pi);
      } else {
        vcp_t *c1 = &vc->p[0], *c2 = &vc->p[1];
        float ax = c1->normal_impulse, ay = c2->normal_impulse;
        v2 dv1 = vadd(vB, vcross_sv(wB, c1->rB)), dv2 = vadd(vB, vcross_sv(wB, c2->rB));
        float vn1 = vdot(dv1, normal), vn2 = vdot(dv2, normal);
        float bx = vn1 - c1->velocity_bias, by = vn2 - c2->velocity_bias;
        bx -= vc->K[0][0] * ax + vc->K[1][0] * ay;
        by -= vc->K[0][1] * ax + vc->K[1][1] * ay;
        float xx, xy;
        int solved = 0;
        xx = -(vc->NM[0][0] * bx + vc->NM[1][0] * by);
        xy = -(vc->NM[0][1] * bx + vc->NM[1][1] * by);
        if (xx >= 0.0f && xy >= 0.0f) solved = 1;
        if (!solved) {
          xx = -c1->normal_mass * bx; xy = 0.0f;
          vn2 = vc->K[0][1] * xx + by;
          if (xx >= 0.0f && vn2 >= 0.0f) solved = 1;
        }
        if (!solved) {
          xx = 0.0f; xy = -c2->normal_mass * by;
          vn1 = vc->K[1][0] * xy + bx;
          if (xy >= 0.0f && vn1 >= 0.0f) solved = 1;
        }
        if (!solved) {
          xx = 0.0f; xy = 0.0f;
          if (bx >= 0.0f && by >= 0.0f) solved = 1;
        }
        if (solved) {
          float dx = xx - ax, dy = xy - ay;
          v2 P1 = vscale(dx, normal), P2 = vscale(dy, normal);
          vB = vadd(vB, vscale(mB, vadd(P1, P2)));
          wB += iB * (vcross(c1->rB, P1) + vcross(c2->rB, P2));
          c1->normal_impulse = xx; c2->normal_impulse = xy;
        }
      }
    }
  /* the TOI impulses are not stored for warm starting; integrate positions over the rest of the step */
  v2 tr = vscale(h, vB);
  if (vdot(tr, tr) > MAX_TRANSLATION * MAX_TRANSLATION) { float ratio = MAX_TRANSLATION / vlen(tr); vB = vscale(ratio, vB); }
  float rotn = h * wB;
  if (rotn * rotn > MAX_ROTATION * MAX_ROTATION) { float ratio = MAX_ROTATION / fabsf(rotn); wB *= ratio; }
  cB = vadd(cB, vscale(h, vB));
  aB += h * wB;
  B->c = cB; B->a = aB; B->vel = vB; B->w = wB;
  sync_transform(B);
}

/* b2World::SolveTOI(step) with m_stepComplete = true on entry (no sub-stepping mode) */
/* The CUDA engine's shortcut around b2TimeOfImpact (gymnasium_b200/csrc/lunarlander.cu: toi_clearly_separated), restated
 * here so that the oracle can count how often it applies and check that it never changes a result.  If every vertex of the
 * polygon stays on one side of the edge's line by more than TOI_CLEAR_MARGIN during the whole sweep, no point of the
 * polygon ever comes within the TOI target distance (0.005 + 0.00125) of any point of the edge, so b2TimeOfImpact cannot
 * return e_touching and SolveTOI's alpha is 1 whatever the root finder does.  The sweep is bounded from the end pose (the
 * body's current transform): a vertex at sweep time t lies within |c - c0| projected on the normal, plus r_max * |a - a0|
 * for the rotation, of its end position. */
#define TOI_CLEAR_MARGIN 0.05f
#define TOI_CLEAR_RMAX 0.8f /* > the largest vertex distance from a body's centre of mass (lander 0.61, legs 0.28) */
static int toi_clearly_separated(const v2* ev, const poly_t* p, const body_t* b) {
  const v2 d = vsub(ev[1], ev[0]);
  const float len = vlen(d);
  if (!(len > 1e-3f)) return 0;
  const v2 n = V(d.y / len, -d.x / len);
  const float da = fabsf(b->a - b->a0);
  if (!(da < 0.5f)) return 0;
  const float slack = fabsf(vdot(n, vsub(b->c, b->c0))) + TOI_CLEAR_RMAX * da + TOI_CLEAR_MARGIN;
  const v2 u = V(d.x / len, d.y / len);
  const float slack_u = fabsf(vdot(u, vsub(b->c, b->c0))) + TOI_CLEAR_RMAX * da + TOI_CLEAR_MARGIN;
  float lo = 1e30f, hi = -1e30f, ulo = 1e30f, uhi = -1e30f;
  for (int i = 0; i < p->count; ++i) {
    const v2 v = p->v[i];
    const v2 w = V((b->xf.q.c * v.x - b->xf.q.s * v.y) + b->xf.p.x, (b->xf.q.s * v.x + b->xf.q.c * v.y) + b->xf.p.y);
    const v2 r = vsub(w, ev[0]);
    const float sd = vdot(n, r), su = vdot(u, r);
    lo = fminf(lo, sd);
    hi = fmaxf(hi, sd);
    ulo = fminf(ulo, su);
    uhi = fmaxf(uhi, su);
  }
  /* beside the line, or beyond one of the segment's ends along it */
  return lo > slack || hi < -slack || ulo > len + slack_u || uhi < -slack_u;
}

static void solve_toi(lander_t* L, float dt, int vel_iters) {
  for (int i = 0; i < NBODY; ++i) L->b[i].alpha0 = 0.0f;
  for (int k = 0; k < NPAIR; ++k) { L->contact[k].toi_flag = 0; L->contact[k].toi_count = 0; L->contact[k].toi = 1.0f; }
  body_t* moon = &L->b[0];
  for (;;) {
    int idx[NPAIR], n = contacts_sorted(L, idx, -1), min_k = -1;
    float min_alpha = 1.0f;
    for (int t = 0; t < n; ++t) {
      contact_t* c = &L->contact[idx[t]];
      if (!c->enabled) continue;
      if (c->toi_count > MAX_SUB_STEPS) continue;
      float alpha = 1.0f;
      if (c->toi_flag) alpha = c->toi;
      else {
        int dyn = idx[t] / NEDGE, e = idx[t] % NEDGE;
        body_t* bB = &L->b[1 + dyn];
        if (!bB->awake) continue; /* activeA is false (static), activeB = awake; collideA is true (not dynamic) */
        float alpha0 = moon->alpha0; /* put the sweeps onto the same time interval */
        if (moon->alpha0 < bB->alpha0) {
          alpha0 = bB->alpha0;
          moon->alpha0 = alpha0; /* b2Sweep::Advance of a sweep with c0 == c, a0 == a */
        } else if (bB->alpha0 < moon->alpha0) {
          alpha0 = moon->alpha0;
          sweep_t s = body_sweep(bB);
          sweep_advance(&s, alpha0);
          body_set_sweep(bB, &s);
        }
        v2 ev[2] = {L->edge_v1[e], L->edge_v2[e]};
        proxy_t pA = {ev, 2}, pB = {L->poly[dyn].v, L->poly[dyn].count};
        float beta;
        int state = time_of_impact(&pA, &pB, body_sweep(bB), &beta);
        ++L->toi_calls;
        alpha = state == 3 ? fminf(alpha0 + (1.0f - alpha0) * beta, 1.0f) : 1.0f;
        if (toi_clearly_separated(ev, &L->poly[dyn], bB)) {
          ++L->toi_clear;
          if (alpha != 1.0f) ++L->toi_clear_wrong;
        }
        c->toi = alpha;
        c->toi_flag = 1;
      }
      if (alpha < min_alpha) { min_k = idx[t]; min_alpha = alpha; }
    }
    if (min_k < 0 || 1.0f - 10.0f * B2_EPSILON < min_alpha) break; /* no more TOI events */
    contact_t* mc = &L->contact[min_k];
    const int dyn = min_k / NEDGE;
    body_t* bB = &L->b[1 + dyn];
    const float moon_backup = moon->alpha0;
    const sweep_t backup = body_sweep(bB);
    moon->alpha0 = min_alpha; /* bA->Advance(minAlpha) on the static body */
    body_advance(bB, min_alpha);
    contact_update(L, min_k); /* the TOI contact likely has some new contact points */
    mc->toi_flag = 0;
    ++mc->toi_count;
    if (!mc->enabled || !mc->touching) { /* not solid: restore the sweeps */
      mc->enabled = 0;
      moon->alpha0 = moon_backup;
      body_set_sweep(bB, &backup);
      sync_transform(bB);
      continue;
    }
    set_awake(bB, 1);
    ++L->toi_events;
    /* island: the TOI contact, then the other contacts of bB (most recent first) that touch at the advanced pose */
    int isl[NPAIR], ni = 0;
    isl[ni++] = min_k;
    {
      int bidx[NPAIR], bn = contacts_sorted(L, bidx, dyn);
      for (int t = 0; t < bn; ++t) {
        if (bidx[t] == min_k) continue;
        contact_update(L, bidx[t]); /* other = the moon, already in the island: not advanced */
        if (!L->contact[bidx[t]].enabled || !L->contact[bidx[t]].touching) continue;
        isl[ni++] = bidx[t];
      }
    }
    solve_toi_island(L, dyn, isl, ni, (1.0f - min_alpha) * dt, vel_iters);
    int moved[NDYN] = {0, 0, 0};
    moved[dyn] = sync_fixtures(L, dyn);
    { /* invalidate all contact TOIs on this displaced body */
      int bidx[NPAIR], bn = contacts_sorted(L, bidx, dyn);
      for (int t = 0; t < bn; ++t) L->contact[bidx[t]].toi_flag = 0;
    }
    find_new_contacts(L, moved);
  }
}

/* b2World::Step(dt, 180, 60): new-fixture pairs, Collide, Solve (+ SynchronizeFixtures / FindNewContacts), SolveTOI,
 * ClearForces */
static void world_step(lander_t* L, float dt, int vel_iters, int pos_iters) {
  int moved[NDYN] = {1, 1, 1};
  if (L->new_fixture) { find_new_contacts(L, moved); L->new_fixture = 0; }
  float inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
  float dt_ratio = L->inv_dt0 * dt;
  collide(L);
  int was_awake = L->b[1].awake || L->b[2].awake || L->b[3].awake;
  solve_island(L, dt, dt_ratio, vel_iters, pos_iters);
  if (was_awake) {
    /* SynchronizeFixtures for island bodies (most recently created first) + FindNewContacts */
    for (int d = NDYN - 1; d >= 0; --d) moved[d] = sync_fixtures(L, d);
    find_new_contacts(L, moved);
  }
  if (dt > 0.0f) solve_toi(L, dt, vel_iters); /* m_continuousPhysics && step.dt > 0 */
  if (dt > 0.0f) L->inv_dt0 = inv_dt;
  for (int i = 1; i < NBODY; ++i) { L->b[i].force = V(0, 0); L->b[i].torque = 0.0f; } /* ClearForces */
}

/* ------------------------------------------------------------------------------------------------------------------
 * environment (lunar_lander.py) */
#define PI_D 3.141592653589793 /* math.pi */
#define FPS 50
#define SCALE 30.0
#define MAIN_ENGINE_POWER 13.0
#define SIDE_ENGINE_POWER 0.6
#define INITIAL_RANDOM 1000.0
#define LEG_AWAY 20
#define LEG_DOWN 18
#define LEG_W 2
#define LEG_H 8
#define LEG_SPRING_TORQUE 40
#define SIDE_ENGINE_HEIGHT 14
#define SIDE_ENGINE_AWAY 12
#define MAIN_ENGINE_Y_LOCATION 4
#define VIEWPORT_W 600
#define VIEWPORT_H 400

static void apply_linear_impulse(body_t* b, v2 impulse, v2 point) {
  if (!b->awake) set_awake(b, 1);
  b->vel = vadd(b->vel, vscale(b->inv_mass, impulse));
  b->w += b->inv_I * vcross(vsub(point, b->c), impulse);
}

typedef struct { double obs[8]; double reward; int terminated; } step_out_t;

static void env_step(lander_t* L, int action, const float* caction, step_out_t* out);

/* LunarLander.reset (lunar_lander.py:321-447); the RNG stream continues unless the caller re-seeded it */
static void env_reset(lander_t* L, float gravity, step_out_t* out) {
  pcg64_t rng = L->rng;
  const int continuous = L->continuous, enable_wind = L->enable_wind;
  const double wind_power = L->wind_power, turbulence_power = L->turbulence_power;
  memset(L, 0, sizeof(*L));
  L->rng = rng;
  L->gravity = gravity;
  L->continuous = continuous; L->enable_wind = enable_wind; L->wind_power = wind_power; L->turbulence_power = turbulence_power;
  const double W = VIEWPORT_W / SCALE, H = VIEWPORT_H / SCALE;
  enum { CHUNKS = 11 };
  double height[CHUNKS + 1], chunk_x[CHUNKS], smooth_y[CHUNKS];
  for (int i = 0; i < CHUNKS + 1; ++i) height[i] = pcg64_uniform(&L->rng, 0, H / 2);
  for (int i = 0; i < CHUNKS; ++i) chunk_x[i] = W / (CHUNKS - 1) * i;
  L->helipad_y = H / 4;
  for (int k = -2; k <= 2; ++k) height[CHUNKS / 2 + k] = L->helipad_y;
  for (int i = 0; i < CHUNKS; ++i) {
    int im = i - 1 < 0 ? CHUNKS : i - 1; /* height[-1] wraps to the last element (index CHUNKS) */
    smooth_y[i] = 0.33 * (height[im] + height[i + 0] + height[i + 1]);
  }
  L->edge_v1[0] = V(0.0f, 0.0f); L->edge_v2[0] = V((float)W, 0.0f); L->edge_friction[0] = 0.2f;
  for (int i = 0; i < CHUNKS - 1; ++i) {
    L->edge_v1[i + 1] = V((float)chunk_x[i], (float)smooth_y[i]);
    L->edge_v2[i + 1] = V((float)chunk_x[i + 1], (float)smooth_y[i + 1]);
    L->edge_friction[i + 1] = 0.1f;
  }
  for (int e = 0; e < NEDGE; ++e) {
    v2 lo = vmin(L->edge_v1[e], L->edge_v2[e]), hi = vmax(L->edge_v1[e], L->edge_v2[e]);
    aabb_t t = {V(lo.x - POLYGON_RADIUS, lo.y - POLYGON_RADIUS), V(hi.x + POLYGON_RADIUS, hi.y + POLYGON_RADIUS)};
    L->edge_fat[e] = fatten(t);
  }
  const double initial_y = VIEWPORT_H / SCALE, initial_x = VIEWPORT_W / SCALE / 2;
  static const int LANDER_POLY[6][2] = {{-14, 17}, {-17, 0}, {-17, -10}, {17, -10}, {17, 0}, {14, 17}};
  v2 lv[6];
  for (int i = 0; i < 6; ++i) lv[i] = V((float)(LANDER_POLY[i][0] / SCALE), (float)(LANDER_POLY[i][1] / SCALE));
  poly_set(&L->poly[0], lv, 6);
  L->poly_friction[0] = 0.1f;
  body_init_dynamic(&L->b[1], &L->poly[0], 5.0f, V((float)initial_x, (float)initial_y), 0.0f);
  /* ApplyForceToCenter(uniform(+-1000), uniform(+-1000)) */
  double fx = pcg64_uniform(&L->rng, -INITIAL_RANDOM, INITIAL_RANDOM);
  double fy = pcg64_uniform(&L->rng, -INITIAL_RANDOM, INITIAL_RANDOM);
  L->b[1].force = vadd(L->b[1].force, V((float)fx, (float)fy));
  if (L->enable_wind) { /* lunar_lander.py:401-403: two bounded integers from the env's stream */
    L->wind_idx = pcg64_integers(&L->rng, -9999, 9999);
    L->torque_idx = pcg64_integers(&L->rng, -9999, 9999);
  }
  for (int k = 0; k < 2; ++k) {
    int i = k == 0 ? -1 : +1;
    poly_set_box(&L->poly[1 + k], (float)(LEG_W / SCALE), (float)(LEG_H / SCALE));
    L->poly_friction[1 + k] = 0.2f;
    body_init_dynamic(&L->b[2 + k], &L->poly[1 + k], 1.0f, V((float)(initial_x - i * LEG_AWAY / SCALE), (float)initial_y),
                      (float)(i * 0.05));
    joint_t* j = &L->joint[k];
    memset(j, 0, sizeof(*j));
    j->bodyB = 2 + k;
    j->local_anchor_a = V(0.0f, 0.0f);
    j->local_anchor_b = V((float)(i * LEG_AWAY / SCALE), (float)(LEG_DOWN / SCALE));
    j->reference_angle = L->b[2 + k].a - L->b[1].a; /* pybox2d: bodyB.angle - bodyA.angle when not given */
    j->max_motor_torque = (float)LEG_SPRING_TORQUE;
    j->motor_speed = (float)(+0.3 * i);
    if (i == -1) { j->lower = (float)(+0.9 - 0.5); j->upper = (float)+0.9; }
    else { j->lower = (float)-0.9; j->upper = (float)(-0.9 + 0.5); }
  }
  L->b[0].dynamic = 0;
  L->b[0].xf.q.c = 1.0f;
  for (int d = 0; d < NDYN; ++d) L->poly_fat[d] = fatten(poly_aabb(&L->poly[d], L->b[1 + d].xf));
  L->new_fixture = 1;
  const float zero2[2] = {0.0f, 0.0f};
  env_step(L, 0, zero2, out); /* return self.step(np.array([0, 0]) if self.continuous else 0)[0] */
}

/* LunarLander.step (lunar_lander.py:471-665): `action` for the discrete env, `caction` (float32 [2]) for continuous=True */
static void env_step(lander_t* L, int action, const float* caction, step_out_t* out) {
  body_t* lander = &L->b[1];
  if (L->enable_wind && !(L->leg_contact[0] || L->leg_contact[1])) { /* lunar_lander.py:476-506 */
    double s1, s2, c_unused;
    det_sincos(0.02 * (double)L->wind_idx, &s1, &c_unused);
    det_sincos((PI_D * 0.01) * (double)L->wind_idx, &s2, &c_unused);
    const double wind_mag = det_tanh(s1 + s2) * L->wind_power;
    L->wind_idx += 1;
    if (!lander->awake) set_awake(lander, 1); /* ApplyForceToCenter(..., wake=True) */
    lander->force = vadd(lander->force, V((float)wind_mag, 0.0f));
    det_sincos(0.02 * (double)L->torque_idx, &s1, &c_unused);
    det_sincos((PI_D * 0.01) * (double)L->torque_idx, &s2, &c_unused);
    const double torque_mag = det_tanh(s1 + s2) * L->turbulence_power;
    L->torque_idx += 1;
    lander->torque += (float)torque_mag; /* ApplyTorque */
  }
  /* continuous: action = np.clip(action, -1, +1).astype(np.float64) (clipped in float32, then widened) */
  double a0 = 0.0, a1 = 0.0;
  if (L->continuous) {
    const float c0 = fminf(fmaxf(caction[0], -1.0f), 1.0f), c1 = fminf(fmaxf(caction[1], -1.0f), 1.0f);
    a0 = (double)c0; a1 = (double)c1;
  }
  const int fire_main = L->continuous ? (a0 > 0.0) : (action == 2);
  const int fire_side = L->continuous ? (fabs(a1) > 0.5) : (action == 1 || action == 3);
  double tip0, tip1; /* tip = (math.sin(angle), math.cos(angle)) */
  det_sincos((double)lander->a, &tip0, &tip1);
  const double side0 = -tip1, side1 = tip0;
  double dispersion[2];
  dispersion[0] = pcg64_uniform(&L->rng, -1.0, +1.0) / SCALE;
  dispersion[1] = pcg64_uniform(&L->rng, -1.0, +1.0) / SCALE;
  double m_power = 0.0, s_power = 0.0;
  if (fire_main) {
    m_power = L->continuous ? (fmin(fmax(a0, 0.0), 1.0) + 1.0) * 0.5 : 1.0; /* 0.5..1.0 */
    double ox = tip0 * (MAIN_ENGINE_Y_LOCATION / SCALE + 2 * dispersion[0]) + side0 * dispersion[1];
    double oy = -tip1 * (MAIN_ENGINE_Y_LOCATION / SCALE + 2 * dispersion[0]) - side1 * dispersion[1];
    double px = (double)lander->xf.p.x + ox, py = (double)lander->xf.p.y + oy;
    apply_linear_impulse(lander, V((float)(-ox * MAIN_ENGINE_POWER * m_power), (float)(-oy * MAIN_ENGINE_POWER * m_power)),
                         V((float)px, (float)py));
  }
  if (fire_side) {
    double direction = L->continuous ? (a1 > 0 ? 1.0 : -1.0) : (double)(action - 2); /* np.sign(action[1]) */
    s_power = L->continuous ? fmin(fmax(fabs(a1), 0.5), 1.0) : 1.0;
    double ox = tip0 * dispersion[0] + side0 * (3 * dispersion[1] + direction * SIDE_ENGINE_AWAY / SCALE);
    double oy = -tip1 * dispersion[0] - side1 * (3 * dispersion[1] + direction * SIDE_ENGINE_AWAY / SCALE);
    double px = (double)lander->xf.p.x + ox - tip0 * 17 / SCALE;
    double py = (double)lander->xf.p.y + oy + tip1 * SIDE_ENGINE_HEIGHT / SCALE;
    apply_linear_impulse(lander, V((float)(-ox * SIDE_ENGINE_POWER * s_power), (float)(-oy * SIDE_ENGINE_POWER * s_power)),
                         V((float)px, (float)py));
  }
  world_step(L, (float)(1.0 / FPS), 6 * 30, 2 * 30);
  const double posx = lander->xf.p.x, posy = lander->xf.p.y, velx = lander->vel.x, vely = lander->vel.y;
  double* s = out->obs;
  s[0] = (posx - VIEWPORT_W / SCALE / 2) / (VIEWPORT_W / SCALE / 2);
  s[1] = (posy - (L->helipad_y + LEG_DOWN / SCALE)) / (VIEWPORT_H / SCALE / 2);
  s[2] = velx * (VIEWPORT_W / SCALE / 2) / FPS;
  s[3] = vely * (VIEWPORT_H / SCALE / 2) / FPS;
  s[4] = (double)lander->a;
  s[5] = 20.0 * (double)lander->w / FPS;
  s[6] = L->leg_contact[0] ? 1.0 : 0.0;
  s[7] = L->leg_contact[1] ? 1.0 : 0.0;
  double reward = 0;
  double shaping = -100 * sqrt(s[0] * s[0] + s[1] * s[1]) - 100 * sqrt(s[2] * s[2] + s[3] * s[3]) - 100 * fabs(s[4]) +
                   10 * s[6] + 10 * s[7];
  if (L->has_prev_shaping) reward = shaping - L->prev_shaping;
  L->prev_shaping = shaping;
  L->has_prev_shaping = 1;
  reward -= m_power * 0.30;
  reward -= s_power * 0.03;
  int terminated = 0;
  if (L->game_over || fabs(s[0]) >= 1.0) { terminated = 1; reward = -100; }
  if (!lander->awake) { terminated = 1; reward = +100; }
  out->reward = reward;
  out->terminated = terminated;
}

/* ------------------------------------------------------------------------------------------------------------------
 * vector API used by oracle/lunar_lander.py (ctypes): SyncVectorEnv NEXT_STEP semantics + TimeLimit */
typedef struct {
  int n, max_episode_steps;
  float gravity;
  lander_t* env;
  int *elapsed, *autoreset;
} ll_vec_t;

ll_vec_t* ll_create(int n, int max_episode_steps, double gravity) {
  ll_vec_t* v = (ll_vec_t*)calloc(1, sizeof(ll_vec_t));
  v->n = n; v->max_episode_steps = max_episode_steps; v->gravity = (float)gravity;
  v->env = (lander_t*)calloc((size_t)n, sizeof(lander_t));
  for (int i = 0; i < n; ++i) { v->env[i].wind_power = 15.0; v->env[i].turbulence_power = 1.5; }
  v->elapsed = (int*)calloc((size_t)n, sizeof(int));
  v->autoreset = (int*)calloc((size_t)n, sizeof(int));
  return v;
}
/* LunarLander(continuous=..., enable_wind=..., wind_power=..., turbulence_power=...); call before the first reset */
void ll_configure(ll_vec_t* v, int continuous, int enable_wind, double wind_power, double turbulence_power) {
  for (int i = 0; i < v->n; ++i) {
    lander_t* L = &v->env[i];
    L->continuous = continuous; L->enable_wind = enable_wind; L->wind_power = wind_power; L->turbulence_power = turbulence_power;
  }
}
void ll_destroy(ll_vec_t* v) {
  if (!v) return;
  free(v->env); free(v->elapsed); free(v->autoreset); free(v);
}
static void write_obs(float* obs, const step_out_t* o) { for (int k = 0; k < 8; ++k) obs[k] = (float)o->obs[k]; }

/* seeds may be NULL (streams continue); mask may be NULL (all) */
void ll_reset(ll_vec_t* v, const uint64_t* seeds, const uint8_t* mask, float* obs) {
  for (int i = 0; i < v->n; ++i) {
    if (mask && !mask[i]) continue;
    if (seeds) pcg64_seed(&v->env[i].rng, seeds[i]);
    step_out_t o;
    env_reset(&v->env[i], v->gravity, &o);
    write_obs(obs + 8 * i, &o);
    v->elapsed[i] = 0; v->autoreset[i] = 0;
  }
}
/* actions: int64 [n] (discrete) or, for continuous=True, float32 [n][2] */
void ll_step(ll_vec_t* v, const void* actions_, float* obs, double* reward, uint8_t* terminated, uint8_t* truncated) {
  const int64_t* actions = (const int64_t*)actions_;
  const float* cactions = (const float*)actions_;
  for (int i = 0; i < v->n; ++i) {
    step_out_t o;
    if (v->autoreset[i]) {
      env_reset(&v->env[i], v->gravity, &o);
      write_obs(obs + 8 * i, &o);
      reward[i] = 0.0; terminated[i] = 0; truncated[i] = 0;
      v->elapsed[i] = 0; v->autoreset[i] = 0;
      continue;
    }
    if (v->env[i].continuous) env_step(&v->env[i], 0, cactions + 2 * i, &o);
    else env_step(&v->env[i], (int)actions[i], cactions, &o);
    write_obs(obs + 8 * i, &o);
    reward[i] = o.reward;
    terminated[i] = (uint8_t)o.terminated;
    v->elapsed[i] += 1;
    truncated[i] = v->max_episode_steps > 0 && v->elapsed[i] >= v->max_episode_steps;
    v->autoreset[i] = terminated[i] || truncated[i];
  }
}
/* introspection for tests: body state [3][7] = x, y, angle, vx, vy, w, awake; misc[8] */
void ll_debug_state(const ll_vec_t* v, int i, float* bodies, float* misc) {
  const lander_t* L = &v->env[i];
  for (int b = 0; b < 3; ++b) {
    const body_t* B = &L->b[1 + b];
    float* o = bodies + 7 * b;
    o[0] = B->xf.p.x; o[1] = B->xf.p.y; o[2] = B->a; o[3] = B->vel.x; o[4] = B->vel.y; o[5] = B->w; o[6] = (float)B->awake;
  }
  misc[0] = L->b[1].mass; misc[1] = L->b[1].I; misc[2] = L->b[2].mass; misc[3] = L->b[2].I;
  misc[4] = L->b[1].local_center.x; misc[5] = L->b[1].local_center.y;
  int nc = 0, nt = 0;
  for (int k = 0; k < NPAIR; ++k) { nc += L->contact[k].exists; nt += L->contact[k].touching; }
  misc[6] = (float)nc; misc[7] = (float)nt;
}
void ll_wind_state(const ll_vec_t* v, int i, int64_t* out /* [2]: wind_idx, torque_idx */) {
  out[0] = v->env[i].wind_idx; out[1] = v->env[i].torque_idx;
}
void ll_toi_stats(const ll_vec_t* v, int i, long* out /* [2]: b2TimeOfImpact calls, solid TOI events since reset */) {
  out[0] = v->env[i].toi_calls; out[1] = v->env[i].toi_events;
}
void ll_toi_shortcut_stats(const ll_vec_t* v, int i, long* out /* [2]: evaluations the shortcut covers, of those not alpha=1 */) {
  out[0] = v->env[i].toi_clear; out[1] = v->env[i].toi_clear_wrong;
}
/* Direct probe of the restated b2TimeOfImpact: a box with half extents (hx, hy) (body origin = centroid) swept from
 * (c0, a0) to (c1, a1) against the edge e1-e2; returns the b2TOIOutput state (1 failed, 2 overlapped, 3 touching,
 * 4 separated) and writes t. */
int ll_toi_probe(const float* e, float hx, float hy, const float* c0, float a0, const float* c1, float a1, float* t) {
  v2 ev[2] = {V(e[0], e[1]), V(e[2], e[3])};
  v2 bv[4] = {V(-hx, -hy), V(hx, -hy), V(hx, hy), V(-hx, hy)};
  proxy_t pA = {ev, 2}, pB = {bv, 4};
  sweep_t s = {V(0, 0), V(c0[0], c0[1]), V(c1[0], c1[1]), a0, a1, 0.0f};
  return time_of_impact(&pA, &pB, s, t);
}
void ll_terrain(const ll_vec_t* v, int i, float* xy /* [11][4] */) {
  const lander_t* L = &v->env[i];
  for (int e = 0; e < NEDGE; ++e) {
    xy[4 * e] = L->edge_v1[e].x; xy[4 * e + 1] = L->edge_v1[e].y; xy[4 * e + 2] = L->edge_v2[e].x; xy[4 * e + 3] = L->edge_v2[e].y;
  }
}
