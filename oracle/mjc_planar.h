/* oracle/mjc_planar.h -- CPU oracle core for the planar MuJoCo robots (Hopper-v5, Walker2d-v5): a plain-C restatement of the
 * MuJoCo subset that gymnasium/envs/mujoco/{hopper,walker2d}_v5.py + mujoco_env.py drive on assets/hopper.xml /
 * assets/walker2d_v5.xml, plus the env logic itself.  Included by oracle/hopper.c and oracle/walker2d.c, which select the
 * robot with ROBOT_HOPPER / ROBOT_WALKER2D.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into or called by gymnasium_b200/.
 *
 * PARITY UNPINNED, like oracle/humanoid.c (the `mujoco` wheel is not installable here; no golden trajectories in the
 * reference).  Same pipeline as that file (mj_step = mj_forward + RK4; kinematics, comPos, crb, factorM, collision,
 * makeConstraint, projectConstraint, comVel, passive, rne, actuation, acceleration, PGS), generalised where the planar
 * robots differ from the humanoid: slide joints (root x / z), no free joint (nq == nv), joint `ref`, per-geom friction /
 * condim / contype / conaffinity, capsules given by size + pos + quat, contact solimp and margin from the geoms (Hopper
 * .8 .8 .01 and 0.001; Walker2d the defaults and 0), actuator gear and ctrlrange +-1.  Anchors on the reference's call sites:
 *   model        gymnasium/envs/mujoco/assets/hopper.xml:1-61, walker2d_v5.xml:1-70 (re-typed below as data)
 *   reset        hopper_v5.py:322-337 / walker2d_v5.py:321-336, mujoco_env.py:132-142, :172-187
 *   step         hopper_v5.py:271-305 / walker2d_v5.py:284-303 (mj_step(nstep=4))
 *   observation  hopper_v5.py:253-261 / walker2d_v5.py:274-282 (qpos[1:], clip(qvel, -10, 10))
 *   health       hopper_v5.py:231-247 / walker2d_v5.py:261-272
 * Known deviation: neither XML names a solver, so MuJoCo runs its default Newton solver (100 iterations, tolerance 1e-8) on
 * the pyramidal-cone problem; this file runs PGS on the same convex problem with the same stopping rule -- the solutions
 * agree to solver tolerance, the iterates do not.
 *
 * All arithmetic float64, no FMA contraction (compile with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(ROBOT_HOPPER)
#define NB 5
#define NQ 6
#define NV 6
#define NU 3
#define NJ 6
#define NG 5
#define MAXCON 8
#define MAXEFC 16
#define API(name) hp_##name
#elif defined(ROBOT_WALKER2D)
#define NB 8
#define NQ 9
#define NV 9
#define NU 6
#define NJ 9
#define NG 8
#define MAXCON 8
#define MAXEFC 32
#define API(name) w2_##name
#elif defined(ROBOT_HALFCHEETAH)
#define NB 8
#define NQ 9
#define NV 9
#define NU 6
#define NJ 9
#define NG 9
#define MAXCON 12
#define MAXEFC 56
#define API(name) hc_##name
#elif defined(ROBOT_INVPEND)
#define NB 3
#define NQ 2
#define NV 2
#define NU 1
#define NJ 2
#define NG 2
#define MAXCON 2
#define MAXEFC 4
#if defined(OPT_EULER) /* test-only instance: the same model under mj_Euler (oracle/inverted_pendulum_euler.c) */
#define API(name) ipe_##name
#else
#define API(name) ip_##name
#endif
#else
#error "define ROBOT_HOPPER, ROBOT_WALKER2D, ROBOT_HALFCHEETAH or ROBOT_INVPEND before including mjc_planar.h"
#endif
#define MINVAL 1e-15
#define PI 3.14159265358979323846

typedef unsigned __int128 u128;
typedef struct { u128 state, inc; } pcg64_t;
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
}
static double pcg64_double(pcg64_t* g) {
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
  g->state = g->state * mult + g->inc;
  uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state, x = hi ^ lo;
  unsigned rot = (unsigned)(hi >> 58);
  x = (x >> rot) | (x << ((64 - rot) & 63));
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* small vector helpers */
static inline void cp3(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return n; }
  double inv = 1.0 / n;
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
  return n;
}
static inline void mulmatvec3(double* r, const double* m, const double* v) { /* row-major 3x3 */
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmat3(double* r, const double* a, const double* b) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(r, t, sizeof(t));
}
/* deterministic sin/cos (same fixed IEEE sequence as oracle/lunar_lander.c and the CUDA engine) */
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
/* quaternions (w, x, y, z) */
static inline void quat_mul(double* r, const double* a, const double* b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof(t));
}
static inline void quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double inv = 1.0 / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
static inline void quat_axisangle(double* q, const double* axis, double angle) {
  double s, c;
  det_sincos(0.5 * angle, &s, &c);
  q[0] = c; q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
static inline void quat2mat(double* m, const double* q) { /* mju_quat2Mat */
  const double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  const double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
static inline void quat_rot(double* r, const double* q, const double* v) {
  double m[9];
  quat2mat(m, q);
  mulmatvec3(r, m, v);
}

/* ------------------------------------------------------------------------------------------------------------------ */
/* model (compiled from humanoid.xml) */
enum { G_PLANE = 0, G_SPHERE = 2, G_CAPSULE = 3 };
typedef struct {
  /* bodies */
  int parent[NB], body_jntadr[NB], body_jntnum[NB], body_dofadr[NB], body_dofnum[NB], body_lastdof[NB];
  double body_pos[NB][3], body_quat[NB][4], body_mass[NB], body_ipos[NB][3], body_inertia[NB][9] /* about com, body frame */;
  double subtree_mass[NB], body_invweight0[NB][2];
  /* joints */
  int jnt_type[NJ] /*2 slide, 3 hinge*/, jnt_body[NJ], jnt_qposadr[NJ], jnt_dofadr[NJ], jnt_limited[NJ];
  double jnt_pos[NJ][3], jnt_axis[NJ][3], jnt_range[NJ][2], jnt_stiffness[NJ];
  /* dofs */
  int dof_body[NV], dof_jnt[NV], dof_parent[NV];
  double dof_armature[NV], dof_damping[NV], dof_invweight0[NV];
  /* geoms */
  int geom_type[NG], geom_body[NG], geom_condim[NG], geom_contype[NG], geom_conaffinity[NG];
  double geom_pos[NG][3], geom_mat[NG][9], geom_size[NG][2], geom_rbound[NG], geom_friction[NG];
  /* actuators */
  int act_dof[NU];
  double act_gear[NU], act_ctrlrange[NU][2];
  /* options; solimp: default (joint limits), solimp_contact: the geoms' (hopper.xml:10) */
  double timestep, gravity[3], meaninertia, margin, solref[2], solimp[5], solimp_contact[5], tolerance;
  int iterations;
  /* collision pair list */
  int npair, pair_g1[160], pair_g2[160];
  double qpos0[NQ];
} model_t;

typedef struct {
  int g1, g2, dim, efc_adr;
  double dist, pos[3], frame[9], mu;
} contact_t;

typedef struct {
  double qpos[NQ], qvel[NV], qacc_warmstart[NV], time, ctrl[NU];
  double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], xanchor[NJ][3], xaxis[NJ][3];
  double geom_xpos[NG][3], geom_xmat[NG][9];
  double subtree_com[NB][3], cinert[NB][10], cdof[NV][6], cdof_dot[NV][6], cvel[NB][6], cfrc_ext[NB][6];
  double qM[NV][NV], qLD[NV][NV], qLDiagInv[NV];
  double qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], qfrc_smooth[NV], qacc_smooth[NV], qacc[NV], qfrc_constraint[NV];
  int ncon, nefc;
  contact_t con[MAXCON];
  double efc_J[MAXEFC][NV], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_D[MAXEFC], efc_R[MAXEFC], efc_aref[MAXEFC],
      efc_b[MAXEFC], efc_force[MAXEFC], efc_diagApprox[MAXEFC], efc_vel[MAXEFC];
  int efc_contact[MAXEFC]; /* row belongs to a contact (solimp_contact) or to a joint limit (solimp) */
  double efc_AR[MAXEFC][MAXEFC];
  int solver_iter;
} data_t;

/* --- model construction -------------------------------------------------------------------------------------------- */
#define DEG (PI / 180.0)
typedef struct { int parent; double pos[3]; } bdef_t;
/* joints in qpos order: type (2 slide, 3 hinge), body, pos, axis, limited, range (deg), armature, damping, ref */
typedef struct { int type, body; double pos[3], axis[3]; int limited; double lo, hi, armature, damping, ref; } jdef_t;
/* geoms: type, body, pos, quat, radius, half length, friction, condim, contype, conaffinity */
typedef struct { int type, body; double pos[3], quat[4], r, half, friction; int condim, contype, conaffinity; } gdef_t;
typedef struct { int joint; double gear; } adef_t;
#define FOOT_QUAT {0.70710678118654757, 0, -0.70710678118654746, 0}
#if defined(ROBOT_HOPPER)
/* hopper.xml: defaults joint armature 1 damping 1 limited (:9), geom condim 1 contype 1 conaffinity 1 margin 0.001
 * solimp .8 .8 .01 (:10); floor condim 3 (:19); RK4, timestep 0.002 (:12) */
static const char* BODY_NAMES[NB] = {"world", "torso", "thigh", "leg", "foot"};
static const bdef_t BODY_DEF[NB] = {
    {0, {0, 0, 0}}, {0, {0, 0, 1.25}}, {1, {0, 0, -0.19999999999999996}}, {2, {0, 0, -0.70000000000000007}}, {3, {0.13, 0, -0.35}}};
static const jdef_t JOINT_DEF[NJ] = {
    {2, 1, {0, 0, -1.25}, {1, 0, 0}, 0, 0, 0, 0, 0, 0},        /* rootx */
    {2, 1, {0, 0, -1.25}, {0, 0, 1}, 0, 0, 0, 0, 0, 1.25},     /* rootz, ref 1.25 */
    {3, 1, {0, 0, 0}, {0, 1, 0}, 0, 0, 0, 0, 0, 0},            /* rooty */
    {3, 2, {0, 0, 0}, {0, -1, 0}, 1, -150, 0, 1, 1, 0},        /* thigh_joint */
    {3, 3, {0, 0, 0.25}, {0, -1, 0}, 1, -150, 0, 1, 1, 0},     /* leg_joint */
    {3, 4, {-0.13, 0, 0.1}, {0, -1, 0}, 1, -45, 45, 1, 1, 0},  /* foot_joint */
};
static const gdef_t GEOM_DEF[NG] = {
    {G_PLANE, 0, {0, 0, 0}, {1, 0, 0, 0}, 0, 0, 1.0, 3, 1, 1},                                           /* floor */
    {G_CAPSULE, 1, {0, 0, 0}, {1, 0, 0, 0}, 0.05, 0.19999999999999996, 0.9, 1, 1, 1},                    /* torso_geom */
    {G_CAPSULE, 2, {0, 0, -0.22500000000000009}, {1, 0, 0, 0}, 0.05, 0.22500000000000003, 0.9, 1, 1, 1}, /* thigh_geom */
    {G_CAPSULE, 3, {0, 0, 0}, {1, 0, 0, 0}, 0.04, 0.25, 0.9, 1, 1, 1},                                   /* leg_geom */
    {G_CAPSULE, 4, {-0.065, 0, 0.1}, FOOT_QUAT, 0.06, 0.195, 2.0, 1, 1, 1},                              /* foot_geom */
};
static const adef_t ACT_DEF[NU] = {{3, 200.0}, {4, 200.0}, {5, 200.0}};
#define OPT_MARGIN 0.001
#define OPT_SOLIMP_CONTACT {0.8, 0.8, 0.01, 0.5, 2.0}
#elif defined(ROBOT_HALFCHEETAH)
/* half_cheetah.xml: compiler angle="radian" settotalmass="14" (:35); defaults joint armature .1 damping .01 limited
 * solimplimit 0 .8 .03 solreflimit .02 1 stiffness 8 (:37; every hinge overrides damping / stiffness), geom conaffinity 0
 * condim 3 contype 1 friction .4 solimp 0 .8 .01 solref .02 1 (:38: the robot collides with the floor only); option
 * timestep 0.01, integrator left at its default = Euler (:42); the root joints carry armature 0 damping 0 stiffness 0 (:55-57);
 * motors gear 120 90 60 120 60 30, ctrlrange -1 1 (:88-93) */
static const char* BODY_NAMES[NB] = {"world", "torso", "bthigh", "bshin", "bfoot", "fthigh", "fshin", "ffoot"};
static const bdef_t BODY_DEF[NB] = {
    {0, {0, 0, 0}}, {0, {0, 0, 0.7}},
    {1, {-0.5, 0, 0}}, {2, {0.16, 0, -0.25}}, {3, {-0.28, 0, -0.14}},
    {1, {0.5, 0, 0}}, {5, {-0.14, 0, -0.24}}, {6, {0.13, 0, -0.18}}};
/* ranges in radians */
static const jdef_t JOINT_DEF[NJ] = {
    {2, 1, {0, 0, 0}, {1, 0, 0}, 0, 0, 0, 0, 0, 0},          /* rootx */
    {2, 1, {0, 0, 0}, {0, 0, 1}, 0, 0, 0, 0, 0, 0},          /* rootz */
    {3, 1, {0, 0, 0}, {0, 1, 0}, 0, 0, 0, 0, 0, 0},          /* rooty */
    {3, 2, {0, 0, 0}, {0, 1, 0}, 1, -0.52, 1.05, 0.1, 6, 0},   /* bthigh */
    {3, 3, {0, 0, 0}, {0, 1, 0}, 1, -0.785, 0.785, 0.1, 4.5, 0}, /* bshin */
    {3, 4, {0, 0, 0}, {0, 1, 0}, 1, -0.4, 0.785, 0.1, 3, 0},   /* bfoot */
    {3, 5, {0, 0, 0}, {0, 1, 0}, 1, -1, 0.7, 0.1, 4.5, 0},     /* fthigh */
    {3, 6, {0, 0, 0}, {0, 1, 0}, 1, -1.2, 0.87, 0.1, 3, 0},    /* fshin */
    {3, 7, {0, 0, 0}, {0, 1, 0}, 1, -0.5, 0.5, 0.1, 1.5, 0},   /* ffoot */
};
static const double JOINT_STIFFNESS[NJ] = {0, 0, 0, 240, 180, 120, 180, 120, 60};
/* geoms: the floor, then torso (fromto), head, and one capsule per leg segment given by pos + axisangle about y (radians) */
static const gdef_t GEOM_DEF[NG] = {
    {G_PLANE, 0, {0, 0, 0}, {1, 0, 0, 0}, 0, 0, 0.4, 3, 1, 1},
    {G_CAPSULE, 1, {0, 0, 0}, {1, 0, 0, 0}, 0.046, 0.5, 0.4, 3, 1, 0},          /* torso: fromto -.5 0 0 .5 0 0 */
    {G_CAPSULE, 1, {0.6, 0, 0.1}, {1, 0, 0, 0}, 0.046, 0.15, 0.4, 3, 1, 0},     /* head: axisangle 0 1 0 .87 */
    {G_CAPSULE, 2, {0.1, 0, -0.13}, {1, 0, 0, 0}, 0.046, 0.145, 0.4, 3, 1, 0},  /* bthigh: -3.8 */
    {G_CAPSULE, 3, {-0.14, 0, -0.07}, {1, 0, 0, 0}, 0.046, 0.15, 0.4, 3, 1, 0}, /* bshin: -2.03 */
    {G_CAPSULE, 4, {0.03, 0, -0.097}, {1, 0, 0, 0}, 0.046, 0.094, 0.4, 3, 1, 0}, /* bfoot: -.27 */
    {G_CAPSULE, 5, {-0.07, 0, -0.12}, {1, 0, 0, 0}, 0.046, 0.133, 0.4, 3, 1, 0}, /* fthigh: .52 */
    {G_CAPSULE, 6, {0.065, 0, -0.09}, {1, 0, 0, 0}, 0.046, 0.106, 0.4, 3, 1, 0}, /* fshin: -.6 */
    {G_CAPSULE, 7, {0.045, 0, -0.07}, {1, 0, 0, 0}, 0.046, 0.07, 0.4, 3, 1, 0},  /* ffoot: -.6 */
};
/* orientation / placement spec per geom: kind 0 = the quat above, 1 = axisangle about +y with angle a[0], 2 = fromto a[0..5] */
typedef struct { int kind; double a[6]; } gspec_t;
static const gspec_t GEOM_SPEC[NG] = {
    {0, {0}}, {2, {-0.5, 0, 0, 0.5, 0, 0}}, {1, {0.87}}, {1, {-3.8}}, {1, {-2.03}}, {1, {-0.27}}, {1, {0.52}}, {1, {-0.6}}, {1, {-0.6}}};
static const adef_t ACT_DEF[NU] = {{3, 120.0}, {4, 90.0}, {5, 60.0}, {6, 120.0}, {7, 60.0}, {8, 30.0}};
#define OPT_MARGIN 0.0
#define OPT_SOLIMP_CONTACT {0.0, 0.8, 0.01, 0.5, 2.0}
#define OPT_SOLIMP_LIMIT {0.0, 0.8, 0.03, 0.5, 2.0}
#define OPT_TIMESTEP 0.01
#define OPT_CTRLRANGE 1.0
#define FRAME_SKIP 5
#define OPT_ANGLE_RADIAN 1
#define OPT_TOTALMASS 14.0
#define OPT_EULER 1
#elif defined(ROBOT_INVPEND)
/* inverted_pendulum.xml: defaults joint armature 0 damping 1 limited (:4), geom contype 0 (:5: nothing collides), motor
 * ctrlrange -3 3 (:7); RK4, timestep 0.02 (:9); the rail (a world geom, :13) takes part in nothing and is left out.  Slide
 * ranges are lengths, hinge ranges degrees.  The pole's capsule is given by fromto (:19): see build_model. */
static const char* BODY_NAMES[NB] = {"world", "cart", "pole"};
static const bdef_t BODY_DEF[NB] = {{0, {0, 0, 0}}, {0, {0, 0, 0}}, {1, {0, 0, 0}}};
static const jdef_t JOINT_DEF[NJ] = {
    {2, 1, {0, 0, 0}, {1, 0, 0}, 1, -1, 1, 0, 1, 0},   /* slider */
    {3, 2, {0, 0, 0}, {0, 1, 0}, 1, -90, 90, 0, 1, 0}, /* hinge */
};
#define POLE_FROMTO {0, 0, 0, 0.001, 0, 0.6}
static const gdef_t GEOM_DEF[NG] = {
    {G_CAPSULE, 1, {0, 0, 0}, {0.707, 0, 0.707, 0}, 0.1, 0.1, 1.0, 3, 0, 1}, /* cart */
    {G_CAPSULE, 2, {0, 0, 0}, {1, 0, 0, 0}, 0.049, 0.3, 1.0, 3, 0, 1},        /* cpole: pos / quat / half length from fromto */
};
static const adef_t ACT_DEF[NU] = {{0, 100.0}};
#define OPT_MARGIN 0.0
#define OPT_SOLIMP_CONTACT {0.9, 0.95, 0.001, 0.5, 2.0}
#define OPT_TIMESTEP 0.02
#define OPT_CTRLRANGE 3.0
#define FRAME_SKIP 2
#else
/* walker2d_v5.xml: defaults joint armature 0.01 damping .1 limited (:10), geom condim 3 contype 1 conaffinity 0 friction .7
 * (:11: the robot's geoms collide with the floor only); floor conaffinity 1 (:16); RK4, timestep 0.002 (:13) */
static const char* BODY_NAMES[NB] = {"world", "torso", "thigh", "leg", "foot", "thigh_left", "leg_left", "foot_left"};
static const bdef_t BODY_DEF[NB] = {
    {0, {0, 0, 0}}, {0, {0, 0, 1.25}},
    {1, {0, 0, -0.19999999999999996}}, {2, {0, 0, -0.70000000000000007}}, {3, {0.20000000000000001, 0, -0.34999999999999998}},
    {1, {0, 0, -0.19999999999999996}}, {5, {0, 0, -0.70000000000000007}}, {6, {0.20000000000000001, 0, -0.34999999999999998}}};
static const jdef_t JOINT_DEF[NJ] = {
    {2, 1, {0, 0, -1.25}, {1, 0, 0}, 0, 0, 0, 0, 0, 0},
    {2, 1, {0, 0, -1.25}, {0, 0, 1}, 0, 0, 0, 0, 0, 1.25},
    {3, 1, {0, 0, 0}, {0, 1, 0}, 0, 0, 0, 0, 0, 0},
    {3, 2, {0, 0, 0}, {0, -1, 0}, 1, -150, 0, 0.01, 0.1, 0},
    {3, 3, {0, 0, 0.25}, {0, -1, 0}, 1, -150, 0, 0.01, 0.1, 0},
    {3, 4, {-0.20000000000000001, 0, 0.10000000000000001}, {0, -1, 0}, 1, -45, 45, 0.01, 0.1, 0},
    {3, 5, {0, 0, 0}, {0, -1, 0}, 1, -150, 0, 0.01, 0.1, 0},
    {3, 6, {0, 0, 0.25}, {0, -1, 0}, 1, -150, 0, 0.01, 0.1, 0},
    {3, 7, {-0.20000000000000001, 0, 0.10000000000000001}, {0, -1, 0}, 1, -45, 45, 0.01, 0.1, 0},
};
static const gdef_t GEOM_DEF[NG] = {
    {G_PLANE, 0, {0, 0, 0}, {1, 0, 0, 0}, 0, 0, 0.7, 3, 1, 1},
    {G_CAPSULE, 1, {0, 0, 0}, {1, 0, 0, 0}, 0.050000000000000003, 0.19999999999999996, 0.9, 3, 1, 0},
    {G_CAPSULE, 2, {0, 0, -0.22500000000000009}, {1, 0, 0, 0}, 0.050000000000000003, 0.22500000000000003, 0.9, 3, 1, 0},
    {G_CAPSULE, 3, {0, 0, 0}, {1, 0, 0, 0}, 0.040000000000000001, 0.25, 0.9, 3, 1, 0},
    {G_CAPSULE, 4, {-0.10000000000000001, 0, 0.10000000000000001}, FOOT_QUAT, 0.059999999999999998, 0.10000000000000001, 1.9, 3, 1, 0},
    {G_CAPSULE, 5, {0, 0, -0.22500000000000009}, {1, 0, 0, 0}, 0.050000000000000003, 0.22500000000000003, 0.9, 3, 1, 0},
    {G_CAPSULE, 6, {0, 0, 0}, {1, 0, 0, 0}, 0.040000000000000001, 0.25, 0.9, 3, 1, 0},
    {G_CAPSULE, 7, {-0.10000000000000001, 0, 0.10000000000000001}, FOOT_QUAT, 0.059999999999999998, 0.10000000000000001, 1.9, 3, 1, 0},
};
static const adef_t ACT_DEF[NU] = {{3, 100.0}, {4, 100.0}, {5, 100.0}, {6, 100.0}, {7, 100.0}, {8, 100.0}};
#define OPT_MARGIN 0.0
#define OPT_SOLIMP_CONTACT {0.9, 0.95, 0.001, 0.5, 2.0}
#endif

#ifndef OPT_TIMESTEP
#define OPT_TIMESTEP 0.002
#define OPT_CTRLRANGE 1.0
#define FRAME_SKIP 4
#endif

static void forward_position(const model_t* m, data_t* d);
static void solve_M(const model_t* m, const data_t* d, double* x);

static void build_model(model_t* m) {
  memset(m, 0, sizeof(*m));
  m->timestep = OPT_TIMESTEP; m->gravity[2] = -9.81; m->margin = OPT_MARGIN; m->tolerance = 1e-8; m->iterations = 100;
  m->solref[0] = 0.02; m->solref[1] = 1.0;
  m->solimp[0] = 0.9; m->solimp[1] = 0.95; m->solimp[2] = 0.001; m->solimp[3] = 0.5; m->solimp[4] = 2.0;
  {
    const double sc[5] = OPT_SOLIMP_CONTACT;
    memcpy(m->solimp_contact, sc, sizeof(sc));
#if defined(OPT_SOLIMP_LIMIT)
    const double sl[5] = OPT_SOLIMP_LIMIT;
    memcpy(m->solimp, sl, sizeof(sl));
#endif
    /* getsolparam (engine_core_constraint.c): dmin and dmax are clipped to [mjMINIMP, mjMAXIMP] = [0.0001, 0.9999] */
    for (int k = 0; k < 2; ++k) {
      m->solimp[k] = m->solimp[k] < 0.0001 ? 0.0001 : (m->solimp[k] > 0.9999 ? 0.9999 : m->solimp[k]);
      m->solimp_contact[k] = m->solimp_contact[k] < 0.0001 ? 0.0001 : (m->solimp_contact[k] > 0.9999 ? 0.9999 : m->solimp_contact[k]);
    }
  }
  for (int b = 0; b < NB; ++b) {
    m->parent[b] = BODY_DEF[b].parent;
    cp3(m->body_pos[b], BODY_DEF[b].pos);
    m->body_quat[b][0] = 1;
    m->body_jntadr[b] = -1; m->body_dofadr[b] = -1;
  }
  for (int j = 0; j < NJ; ++j) { /* one dof and one qpos entry per joint */
    const jdef_t* J = &JOINT_DEF[j];
    m->jnt_type[j] = J->type; m->jnt_body[j] = J->body; m->jnt_qposadr[j] = j; m->jnt_dofadr[j] = j;
    cp3(m->jnt_pos[j], J->pos);
    cp3(m->jnt_axis[j], J->axis);
    normalize3(m->jnt_axis[j]);
    m->jnt_limited[j] = J->limited;
#if defined(OPT_ANGLE_RADIAN)
    const double unit = 1.0;                      /* compiler angle="radian" */
#else
    const double unit = J->type == 3 ? DEG : 1.0; /* compiler angle="degree" applies to hinges only */
#endif
    m->jnt_range[j][0] = J->lo * unit; m->jnt_range[j][1] = J->hi * unit;
#if defined(ROBOT_HALFCHEETAH)
    m->jnt_stiffness[j] = JOINT_STIFFNESS[j];
#else
    m->jnt_stiffness[j] = 0.0;
#endif
    m->dof_armature[j] = J->armature; m->dof_damping[j] = J->damping;
    m->qpos0[j] = J->ref;
  }
  for (int j = 0; j < NJ; ++j) {
    int b = m->jnt_body[j];
    if (m->body_jntadr[b] < 0) { m->body_jntadr[b] = j; m->body_dofadr[b] = m->jnt_dofadr[j]; }
    m->body_jntnum[b] += 1;
    m->body_dofnum[b] += 1;
  }
  for (int i = 0; i < NV; ++i) { m->dof_jnt[i] = i; m->dof_body[i] = m->jnt_body[i]; }
  m->body_lastdof[0] = -1;
  for (int b = 1; b < NB; ++b)
    m->body_lastdof[b] = m->body_dofnum[b] ? m->body_dofadr[b] + m->body_dofnum[b] - 1 : m->body_lastdof[m->parent[b]];
  for (int i = 0; i < NV; ++i) {
    int b = m->dof_body[i];
    m->dof_parent[i] = i > m->body_dofadr[b] ? i - 1 : m->body_lastdof[m->parent[b]];
  }
  /* geoms + inertia from geoms (density 1000) */
  double bm[NB] = {0}, bcom[NB][3] = {{0}};
  double gmass[NG], gI[NG][9];
  for (int g = 0; g < NG; ++g) {
    const gdef_t* G = &GEOM_DEF[g];
    m->geom_type[g] = G->type; m->geom_body[g] = G->body; m->geom_condim[g] = G->condim;
    m->geom_contype[g] = G->contype; m->geom_conaffinity[g] = G->conaffinity;
    m->geom_friction[g] = G->friction;
    double I[3] = {0, 0, 0}, q[4] = {G->quat[0], G->quat[1], G->quat[2], G->quat[3]}, gpos[3], ghalf = G->half;
    cp3(gpos, G->pos);
#if defined(ROBOT_INVPEND)
    if (g == 1) { /* capsule from `fromto` (user_objects.cc mjCGeom::Compile / mjuu_z2quat): centre = midpoint, half length =
                     |to - from| / 2, frame = the rotation taking z onto the segment about z x segment */
      const double ft[6] = POLE_FROMTO;
      double vec[3] = {ft[3] - ft[0], ft[4] - ft[1], ft[5] - ft[2]}, z[3] = {0, 0, 1}, axis[3];
      for (int k = 0; k < 3; ++k) gpos[k] = 0.5 * (ft[k] + ft[3 + k]);
      ghalf = 0.5 * normalize3(vec);
      cross3(axis, z, vec);
      const double sn = norm3(axis);
      if (sn < 1e-10) { axis[0] = 1; axis[1] = 0; axis[2] = 0; } else { axis[0] /= sn; axis[1] /= sn; axis[2] /= sn; }
      const double ang = atan2(sn, vec[2]);
      q[0] = cos(0.5 * ang); q[1] = axis[0] * sin(0.5 * ang); q[2] = axis[1] * sin(0.5 * ang); q[3] = axis[2] * sin(0.5 * ang);
    }
#endif
#if defined(ROBOT_HALFCHEETAH)
    if (GEOM_SPEC[g].kind == 1) { /* axisangle="0 1 0 a" (radians): quat = (cos a/2, 0, sin a/2, 0) */
      const double ang = GEOM_SPEC[g].a[0];
      q[0] = cos(0.5 * ang); q[1] = 0; q[2] = sin(0.5 * ang); q[3] = 0;
    } else if (GEOM_SPEC[g].kind == 2) { /* fromto: as for the inverted pendulum's pole */
      const double* ft = GEOM_SPEC[g].a;
      double vec[3] = {ft[3] - ft[0], ft[4] - ft[1], ft[5] - ft[2]}, z[3] = {0, 0, 1}, axis[3];
      for (int k = 0; k < 3; ++k) gpos[k] = 0.5 * (ft[k] + ft[3 + k]);
      ghalf = 0.5 * normalize3(vec);
      cross3(axis, z, vec);
      const double sn = norm3(axis);
      if (sn < 1e-10) { axis[0] = 1; axis[1] = 0; axis[2] = 0; } else { axis[0] /= sn; axis[1] /= sn; axis[2] /= sn; }
      const double ang = atan2(sn, vec[2]);
      q[0] = cos(0.5 * ang); q[1] = axis[0] * sin(0.5 * ang); q[2] = axis[1] * sin(0.5 * ang); q[3] = axis[2] * sin(0.5 * ang);
    }
#endif
    quat_normalize(q);
    quat2mat(m->geom_mat[g], q);
    cp3(m->geom_pos[g], gpos);
    gmass[g] = 0;
    if (G->type == G_PLANE) {
      m->geom_rbound[g] = 0;
    } else {
      double r = G->r, half = ghalf, h = 2.0 * half;
      m->geom_size[g][0] = r; m->geom_size[g][1] = half;
      m->geom_rbound[g] = r + half;
      double ms = 1000.0 * (4.0 / 3.0) * PI * r * r * r, mc = 1000.0 * PI * r * r * h;
      gmass[g] = ms + mc;
      I[0] = I[1] = mc * (3 * r * r + h * h) / 12.0 + ms * (0.4 * r * r + 0.25 * h * h + 0.375 * r * h);
      I[2] = 0.5 * mc * r * r + 0.4 * ms * r * r;
    }
    const double* R = m->geom_mat[g];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        gI[g][3 * i + j] = R[3 * i] * I[0] * R[3 * j] + R[3 * i + 1] * I[1] * R[3 * j + 1] + R[3 * i + 2] * I[2] * R[3 * j + 2];
    int b = G->body;
    bm[b] += gmass[g];
    for (int k = 0; k < 3; ++k) bcom[b][k] += gmass[g] * m->geom_pos[g][k];
  }
  for (int b = 1; b < NB; ++b) {
    m->body_mass[b] = bm[b];
    for (int k = 0; k < 3; ++k) m->body_ipos[b][k] = bcom[b][k] / bm[b];
  }
  for (int g = 0; g < NG; ++g) { /* parallel-axis accumulation about the body com */
    int b = GEOM_DEF[g].body;
    if (b == 0) continue; /* world geoms (the floor) carry no inertia */
    double dv[3] = {m->geom_pos[g][0] - m->body_ipos[b][0], m->geom_pos[g][1] - m->body_ipos[b][1],
                    m->geom_pos[g][2] - m->body_ipos[b][2]};
    double d2 = dot3(dv, dv);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        m->body_inertia[b][3 * i + j] += gI[g][3 * i + j] + gmass[g] * ((i == j ? d2 : 0.0) - dv[i] * dv[j]);
  }
#if defined(OPT_TOTALMASS)
  { /* compiler settotalmass: every body's mass and inertia scaled so that the masses add up to it (user_model.cc) */
    double total = 0;
    for (int b = 1; b < NB; ++b) total += m->body_mass[b];
    const double scale = OPT_TOTALMASS / total;
    for (int b = 1; b < NB; ++b) {
      m->body_mass[b] *= scale;
      for (int k = 0; k < 9; ++k) m->body_inertia[b][k] *= scale;
    }
  }
#endif
  for (int b = NB - 1; b >= 0; --b) m->subtree_mass[b] = m->body_mass[b];
  for (int b = NB - 1; b >= 1; --b) m->subtree_mass[m->parent[b]] += m->subtree_mass[b];
  for (int u = 0; u < NU; ++u) {
    m->act_dof[u] = m->jnt_dofadr[ACT_DEF[u].joint]; m->act_gear[u] = ACT_DEF[u].gear;
    m->act_ctrlrange[u][0] = -OPT_CTRLRANGE; m->act_ctrlrange[u][1] = OPT_CTRLRANGE;
  }
  /* collision pairs: different bodies, not parent-child (unless the parent is the world), ordered by (body, body, geom) */
  m->npair = 0;
  for (int b1 = 0; b1 < NB; ++b1)
    for (int b2 = b1 + 1; b2 < NB; ++b2) {
      if (b1 != 0 && (m->parent[b2] == b1 || m->parent[b1] == b2)) continue;
      for (int g1 = 0; g1 < NG; ++g1)
        for (int g2 = 0; g2 < NG; ++g2) {
          if (m->geom_body[g1] != b1 || m->geom_body[g2] != b2) continue;
          if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
          int a = g1, c = g2;
          if (m->geom_type[a] > m->geom_type[c]) { int t = a; a = c; c = t; } /* type1 <= type2 */
          m->pair_g1[m->npair] = a; m->pair_g2[m->npair] = c; m->npair++;
        }
    }
  /* the constants derived at qpos0 (mj_setConst): invweight0, meaninertia */
  data_t* d = (data_t*)calloc(1, sizeof(data_t));
  memcpy(d->qpos, m->qpos0, sizeof(m->qpos0));
  forward_position(m, d);
  double Minv[NV][NV];
  for (int i = 0; i < NV; ++i) {
    double e[NV] = {0};
    e[i] = 1.0;
    solve_M(m, d, e);
    for (int j = 0; j < NV; ++j) Minv[j][i] = e[j];
  }
  double mi = 0;
  for (int i = 0; i < NV; ++i) mi += d->qM[i][i];
  m->meaninertia = mi / NV;
  for (int i = 0; i < NV; ++i) m->dof_invweight0[i] = Minv[i][i];
  for (int b = 1; b < NB; ++b) { /* body_invweight0 = trace(J Minv J^T)/3 for the com translational / rotational Jacobians */
    double Jp[3][NV] = {{0}}, Jr[3][NV] = {{0}};
    int i = m->body_lastdof[b];
    while (i >= 0) {
      double off[3] = {d->xipos[b][0] - d->subtree_com[1][0], d->xipos[b][1] - d->subtree_com[1][1],
                       d->xipos[b][2] - d->subtree_com[1][2]};
      double t[3];
      cross3(t, d->cdof[i], off);
      for (int k = 0; k < 3; ++k) { Jp[k][i] = d->cdof[i][3 + k] + t[k]; Jr[k][i] = d->cdof[i][k]; }
      i = m->dof_parent[i];
    }
    double tp = 0, tr = 0;
    for (int k = 0; k < 3; ++k)
      for (int a = 0; a < NV; ++a)
        for (int c = 0; c < NV; ++c) { tp += Jp[k][a] * Minv[a][c] * Jp[k][c]; tr += Jr[k][a] * Minv[a][c] * Jr[k][c]; }
    m->body_invweight0[b][0] = tp / 3; m->body_invweight0[b][1] = tr / 3;
  }
  free(d);
}

/* --- mj_kinematics + mj_comPos + tendon + mj_crb + mj_factorM ----------------------------------------------------------- */
static void kinematics(const model_t* m, data_t* d) {
  double* q = d->qpos;
  double id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  memset(d->xpos[0], 0, 24); d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  memcpy(d->xmat[0], id, sizeof(id)); memset(d->xipos[0], 0, 24);
  for (int b = 1; b < NB; ++b) {
    double xpos[3], xquat[4];
    int p = m->parent[b];
    double t[3];
    mulmatvec3(t, d->xmat[p], m->body_pos[b]);
    for (int k = 0; k < 3; ++k) xpos[k] = d->xpos[p][k] + t[k];
    quat_mul(xquat, d->xquat[p], m->body_quat[b]);
    for (int jj = 0; jj < m->body_jntnum[b]; ++jj) {
      int j = m->body_jntadr[b] + jj;
      double v[3];
      quat_rot(v, xquat, m->jnt_pos[j]);
      for (int k = 0; k < 3; ++k) d->xanchor[j][k] = xpos[k] + v[k];
      quat_rot(d->xaxis[j], xquat, m->jnt_axis[j]);
      const double disp = q[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
      if (m->jnt_type[j] == 2) { /* slide: translate along the axis */
        for (int k = 0; k < 3; ++k) xpos[k] += d->xaxis[j][k] * disp;
      } else { /* hinge: rotate about the axis through the anchor */
        double ql[4];
        quat_axisangle(ql, m->jnt_axis[j], disp);
        quat_mul(xquat, xquat, ql);
        quat_rot(v, xquat, m->jnt_pos[j]);
        for (int k = 0; k < 3; ++k) xpos[k] = d->xanchor[j][k] - v[k];
      }
    }
    quat_normalize(xquat);
    cp3(d->xpos[b], xpos);
    memcpy(d->xquat[b], xquat, 32);
    quat2mat(d->xmat[b], xquat);
    mulmatvec3(t, d->xmat[b], m->body_ipos[b]);
    for (int k = 0; k < 3; ++k) d->xipos[b][k] = xpos[k] + t[k];
  }
  for (int g = 0; g < NG; ++g) {
    int b = m->geom_body[g];
    double t[3];
    mulmatvec3(t, d->xmat[b], m->geom_pos[g]);
    for (int k = 0; k < 3; ++k) d->geom_xpos[g][k] = d->xpos[b][k] + t[k];
    mulmat3(d->geom_xmat[g], d->xmat[b], m->geom_mat[g]);
  }
}

static void com_pos(const model_t* m, data_t* d) {
  for (int b = 0; b < NB; ++b)
    for (int k = 0; k < 3; ++k) d->subtree_com[b][k] = m->body_mass[b] * d->xipos[b][k];
  for (int b = NB - 1; b >= 1; --b)
    for (int k = 0; k < 3; ++k) d->subtree_com[m->parent[b]][k] += d->subtree_com[b][k];
  for (int b = 0; b < NB; ++b) {
    if (m->subtree_mass[b] < MINVAL) cp3(d->subtree_com[b], d->xipos[b]);
    else for (int k = 0; k < 3; ++k) d->subtree_com[b][k] /= m->subtree_mass[b];
  }
  memset(d->cinert[0], 0, sizeof(d->cinert[0]));
  const double* root = d->subtree_com[1]; /* body_rootid = torso for every body */
  for (int b = 1; b < NB; ++b) {
    double off[3] = {d->xipos[b][0] - root[0], d->xipos[b][1] - root[1], d->xipos[b][2] - root[2]};
    const double* R = d->xmat[b];
    double RI[9], W[9], Rt[9];
    mulmat3(RI, R, m->body_inertia[b]);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rt[3 * i + j] = R[3 * j + i];
    mulmat3(W, RI, Rt);
    double mass = m->body_mass[b], *c = d->cinert[b];
    c[0] = W[0] + mass * (off[1] * off[1] + off[2] * off[2]);
    c[1] = W[4] + mass * (off[0] * off[0] + off[2] * off[2]);
    c[2] = W[8] + mass * (off[0] * off[0] + off[1] * off[1]);
    c[3] = W[1] - mass * off[0] * off[1];
    c[4] = W[2] - mass * off[0] * off[2];
    c[5] = W[5] - mass * off[1] * off[2];
    c[6] = mass * off[0]; c[7] = mass * off[1]; c[8] = mass * off[2]; c[9] = mass;
  }
  for (int j = 0; j < NJ; ++j) {
    double off[3] = {root[0] - d->xanchor[j][0], root[1] - d->xanchor[j][1], root[2] - d->xanchor[j][2]};
    int da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == 2) { /* slide: pure translation along the axis */
      d->cdof[da][0] = d->cdof[da][1] = d->cdof[da][2] = 0;
      cp3(d->cdof[da] + 3, d->xaxis[j]);
    } else {
      cp3(d->cdof[da], d->xaxis[j]);
      cross3(d->cdof[da] + 3, d->xaxis[j], off);
    }
  }
}

static void mul_inert_vec(double* res, const double* i, const double* v) {
  res[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  res[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  res[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  res[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  res[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  res[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static inline double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

static void crb_and_factor(const model_t* m, data_t* d) {
  double crb[NB][10];
  memcpy(crb, d->cinert, sizeof(crb));
  for (int b = NB - 1; b >= 1; --b)
    if (m->parent[b] > 0)
      for (int k = 0; k < 10; ++k) crb[m->parent[b]][k] += crb[b][k];
  memset(d->qM, 0, sizeof(d->qM));
  for (int i = 0; i < NV; ++i) {
    double buf[6];
    mul_inert_vec(buf, crb[m->dof_body[i]], d->cdof[i]);
    d->qM[i][i] = m->dof_armature[i] + dot6(d->cdof[i], buf);
    for (int j = m->dof_parent[i]; j >= 0; j = m->dof_parent[j]) { d->qM[i][j] = dot6(d->cdof[j], buf); d->qM[j][i] = d->qM[i][j]; }
  }
  /* L^T D L (mj_factorM): dense storage, tree sparsity via the ancestor chains */
  memcpy(d->qLD, d->qM, sizeof(d->qM));
  for (int k = NV - 1; k >= 0; --k) {
    for (int i = m->dof_parent[k]; i >= 0; i = m->dof_parent[i]) {
      double tmp = d->qLD[k][i] / d->qLD[k][k];
      for (int j = i; j >= 0; j = m->dof_parent[j]) d->qLD[i][j] -= d->qLD[k][j] * tmp;
      d->qLD[k][i] = tmp;
    }
    d->qLDiagInv[k] = 1.0 / d->qLD[k][k];
  }
}
/* x <- M^-1 x (mj_solveM) */
static void solve_M(const model_t* m, const data_t* d, double* x) {
  for (int i = NV - 1; i >= 0; --i)
    for (int j = m->dof_parent[i]; j >= 0; j = m->dof_parent[j]) x[j] -= d->qLD[i][j] * x[i];
  for (int i = 0; i < NV; ++i) x[i] *= d->qLDiagInv[i];
  for (int i = 0; i < NV; ++i)
    for (int j = m->dof_parent[i]; j >= 0; j = m->dof_parent[j]) x[i] -= d->qLD[i][j] * x[j];
}

/* --- collision --------------------------------------------------------------------------------------------------------- */
static void make_frame(double* f) { /* mju_makeFrame: f[0:3] unit normal; f[3:6] hint or ~0 */
  double* n = f; double* t1 = f + 3; double* t2 = f + 6;
  if (norm3(t1) < 0.5) {
    t1[0] = t1[1] = t1[2] = 0;
    if (n[1] < 0.5 && n[1] > -0.5) t1[1] = 1; else t1[2] = 1;
  }
  double dd = dot3(n, t1);
  for (int k = 0; k < 3; ++k) t1[k] -= dd * n[k];
  normalize3(t1);
  cross3(t2, n, t1);
}
static int add_contact(const model_t* m, data_t* d, int g1, int g2, double dist, const double* pos, const double* normal,
                       const double* hint) {
  if (dist >= m->margin || d->ncon >= MAXCON) return 0;
  contact_t* c = &d->con[d->ncon++];
  c->g1 = g1; c->g2 = g2; c->dist = dist;
  cp3(c->pos, pos);
  cp3(c->frame, normal);
  if (hint) cp3(c->frame + 3, hint); else c->frame[3] = c->frame[4] = c->frame[5] = 0;
  make_frame(c->frame);
  int cd = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
  c->dim = cd;
  c->mu = m->geom_friction[g1] > m->geom_friction[g2] ? m->geom_friction[g1] : m->geom_friction[g2];
  c->efc_adr = -1;
  return 1;
}
static void sphere_sphere(const model_t* m, data_t* d, int g1, int g2, const double* p1, double r1, const double* p2, double r2) {
  double n[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double len = norm3(n);
  double dist = len - r1 - r2;
  if (dist >= m->margin) return;
  if (len < MINVAL) { n[0] = 1; n[1] = 0; n[2] = 0; } else { n[0] /= len; n[1] /= len; n[2] /= len; }
  double pos[3];
  for (int k = 0; k < 3; ++k) pos[k] = p1[k] + n[k] * (r1 + 0.5 * dist);
  add_contact(m, d, g1, g2, dist, pos, n, NULL);
}
static void plane_sphere(const model_t* m, data_t* d, int g1, int g2, const double* c, double r, const double* hint) {
  const double* R = d->geom_xmat[g1];
  double n[3] = {R[2], R[5], R[8]};
  double df[3] = {c[0] - d->geom_xpos[g1][0], c[1] - d->geom_xpos[g1][1], c[2] - d->geom_xpos[g1][2]};
  double dist = dot3(df, n) - r;
  if (dist >= m->margin) return;
  double pos[3];
  for (int k = 0; k < 3; ++k) pos[k] = c[k] - n[k] * (r + 0.5 * dist);
  add_contact(m, d, g1, g2, dist, pos, n, hint);
}
static void collide_pair(const model_t* m, data_t* d, int g1, int g2) {
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  const double *p1 = d->geom_xpos[g1], *p2 = d->geom_xpos[g2];
  if (t1 != G_PLANE) { /* bounding-sphere prefilter */
    double df[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    double bound = m->geom_rbound[g1] + m->geom_rbound[g2] + m->margin;
    if (dot3(df, df) > bound * bound) return;
  }
  if (t1 == G_PLANE && t2 == G_SPHERE) plane_sphere(m, d, g1, g2, p2, m->geom_size[g2][0], NULL);
  else if (t1 == G_PLANE && t2 == G_CAPSULE) {
    const double* R = d->geom_xmat[g2];
    double axis[3] = {R[2], R[5], R[8]}, e1[3], e2[3], h = m->geom_size[g2][1];
    for (int k = 0; k < 3; ++k) { e1[k] = p2[k] + axis[k] * h; e2[k] = p2[k] - axis[k] * h; }
    plane_sphere(m, d, g1, g2, e1, m->geom_size[g2][0], axis);
    plane_sphere(m, d, g1, g2, e2, m->geom_size[g2][0], axis);
  } else if (t1 == G_SPHERE && t2 == G_SPHERE) sphere_sphere(m, d, g1, g2, p1, m->geom_size[g1][0], p2, m->geom_size[g2][0]);
  else if (t1 == G_SPHERE && t2 == G_CAPSULE) {
    const double* R = d->geom_xmat[g2];
    double axis[3] = {R[2], R[5], R[8]}, h = m->geom_size[g2][1];
    double df[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double x = dot3(axis, df);
    x = x > h ? h : (x < -h ? -h : x);
    double cp[3] = {p2[0] + axis[0] * x, p2[1] + axis[1] * x, p2[2] + axis[2] * x};
    sphere_sphere(m, d, g1, g2, p1, m->geom_size[g1][0], cp, m->geom_size[g2][0]);
  } else if (t1 == G_CAPSULE && t2 == G_CAPSULE) {
    const double *R1 = d->geom_xmat[g1], *R2 = d->geom_xmat[g2];
    double a1[3] = {R1[2], R1[5], R1[8]}, a2[3] = {R2[2], R2[5], R2[8]}, l1 = m->geom_size[g1][1], l2 = m->geom_size[g2][1];
    double df[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double ma = dot3(a1, a1), mb = -dot3(a1, a2), mc = dot3(a2, a2), u = -dot3(a1, df), v = dot3(a2, df);
    double det = ma * mc - mb * mb;
    double x1, x2;
    if (fabs(det) >= MINVAL) {
      x1 = (mc * u - mb * v) / det; x2 = (ma * v - mb * u) / det;
      if (x1 > l1) { x1 = l1; x2 = (v - mb * l1) / mc; }
      else if (x1 < -l1) { x1 = -l1; x2 = (v + mb * l1) / mc; }
      if (x2 > l2) { x2 = l2; x1 = (u - mb * l2) / ma; if (x1 > l1) x1 = l1; else if (x1 < -l1) x1 = -l1; }
      else if (x2 < -l2) { x2 = -l2; x1 = (u + mb * l2) / ma; if (x1 > l1) x1 = l1; else if (x1 < -l1) x1 = -l1; }
      double c1[3], c2[3];
      for (int k = 0; k < 3; ++k) { c1[k] = p1[k] + a1[k] * x1; c2[k] = p2[k] + a2[k] * x2; }
      sphere_sphere(m, d, g1, g2, c1, m->geom_size[g1][0], c2, m->geom_size[g2][0]);
    } else { /* parallel axes: test both ends of segment 1 against segment 2 */
      for (int s = -1; s <= 1; s += 2) {
        double c1[3], c2[3];
        for (int k = 0; k < 3; ++k) c1[k] = p1[k] + a1[k] * (s * l1);
        double dd[3] = {c1[0] - p2[0], c1[1] - p2[1], c1[2] - p2[2]};
        double x = dot3(a2, dd);
        x = x > l2 ? l2 : (x < -l2 ? -l2 : x);
        for (int k = 0; k < 3; ++k) c2[k] = p2[k] + a2[k] * x;
        sphere_sphere(m, d, g1, g2, c1, m->geom_size[g1][0], c2, m->geom_size[g2][0]);
      }
    }
  }
}
static void collision(const model_t* m, data_t* d) {
  d->ncon = 0;
  for (int p = 0; p < m->npair; ++p) collide_pair(m, d, m->pair_g1[p], m->pair_g2[p]);
}

/* --- constraints --------------------------------------------------------------------------------------------------------- */
/* translational Jacobian of a world point attached to body b (mj_jac) */
static void jac_point(const model_t* m, const data_t* d, int b, const double* point, double J[3][NV]) {
  memset(J, 0, 3 * NV * sizeof(double));
  double off[3] = {point[0] - d->subtree_com[1][0], point[1] - d->subtree_com[1][1], point[2] - d->subtree_com[1][2]};
  for (int i = m->body_lastdof[b]; i >= 0; i = m->dof_parent[i]) {
    double t[3];
    cross3(t, d->cdof[i], off);
    for (int k = 0; k < 3; ++k) J[k][i] = d->cdof[i][3 + k] + t[k];
  }
}
static double impedance(const double* solimp, double pos, double margin) {
  double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3];
  double x = (pos - margin) / width;
  if (x < 0) x = -x;
  if (x >= 1.0) return dmax;
  if (x <= 0.0) return dmin;
  double y; /* power = 2 */
  if (x <= mid) { double a = 1.0 / mid; y = a * x * x; }
  else { double b = 1.0 / (1.0 - mid); y = 1.0 - b * (1.0 - x) * (1.0 - x); }
  return dmin + y * (dmax - dmin);
}
static void make_constraint(const model_t* m, data_t* d) {
  int n = 0;
  for (int j = 0; j < NJ; ++j) { /* joint limits: lower side then upper side */
    if (!m->jnt_limited[j]) continue;
    double value = d->qpos[m->jnt_qposadr[j]];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[j][(side + 1) / 2] - value);
      if (dist < 0.0 && n < MAXEFC) {
        memset(d->efc_J[n], 0, sizeof(d->efc_J[n]));
        d->efc_J[n][m->jnt_dofadr[j]] = -side;
        d->efc_pos[n] = dist; d->efc_margin[n] = 0.0;
        d->efc_diagApprox[n] = m->dof_invweight0[m->jnt_dofadr[j]];
        d->efc_contact[n] = 0;
        ++n;
      }
    }
  }
  for (int c = 0; c < d->ncon; ++c) {
    contact_t* con = &d->con[c];
    int b1 = m->geom_body[con->g1], b2 = m->geom_body[con->g2];
    double J1[3][NV], J2[3][NV], Jc[3][NV];
    jac_point(m, d, b1, con->pos, J1);
    jac_point(m, d, b2, con->pos, J2);
    for (int r = 0; r < 3; ++r)
      for (int i = 0; i < NV; ++i) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += con->frame[3 * r + k] * (J2[k][i] - J1[k][i]);
        Jc[r][i] = s;
      }
    double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    int rows = con->dim == 1 ? 1 : 4;
    if (n + rows > MAXEFC) { con->efc_adr = -1; continue; }
    con->efc_adr = n;
    if (con->dim == 1) {
      memcpy(d->efc_J[n], Jc[0], sizeof(Jc[0]));
      d->efc_pos[n] = con->dist; d->efc_margin[n] = m->margin; d->efc_diagApprox[n] = tran;
      d->efc_contact[n] = 1;
      ++n;
    } else { /* pyramidal: normal +- mu * tangent_k */
      for (int k = 1; k <= 2; ++k)
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          for (int i = 0; i < NV; ++i) d->efc_J[n][i] = Jc[0][i] + sgn * con->mu * Jc[k][i];
          d->efc_pos[n] = con->dist; d->efc_margin[n] = m->margin;
          d->efc_diagApprox[n] = tran + con->mu * con->mu * tran;
          d->efc_contact[n] = 1;
          ++n;
        }
    }
  }
  d->nefc = n;
  /* impedance -> R, D; reference acceleration */
  const double timeconst = m->solref[0] > 2 * m->timestep ? m->solref[0] : 2 * m->timestep, dampratio = m->solref[1];
  for (int i = 0; i < n; ++i) {
    const double* solimp = d->efc_contact[i] ? m->solimp_contact : m->solimp;
    const double dmax = solimp[1];
    const double K = 1.0 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio), B = 2.0 / (dmax * timeconst);
    double imp = impedance(solimp, d->efc_pos[i], d->efc_margin[i]);
    double R = (1.0 - imp) / imp * d->efc_diagApprox[i];
    d->efc_R[i] = R < MINVAL ? MINVAL : R;
    double vel = 0;
    for (int k = 0; k < NV; ++k) vel += d->efc_J[i][k] * d->qvel[k];
    d->efc_vel[i] = vel;
    d->efc_aref[i] = -B * vel - K * imp * (d->efc_pos[i] - d->efc_margin[i]);
  }
  for (int c = 0; c < d->ncon; ++c) { /* pyramidal rows share Rpy = 2 mu^2 R(first row) */
    contact_t* con = &d->con[c];
    if (con->efc_adr >= 0 && con->dim > 1) {
      double Rpy = 2.0 * con->mu * con->mu * d->efc_R[con->efc_adr];
      for (int k = 0; k < 4; ++k) d->efc_R[con->efc_adr + k] = Rpy;
    }
  }
  for (int i = 0; i < n; ++i) d->efc_D[i] = 1.0 / d->efc_R[i];
  /* mj_projectConstraint: AR = J M^-1 J^T + diag(R) */
  double X[MAXEFC][NV];  /* on the stack: the library is called from several host threads */
  for (int i = 0; i < n; ++i) {
    memcpy(X[i], d->efc_J[i], sizeof(X[i]));
    solve_M(m, d, X[i]);
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = 0;
      for (int k = 0; k < NV; ++k) s += d->efc_J[j][k] * X[i][k];
      d->efc_AR[i][j] = d->efc_AR[j][i] = s;
    }
  for (int i = 0; i < n; ++i) d->efc_AR[i][i] += d->efc_R[i];
}

static void forward_position(const model_t* m, data_t* d) {
  kinematics(m, d);
  com_pos(m, d);
  crb_and_factor(m, d);
  collision(m, d);
  make_constraint(m, d);
}

/* --- velocity stage --------------------------------------------------------------------------------------------------------- */
static void cross_motion(double* r, const double* vel, const double* v) {
  double t[6];
  t[0] = -vel[2] * v[1] + vel[1] * v[2]; t[1] = vel[2] * v[0] - vel[0] * v[2]; t[2] = -vel[1] * v[0] + vel[0] * v[1];
  t[3] = -vel[2] * v[4] + vel[1] * v[5]; t[4] = vel[2] * v[3] - vel[0] * v[5]; t[5] = -vel[1] * v[3] + vel[0] * v[4];
  t[3] += -vel[5] * v[1] + vel[4] * v[2]; t[4] += vel[5] * v[0] - vel[3] * v[2]; t[5] += -vel[4] * v[0] + vel[3] * v[1];
  memcpy(r, t, sizeof(t));
}
static void cross_force(double* r, const double* vel, const double* f) {
  double t[6];
  t[0] = -vel[2] * f[1] + vel[1] * f[2]; t[1] = vel[2] * f[0] - vel[0] * f[2]; t[2] = -vel[1] * f[0] + vel[0] * f[1];
  t[3] = -vel[2] * f[4] + vel[1] * f[5]; t[4] = vel[2] * f[3] - vel[0] * f[5]; t[5] = -vel[1] * f[3] + vel[0] * f[4];
  t[0] += -vel[5] * f[4] + vel[4] * f[5]; t[1] += vel[5] * f[3] - vel[3] * f[5]; t[2] += -vel[4] * f[3] + vel[3] * f[4];
  memcpy(r, t, sizeof(t));
}
static void com_vel(const model_t* m, data_t* d) {
  memset(d->cvel[0], 0, 48);
  for (int b = 1; b < NB; ++b) {
    double cvel[6];
    memcpy(cvel, d->cvel[m->parent[b]], 48);
    for (int jj = 0; jj < m->body_jntnum[b]; ++jj) {
      int j = m->body_jntadr[b] + jj, da = m->jnt_dofadr[j];
      cross_motion(d->cdof_dot[da], cvel, d->cdof[da]); /* slide and hinge alike */
      for (int k = 0; k < 6; ++k) cvel[k] += d->cdof[da][k] * d->qvel[da];
    }
    memcpy(d->cvel[b], cvel, 48);
  }
}
/* mj_rne without accelerations -> qfrc_bias */
static void rne_bias(const model_t* m, data_t* d) {
  double cacc[NB][6], cfrc[NB][6];
  memset(cacc[0], 0, 48);
  cacc[0][3] = -m->gravity[0]; cacc[0][4] = -m->gravity[1]; cacc[0][5] = -m->gravity[2];
  for (int b = 1; b < NB; ++b) {
    memcpy(cacc[b], cacc[m->parent[b]], 48);
    for (int i = 0; i < m->body_dofnum[b]; ++i) {
      int da = m->body_dofadr[b] + i;
      for (int k = 0; k < 6; ++k) cacc[b][k] += d->cdof_dot[da][k] * d->qvel[da];
    }
    double t1[6], t2[6], t3[6];
    mul_inert_vec(t1, d->cinert[b], cacc[b]);
    mul_inert_vec(t2, d->cinert[b], d->cvel[b]);
    cross_force(t3, d->cvel[b], t2);
    for (int k = 0; k < 6; ++k) cfrc[b][k] = t1[k] + t3[k];
  }
  memset(cfrc[0], 0, 48);
  for (int b = NB - 1; b >= 1; --b)
    for (int k = 0; k < 6; ++k) cfrc[m->parent[b]][k] += cfrc[b][k];
  for (int i = 0; i < NV; ++i) d->qfrc_bias[i] = dot6(d->cdof[i], cfrc[m->dof_body[i]]);
}
static void forward_velocity(const model_t* m, data_t* d) {
  com_vel(m, d);
  for (int i = 0; i < NV; ++i) d->qfrc_passive[i] = 0;
  for (int j = 0; j < NJ; ++j) {
    int da = m->jnt_dofadr[j], qa = m->jnt_qposadr[j];
    d->qfrc_passive[da] = -m->jnt_stiffness[j] * (d->qpos[qa] - m->qpos0[qa]) - m->dof_damping[da] * d->qvel[da];
  }
  rne_bias(m, d);
}
static void forward_actuation_acceleration(const model_t* m, data_t* d) {
  for (int i = 0; i < NV; ++i) d->qfrc_actuator[i] = 0;
  for (int u = 0; u < NU; ++u) {
    double c = d->ctrl[u];
    c = c < m->act_ctrlrange[u][0] ? m->act_ctrlrange[u][0] : (c > m->act_ctrlrange[u][1] ? m->act_ctrlrange[u][1] : c);
    d->qfrc_actuator[m->act_dof[u]] += m->act_gear[u] * c;
  }
  for (int i = 0; i < NV; ++i) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  memcpy(d->qacc_smooth, d->qfrc_smooth, sizeof(d->qacc_smooth));
  solve_M(m, d, d->qacc_smooth);
}
/* mj_fwdConstraint with the PGS solver */
static void forward_constraint(const model_t* m, data_t* d) {
  int n = d->nefc;
  memset(d->qfrc_constraint, 0, sizeof(d->qfrc_constraint));
  d->solver_iter = 0;
  if (n == 0) { memcpy(d->qacc, d->qacc_smooth, sizeof(d->qacc)); return; }
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int k = 0; k < NV; ++k) s += d->efc_J[i][k] * d->qacc_smooth[k];
    d->efc_b[i] = s - d->efc_aref[i];
  }
  /* warm start: forces implied by qacc_warmstart, kept only if their dual cost beats zero force */
  double cost = 0;
  for (int i = 0; i < n; ++i) {
    double jar = 0;
    for (int k = 0; k < NV; ++k) jar += d->efc_J[i][k] * d->qacc_warmstart[k];
    jar -= d->efc_aref[i];
    d->efc_force[i] = jar < 0 ? -d->efc_D[i] * jar : 0.0;
  }
  double res[MAXEFC], ainv[MAXEFC];  /* running residual AR f + b of every row, 1 / AR[i][i] */
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int j = 0; j < n; ++j) s += d->efc_AR[i][j] * d->efc_force[j];
    cost += d->efc_force[i] * (0.5 * s + d->efc_b[i]);
    res[i] = d->efc_b[i] + s;
    ainv[i] = 1.0 / d->efc_AR[i][i];
  }
  if (cost > 0)
    for (int i = 0; i < n; ++i) { d->efc_force[i] = 0; res[i] = d->efc_b[i]; }
  const double scale = 1.0 / (m->meaninertia * (NV > 1 ? NV : 1));
  /* Gauss-Seidel sweeps in residual-update form: after row j moves by delta, every residual moves by AR[:, j] * delta
     (same sweep as recomputing AR[j, :] f + b per row; the update form is what maps onto one lane per row) */
  for (int it = 0; it < m->iterations; ++it) {
    double improvement = 0;
    for (int j = 0; j < n; ++j) {
      const double old = d->efc_force[j], r = res[j];
      double f = old - r * ainv[j];
      if (f < 0) f = 0;
      d->efc_force[j] = f;
      const double delta = f - old;
      improvement -= 0.5 * delta * delta * d->efc_AR[j][j] + delta * r;
      for (int i = 0; i < n; ++i) res[i] += d->efc_AR[i][j] * delta;
    }
    d->solver_iter = it + 1;
    if (improvement * scale < m->tolerance) break;
  }
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < NV; ++k) d->qfrc_constraint[k] += d->efc_J[i][k] * d->efc_force[i];
  double t[NV];
  memcpy(t, d->qfrc_constraint, sizeof(t));
  solve_M(m, d, t);
  for (int k = 0; k < NV; ++k) d->qacc[k] = d->qacc_smooth[k] + t[k];
}
static void mj_forward(const model_t* m, data_t* d) {
  forward_position(m, d);
  forward_velocity(m, d);
  forward_actuation_acceleration(m, d);
  forward_constraint(m, d);
}

/* --- integration --------------------------------------------------------------------------------------------------------- */
static void integrate_pos(double* qpos, const double* vel, double h) {
  for (int i = 0; i < NV; ++i) qpos[i] += h * vel[i]; /* slide and hinge joints only */
}
#if !defined(OPT_EULER)
static void mj_step_rk4(const model_t* m, data_t* d) {
  mj_forward(m, d);
  const double h = m->timestep;
  static const double A[3][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 1.0}}, Bw[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6},
                      C[3] = {0.5, 0.5, 1.0};
  double X0q[NQ], X0v[NV], Xv[4][NV], F[4][NV], time0 = d->time;
  memcpy(X0q, d->qpos, sizeof(X0q)); memcpy(X0v, d->qvel, sizeof(X0v));
  memcpy(Xv[0], d->qvel, sizeof(X0v)); memcpy(F[0], d->qacc, sizeof(X0v));
  for (int i = 1; i < 4; ++i) {
    double dXv[NV] = {0}, dXa[NV] = {0};
    for (int j = 0; j < i; ++j)
      for (int k = 0; k < NV; ++k) { dXv[k] += A[i - 1][j] * Xv[j][k]; dXa[k] += A[i - 1][j] * F[j][k]; }
    memcpy(d->qpos, X0q, sizeof(X0q));
    integrate_pos(d->qpos, dXv, h);
    for (int k = 0; k < NV; ++k) d->qvel[k] = X0v[k] + h * dXa[k];
    d->time = time0 + C[i - 1] * h;
    memcpy(Xv[i], d->qvel, sizeof(X0v));
    mj_forward(m, d);
    memcpy(F[i], d->qacc, sizeof(X0v));
  }
  double dXv[NV] = {0}, dXa[NV] = {0};
  for (int j = 0; j < 4; ++j)
    for (int k = 0; k < NV; ++k) { dXv[k] += Bw[j] * Xv[j][k]; dXa[k] += Bw[j] * F[j][k]; }
  memcpy(d->qpos, X0q, sizeof(X0q));
  for (int k = 0; k < NV; ++k) d->qvel[k] = X0v[k] + h * dXa[k];
  integrate_pos(d->qpos, dXv, h);
  d->time = time0 + h;
  memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc)); /* mj_advance */
}
#endif
#if defined(OPT_EULER)
/* mj_Euler (engine_forward.c: mj_EulerSkip + mj_advance): semi-implicit Euler with the joint damping treated implicitly --
 * qacc' = (M + h diag(B))^-1 (qfrc_smooth + qfrc_constraint); qvel += h qacc'; qpos += h qvel (the NEW velocity);
 * qacc_warmstart = the forward dynamics' qacc */
static void mj_step_euler(const model_t* m, data_t* d) {
  const double h = m->timestep;
  mj_forward(m, d);
  double qacc[NV];
  int damped = 0;
  for (int i = 0; i < NV; ++i) damped = damped || m->dof_damping[i] > 0;
  if (!damped) {
    memcpy(qacc, d->qacc, sizeof(qacc));
  } else {
    double H[NV][NV], Hinv[NV];
    memcpy(H, d->qM, sizeof(H));
    for (int i = 0; i < NV; ++i) H[i][i] += h * m->dof_damping[i];
    for (int k = NV - 1; k >= 0; --k) { /* mj_factorI: the same L^T D L as crb_and_factor */
      for (int i = m->dof_parent[k]; i >= 0; i = m->dof_parent[i]) {
        double tmp = H[k][i] / H[k][k];
        for (int j = i; j >= 0; j = m->dof_parent[j]) H[i][j] -= H[k][j] * tmp;
        H[k][i] = tmp;
      }
      Hinv[k] = 1.0 / H[k][k];
    }
    for (int i = 0; i < NV; ++i) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    for (int i = NV - 1; i >= 0; --i) /* mj_solveLD */
      for (int j = m->dof_parent[i]; j >= 0; j = m->dof_parent[j]) qacc[j] -= H[i][j] * qacc[i];
    for (int i = 0; i < NV; ++i) qacc[i] *= Hinv[i];
    for (int i = 0; i < NV; ++i)
      for (int j = m->dof_parent[i]; j >= 0; j = m->dof_parent[j]) qacc[i] -= H[i][j] * qacc[j];
  }
  for (int k = 0; k < NV; ++k) d->qvel[k] += h * qacc[k];
  integrate_pos(d->qpos, d->qvel, h);
  d->time += h;
  memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc)); /* mj_advance */
}
#define MJ_STEP mj_step_euler
#else
#define MJ_STEP mj_step_rk4
#endif
/* ------------------------------------------------------------------------------------------------------------------ */
/* environment (hopper_v5.py / walker2d_v5.py); info rows: x_position, z_distance_from_origin, x_velocity, reward_forward,
 * reward_ctrl, reward_survive */
#if defined(ROBOT_INVPEND)
#define OBS (NQ + NV) /* inverted_pendulum_v5.py:185-186: qpos | qvel, nothing skipped or clipped */
#else
#define OBS (NQ - 1 + NV)
#endif
#define NINFO 6 /* InvertedPendulum uses row 5 (reward_survive) only */
typedef struct {
  data_t d;
  pcg64_t rng;
} henv_t;
typedef struct {
  int n, max_episode_steps;
  double reset_noise_scale;
  model_t model;
  henv_t* env;
  int *elapsed, *autoreset;
} pl_vec_t;

#if defined(ROBOT_INVPEND)
static void get_obs(const data_t* d, double* obs) { /* inverted_pendulum_v5.py:185-186 */
  for (int i = 0; i < NQ; ++i) obs[i] = d->qpos[i];
  for (int i = 0; i < NV; ++i) obs[NQ + i] = d->qvel[i];
}
static void env_reset(pl_vec_t* v, int i, double* obs, double* info) { /* inverted_pendulum_v5.py:168-183 */
  const model_t* m = &v->model;
  henv_t* e = &v->env[i];
  pcg64_t rng = e->rng;
  memset(&e->d, 0, sizeof(e->d)); /* mj_resetData */
  e->rng = rng;
  data_t* d = &e->d;
  const double c = v->reset_noise_scale;
  for (int k = 0; k < NQ; ++k) d->qpos[k] = m->qpos0[k] + (-c + (c - -c) * pcg64_double(&e->rng));
  for (int k = 0; k < NV; ++k) d->qvel[k] = 0.0 + (-c + (c - -c) * pcg64_double(&e->rng));
  mj_forward(m, d); /* set_state */
  get_obs(d, obs);
  memset(info, 0, NINFO * sizeof(double)); /* MujocoEnv._get_reset_info: {} */
}
static void env_step(pl_vec_t* v, int i, const float* action, double* obs, double* reward, int* terminated, double* info) {
  const model_t* m = &v->model;
  data_t* d = &v->env[i].d;
  for (int u = 0; u < NU; ++u) d->ctrl[u] = (double)action[u];
  for (int k = 0; k < FRAME_SKIP; ++k) MJ_STEP(m, d);
  get_obs(d, obs);
  int finite = 1;
  for (int k = 0; k < OBS; ++k) finite = finite && isfinite(obs[k]);
  *terminated = !finite || fabs(obs[1]) > 0.2; /* inverted_pendulum_v5.py:152-156 */
  *reward = *terminated ? 0.0 : 1.0;            /* int(not terminated) */
  memset(info, 0, NINFO * sizeof(double));
  info[5] = *reward;                            /* info["reward_survive"] */
}
#elif defined(ROBOT_HALFCHEETAH)
#include "ziggurat_tables.h"
/* exp and log from fixed sequences of IEEE operations (the CUDA engine runs the same ones), for the two slow paths of the
 * ziggurat: accurate to ~2e-16, so a wedge decision differs from libm's only when the two sides agree to 1e-15, and a tail
 * sample (0.03 % of the draws) can differ from numpy's by one ulp. */
static inline double det_exp(double y) { /* y <= 0 here */
  const double k = rint(y * 1.44269504088896338700e+00);
  const double r = (y - k * 6.93147180369123816490e-01) - k * 1.90821492927058770002e-10;
  double p = 1.0 / 6227020800.0;
  p = p * r + 1.0 / 479001600.0;
  p = p * r + 1.0 / 39916800.0;
  p = p * r + 1.0 / 3628800.0;
  p = p * r + 1.0 / 362880.0;
  p = p * r + 1.0 / 40320.0;
  p = p * r + 1.0 / 5040.0;
  p = p * r + 1.0 / 720.0;
  p = p * r + 1.0 / 120.0;
  p = p * r + 1.0 / 24.0;
  p = p * r + 1.0 / 6.0;
  p = p * r + 0.5;
  p = p * r + 1.0;
  p = p * r + 1.0;
  const int ik = (int)k;
  return ik >= 0 ? p * (double)(1ull << ik) : (ik > -63 ? p / (double)(1ull << -ik) : 0.0);
}
static inline double det_log(double y) { /* y in (0, 1] */
  uint64_t bits;
  memcpy(&bits, &y, 8);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  bits = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
  double mant;
  memcpy(&mant, &bits, 8); /* [1, 2) */
  if (mant > 1.41421356237309514547) { mant *= 0.5; e += 1; }
  const double s = (mant - 1.0) / (mant + 1.0), z = s * s;
  double p = 1.0 / 23.0;
  p = p * z + 1.0 / 21.0;
  p = p * z + 1.0 / 19.0;
  p = p * z + 1.0 / 17.0;
  p = p * z + 1.0 / 15.0;
  p = p * z + 1.0 / 13.0;
  p = p * z + 1.0 / 11.0;
  p = p * z + 1.0 / 9.0;
  p = p * z + 1.0 / 7.0;
  p = p * z + 1.0 / 5.0;
  p = p * z + 1.0 / 3.0;
  p = p * z + 1.0;
  return (double)e * 6.93147180369123816490e-01 + ((double)e * 1.90821492927058770002e-10 + 2.0 * s * p);
}
/* Generator.standard_normal: numpy/random/src/distributions/distributions.c random_standard_normal (ziggurat, 256 strips);
 * pinned against numpy in oracle/np_rng.py (PCG64.standard_normal) */
static uint64_t pcg64_u64(pcg64_t* g) {
  const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
  g->state = g->state * mult + g->inc;
  uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state, x = hi ^ lo;
  unsigned rot = (unsigned)(hi >> 58);
  return (x >> rot) | (x << ((64 - rot) & 63));
}
static double pcg64_standard_normal(pcg64_t* g) {
  const double nor_r = 3.6541528853610087963519472518, nor_inv_r = 0.27366123732975827203338247596;
  for (;;) {
    uint64_t r = pcg64_u64(g);
    const int idx = (int)(r & 0xff);
    r >>= 8;
    const int sign = (int)(r & 0x1);
    const uint64_t rabs = (r >> 1) & 0x000fffffffffffffull;
    double x = (double)rabs * zig_wi[idx];
    if (sign) x = -x;
    if (rabs < zig_ki[idx]) return x;
    if (idx == 0) {
      for (;;) {
        const double xx = -nor_inv_r * det_log(1.0 - pcg64_double(g));
        const double yy = -det_log(1.0 - pcg64_double(g));
        if (yy + yy > xx * xx) return ((rabs >> 8) & 0x1) ? -(nor_r + xx) : nor_r + xx;
      }
    } else if ((zig_fi[idx - 1] - zig_fi[idx]) * pcg64_double(g) + zig_fi[idx] < det_exp(-0.5 * x * x)) {
      return x;
    }
  }
}
static void get_obs(const data_t* d, double* obs) { /* half_cheetah_v5.py:251-259: qpos[1:] | qvel, nothing clipped */
  int o = 0;
  for (int i = 1; i < NQ; ++i) obs[o++] = d->qpos[i];
  for (int i = 0; i < NV; ++i) obs[o++] = d->qvel[i];
}
static void env_reset(pl_vec_t* v, int i, double* obs, double* info) { /* half_cheetah_v5.py:261-276 */
  const model_t* m = &v->model;
  henv_t* e = &v->env[i];
  pcg64_t rng = e->rng;
  memset(&e->d, 0, sizeof(e->d)); /* mj_resetData */
  e->rng = rng;
  data_t* d = &e->d;
  const double c = v->reset_noise_scale;
  for (int k = 0; k < NQ; ++k) d->qpos[k] = m->qpos0[k] + (-c + (c - -c) * pcg64_double(&e->rng));
  for (int k = 0; k < NV; ++k) d->qvel[k] = 0.0 + c * pcg64_standard_normal(&e->rng);
  mj_forward(m, d); /* set_state */
  get_obs(d, obs);
  memset(info, 0, NINFO * sizeof(double));
  info[0] = d->qpos[0]; /* _get_reset_info: x_position */
}
static void env_step(pl_vec_t* v, int i, const float* action, double* obs, double* reward, int* terminated, double* info) {
  const model_t* m = &v->model;
  data_t* d = &v->env[i].d;
  const double x_before = d->qpos[0];
  for (int u = 0; u < NU; ++u) d->ctrl[u] = (double)action[u];
  for (int k = 0; k < FRAME_SKIP; ++k) MJ_STEP(m, d);
  const double x_after = d->qpos[0];
  const double dt = m->timestep * FRAME_SKIP;
  const double xv = (x_after - x_before) / dt;
  get_obs(d, obs);
  const double forward_reward = 1.0 * xv;
  float sq = 0.0f; /* control_cost squares and sums the float32 action; NumPy 2 keeps weight * sum in float32 (NEP 50) */
  for (int u = 0; u < NU; ++u) sq += action[u] * action[u];
  const float ctrl_cost32 = (float)0.1 * sq;
  const double ctrl_cost = (double)ctrl_cost32;
  *reward = forward_reward - ctrl_cost; /* half_cheetah_v5.py:239-243 */
  *terminated = 0;                      /* never terminates */
  memset(info, 0, NINFO * sizeof(double));
  info[0] = x_after; info[2] = xv; info[3] = forward_reward; info[4] = -ctrl_cost;
}
#else
static void get_obs(const data_t* d, double* obs) { /* hopper_v5.py:253-261 */
  int o = 0;
  for (int i = 1; i < NQ; ++i) obs[o++] = d->qpos[i];
  for (int i = 0; i < NV; ++i) { double v = d->qvel[i]; obs[o++] = v < -10.0 ? -10.0 : (v > 10.0 ? 10.0 : v); }
}
static int is_healthy(const data_t* d) {
  const double z = d->qpos[1], angle = d->qpos[2];
#if defined(ROBOT_HOPPER) /* hopper_v5.py:231-247: state = qpos[2:] + qvel in (-100, 100), z > 0.7, |angle| < 0.2 */
  int ok = 1;
  for (int i = 2; i < NQ; ++i) ok = ok && (-100.0 < d->qpos[i] && d->qpos[i] < 100.0);
  for (int i = 0; i < NV; ++i) ok = ok && (-100.0 < d->qvel[i] && d->qvel[i] < 100.0);
  return ok && (0.7 < z) && (-0.2 < angle && angle < 0.2);
#else /* walker2d_v5.py:261-272: 0.8 < z < 2.0, |angle| < 1 */
  return (0.8 < z && z < 2.0) && (-1.0 < angle && angle < 1.0);
#endif
}
static void env_reset(pl_vec_t* v, int i, double* obs, double* info) {
  const model_t* m = &v->model;
  henv_t* e = &v->env[i];
  pcg64_t rng = e->rng;
  memset(&e->d, 0, sizeof(e->d)); /* mj_resetData */
  e->rng = rng;
  data_t* d = &e->d;
  const double c = v->reset_noise_scale;
  for (int k = 0; k < NQ; ++k) d->qpos[k] = m->qpos0[k] + (-c + (c - -c) * pcg64_double(&e->rng));
  for (int k = 0; k < NV; ++k) d->qvel[k] = 0.0 + (-c + (c - -c) * pcg64_double(&e->rng));
  mj_forward(m, d); /* set_state */
  get_obs(d, obs);
  memset(info, 0, NINFO * sizeof(double));
  info[0] = d->qpos[0]; info[1] = d->qpos[1] - m->qpos0[1]; /* _get_reset_info, hopper_v5.py:339-343 */
}
static void env_step(pl_vec_t* v, int i, const float* action, double* obs, double* reward, int* terminated, double* info) {
  const model_t* m = &v->model;
  data_t* d = &v->env[i].d;
  const double x_before = d->qpos[0];
  for (int u = 0; u < NU; ++u) d->ctrl[u] = (double)action[u];
  for (int k = 0; k < FRAME_SKIP; ++k) MJ_STEP(m, d); /* frame_skip 4 */
  const double x_after = d->qpos[0];
  const double dt = m->timestep * 4;
  const double xv = (x_after - x_before) / dt;
  get_obs(d, obs);
  const int healthy = is_healthy(d);
  const double forward_reward = 1.0 * xv, healthy_reward = healthy * 1.0;
  /* control_cost (hopper_v5.py:226-228) squares and sums the ACTION as passed -- float32 -- and NumPy 2 keeps the product
   * with the Python float weight in float32 (NEP 50): one float32 rounding per operation, left to right */
  float sq = 0.0f;
  for (int u = 0; u < NU; ++u) sq += action[u] * action[u];
  const float ctrl_cost32 = (float)1e-3 * sq;
  const double ctrl_cost = (double)ctrl_cost32;
  *reward = (forward_reward + healthy_reward) - ctrl_cost; /* hopper_v5.py:305-312 */
  *terminated = !healthy;
  info[0] = x_after; info[1] = d->qpos[1] - m->qpos0[1]; info[2] = xv;
  info[3] = forward_reward; info[4] = -ctrl_cost; info[5] = healthy_reward;
}
#endif

pl_vec_t* API(create)(int n, int max_episode_steps, double reset_noise_scale) {
  pl_vec_t* v = (pl_vec_t*)calloc(1, sizeof(pl_vec_t));
  v->n = n; v->max_episode_steps = max_episode_steps; v->reset_noise_scale = reset_noise_scale;
  build_model(&v->model);
  v->env = (henv_t*)calloc((size_t)n, sizeof(henv_t));
  v->elapsed = (int*)calloc((size_t)n, sizeof(int));
  v->autoreset = (int*)calloc((size_t)n, sizeof(int));
  return v;
}
void API(destroy)(pl_vec_t* v) { if (v) { free(v->env); free(v->elapsed); free(v->autoreset); free(v); } }
void API(reset)(pl_vec_t* v, const uint64_t* seeds, const uint8_t* mask, double* obs, double* info) {
  for (int i = 0; i < v->n; ++i) {
    if (mask && !mask[i]) continue;
    if (seeds) pcg64_seed(&v->env[i].rng, seeds[i]);
    env_reset(v, i, obs + OBS * i, info + NINFO * i);
    v->elapsed[i] = 0; v->autoreset[i] = 0;
  }
}
void API(step)(pl_vec_t* v, const float* actions, double* obs, double* reward, uint8_t* terminated, uint8_t* truncated, double* info) {
  for (int i = 0; i < v->n; ++i) {
    if (v->autoreset[i]) {
      env_reset(v, i, obs + OBS * i, info + NINFO * i);
      reward[i] = 0; terminated[i] = 0; truncated[i] = 0; v->elapsed[i] = 0; v->autoreset[i] = 0;
      continue;
    }
    int term;
    env_step(v, i, actions + NU * i, obs + OBS * i, reward + i, &term, info + NINFO * i);
    terminated[i] = (uint8_t)term;
    v->elapsed[i] += 1;
    truncated[i] = v->max_episode_steps > 0 && v->elapsed[i] >= v->max_episode_steps;
    v->autoreset[i] = terminated[i] || truncated[i];
  }
}
/* introspection for tests */
void API(model_info)(const pl_vec_t* v, double* body_mass /*NB*/, double* misc /*8*/, double* invweight /*NB*2 + NV*/) {
  const model_t* m = &v->model;
  for (int b = 0; b < NB; ++b) body_mass[b] = m->body_mass[b];
  misc[0] = m->meaninertia; misc[1] = m->npair; misc[2] = m->subtree_mass[0];
  for (int b = 0; b < NB; ++b) { invweight[2 * b] = m->body_invweight0[b][0]; invweight[2 * b + 1] = m->body_invweight0[b][1]; }
  for (int i = 0; i < NV; ++i) invweight[2 * NB + i] = m->dof_invweight0[i];
}
void API(debug)(const pl_vec_t* v, int i, double* qpos, double* qvel, double* qacc, int* counts /* ncon, nefc, iters */,
              double* xipos /* NB*3 */) {
  const data_t* d = &v->env[i].d;
  memcpy(qpos, d->qpos, sizeof(d->qpos)); memcpy(qvel, d->qvel, sizeof(d->qvel)); memcpy(qacc, d->qacc, sizeof(d->qacc));
  counts[0] = d->ncon; counts[1] = d->nefc; counts[2] = d->solver_iter;
  memcpy(xipos, d->xipos, sizeof(d->xipos));
}
void API(set_state)(pl_vec_t* v, int i, const double* qpos, const double* qvel) {
  data_t* d = &v->env[i].d;
  memcpy(d->qpos, qpos, sizeof(d->qpos)); memcpy(d->qvel, qvel, sizeof(d->qvel));
  mj_forward(&v->model, d);
}
/* contact forces of the last forward evaluation: per contact the normal force (pyramidal cone: the sum of its four rows) and
 * the contact frame's normal; returns ncon */
int API(contact_forces)(const pl_vec_t* v, int i, double* normal_force /*MAXCON*/, double* normal /*MAXCON*3*/) {
  const data_t* d = &v->env[i].d;
  for (int c = 0; c < d->ncon; ++c) {
    const contact_t* con = &d->con[c];
    double f = 0;
    if (con->efc_adr >= 0) {
      if (con->dim == 1) f = d->efc_force[con->efc_adr];
      else for (int k = 0; k < 4; ++k) f += d->efc_force[con->efc_adr + k];
    }
    normal_force[c] = f;
    for (int k = 0; k < 3; ++k) normal[3 * c + k] = con->frame[k];
  }
  return d->ncon;
}
const char* API(body_name)(int b) { return BODY_NAMES[b]; }
