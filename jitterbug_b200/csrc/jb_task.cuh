// Task layer fused behind the physics: observation, reward, reset randomisation and the heuristic policies of the
// Jitterbug domain, per environment, on the device.
//   observations : jitterbug_dmc/jitterbug.py:673-763 with the Physics helpers :180-317 and the normalisation
//                  tables :324-372 (always normalised -- the reference never reads norm_obs)
//   rewards      : jitterbug_dmc/jitterbug.py:840-925; closed forms of dm_control.utils.rewards.tolerance
//   reset        : jitterbug_dmc/jitterbug.py:601-666 (distributions reproduced with Philox4x32-10 streams keyed by
//                  the *global* env id and the episode index, so results do not depend on how envs are sharded)
//   heuristics   : jitterbug_dmc/heuristic_policies.py:6-136 (operate on the normalised observation, App. C7)
#ifndef JB_TASK_CUH
#define JB_TASK_CUH

#include "jb_math.cuh"
#include "jb_packed.h"
#include <stdint.h>

#define JB_PI 3.14159265358979323846

JB_HD float jb_atan2(float y, float x) { return atan2f(y, x); }
JB_HD double jb_atan2(double y, double x) { return atan2(y, x); }
JB_HD float jb_exp(float x) { return expf(x); }
JB_HD double jb_exp(double x) { return exp(x); }
JB_HD float jb_cos(float x) { return cosf(x); }
JB_HD double jb_cos(double x) { return cos(x); }
JB_HD float jb_ceil(float x) { return ceilf(x); }
JB_HD double jb_ceil(double x) { return ceil(x); }

// (-pi, pi] like the reference's while-loops (jitterbug.py:235-238, :313-316)
template <typename T> JB_HD T jb_wrap_pi(T a) {
  const T twopi = (T)(2 * JB_PI);
  if (a > (T)JB_PI || a <= -(T)JB_PI) a -= twopi * jb_ceil((a - (T)JB_PI) / twopi);
  return a;
}
template <typename T> JB_HD T jb_norm(T v, T lo, T hi) { return (v - lo) / (hi - lo) * (T)2 - (T)1; }

JB_HD int jb_obs_dim(int task) {
  return task == JB_TASK_MOVE_FROM_ORIGIN ? 15 : task == JB_TASK_FACE_DIRECTION ? 16
       : task == JB_TASK_MOVE_TO_POSITION ? 18 : 19;
}

// tolerance(x, bounds=(0,0)|(1,1), margin, gaussian, value_at_margin=0.1) == 0.1^((|x-b|/margin)^2)
template <typename T> JB_HD T jb_tol_gauss(T dist_to_bound, T margin) {
  if (dist_to_bound <= (T)0) return (T)1;
  T d = dist_to_bound / margin;
  return jb_exp(-(T)2.302585092994046 * d * d);
}

// Target pose parameters of one env: position x,y and the yaw quaternion (w, z) written by the reset
// (jitterbug.py:613-648 writes model.body_pos["target",x|y] and model.body_quat["target"]).
template <typename T> struct JbTarget { T x, y, qw, qz; };

// Observation (normalised, OrderedDict order) + reward for one env.  Returns the observation length.
// `s` is any state with fields pos[3], quat[4], vlin[3], w[3], qm, vm; `g` the packed model (jb_packed.h).
template <typename T, typename S>
JB_HD int jb_obs_reward(const T* g, const S& s, int task, const JbTarget<T>& tg, T* obs, T* reward) {
  T R[9];
  quat2mat(R, s.quat);
  T jyaw = jb_atan2(R[3], R[0]) - (T)(JB_PI / 2);
  // target rotation (about z): R00 = w^2 - z^2, R10 = 2wz
  T tc = tg.qw * tg.qw - tg.qz * tg.qz, ts = (T)2 * tg.qw * tg.qz;
  T tyaw = jb_atan2(ts, tc);
  T angle = jb_wrap_pi(tyaw - jyaw);
  T motor = jb_wrap_pi(s.qm + (T)(JB_PI / 2));
  T rel[3] = {tg.x - s.pos[0], tg.y - s.pos[1], g[G_TGT0 + 2] - s.pos[2]};
  T tib[3];
  mulMTV(tib, R, rel);
  // framelinvel sensor: world velocity of the root body's own COM = v + R (w x c0)
  T c0[3] = {g[G_C0], g[G_C0 + 1], g[G_C0 + 2]}, wc[3], sv[3];
  cross3(wc, s.w, c0);
  mulMV(sv, R, wc);
  sv[0] += s.vlin[0]; sv[1] += s.vlin[1]; sv[2] += s.vlin[2];
  // rotate into the target frame (inverse of a z-rotation by tyaw)
  T vit[3] = {tc * sv[0] + ts * sv[1], -ts * sv[0] + tc * sv[1], sv[2]};
  int n = 0;
  obs[n++] = jb_norm(s.pos[0], (T)-2, (T)2);
  obs[n++] = jb_norm(s.pos[1], (T)-2, (T)2);
  obs[n++] = jb_norm(s.pos[2], (T)0, (T)0.1);
#pragma unroll
  for (int i = 0; i < 4; i++) obs[n++] = jb_norm(s.quat[i], (T)-1, (T)1);
#pragma unroll
  for (int i = 0; i < 3; i++) obs[n++] = jb_norm(s.vlin[i], (T)-1, (T)1);
#pragma unroll
  for (int i = 0; i < 3; i++) obs[n++] = jb_norm(s.w[i], (T)-35, (T)35);
  obs[n++] = jb_norm(motor, -(T)JB_PI, (T)JB_PI);
  obs[n++] = jb_norm(s.vm, (T)-180, (T)180);

  T dist = jb_sqrt(dot3(tib, tib));
  T r_pos = jb_tol_gauss(dist, (T)0.05);
  T aa = jb_abs(angle);
  T r_head = aa < (T)(JB_PI / 2) ? (T)0.5 * ((T)1 + jb_cos((T)2 * aa)) : (T)0;
  T r_vel = vit[0] >= (T)0.1 ? (T)1 : (vit[0] > (T)0 ? vit[0] / (T)0.1 : (T)0);
  T r_up = jb_tol_gauss(jb_abs((T)1 - R[8]), (T)0.5);
  T r;
  switch (task) {
    case JB_TASK_MOVE_FROM_ORIGIN: r = (T)1 - r_pos; break;
    case JB_TASK_FACE_DIRECTION:
      obs[n++] = jb_norm(angle, -(T)JB_PI, (T)JB_PI); r = r_head; break;
    case JB_TASK_MOVE_IN_DIRECTION:
      obs[n++] = jb_norm(angle, -(T)JB_PI, (T)JB_PI);
#pragma unroll
      for (int i = 0; i < 3; i++) obs[n++] = jb_norm(vit[i], (T)-1, (T)1);
      r = r_vel; break;
    case JB_TASK_MOVE_TO_POSITION:
      obs[n++] = jb_norm(tib[0], (T)-3, (T)3); obs[n++] = jb_norm(tib[1], (T)-3, (T)3);
      obs[n++] = jb_norm(tib[2], (T)-0.1, (T)0.1);
      r = r_pos; break;
    default:
      obs[n++] = jb_norm(tib[0], (T)-3, (T)3); obs[n++] = jb_norm(tib[1], (T)-3, (T)3);
      obs[n++] = jb_norm(tib[2], (T)-0.1, (T)0.1);
      obs[n++] = jb_norm(angle, -(T)JB_PI, (T)JB_PI);
      r = r_pos * r_head; break;
  }
  *reward = r * r_up;
  return n;
}

// ------------------------------------------------------------------------------------------------ Philox4x32-10
struct JbPhilox { uint32_t v[4]; };
JB_HD uint32_t jb_mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
JB_HD JbPhilox jb_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  JbPhilox o;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = jb_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = jb_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
JB_HD float jb_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0, 1)

// Episode initial conditions (jitterbug.py:601-666) as a function of the six uniforms in [0, 1) the reference draws:
// u[0] -> angle = self.random.uniform(0, 2*pi), u[1] -> radius = self.random.uniform(.05, .2), u[2] -> yaw =
// np.random.uniform(0, 2*pi), u[3] -> rotation_angle = np.random.random() * 2*pi, u[4], u[5] -> rotation axis x, y =
// np.random.random(2) * .05 - .025.  Fills the root part of the state (pos, quat, zero velocities / motor); the caller
// zeroes its leg state and the warm start.
template <typename T, typename S>
JB_HD void jb_reset_from_uniforms(const T* g, S& s, JbTarget<T>& tg, int task, int random_pose, const T* u) {
  T angle = (T)(2 * JB_PI) * u[0];
  T radius = (T)0.05 + (T)0.15 * u[1];
  T yaw = (T)(2 * JB_PI) * u[2];
  T rot = (T)(2 * JB_PI) * u[3];
  T ax = (T)0.05 * u[4] - (T)0.025;
  T ay = (T)0.05 * u[5] - (T)0.025;
  tg.x = g[G_TGT0]; tg.y = g[G_TGT0 + 1]; tg.qw = (T)1; tg.qz = (T)0;
  T sa, ca;
  if (task == JB_TASK_MOVE_TO_POSITION || task == JB_TASK_MOVE_TO_POSE) {
    jb_sincos(angle, &sa, &ca);
    tg.x = radius * ca; tg.y = radius * sa;
  }
  if (task == JB_TASK_FACE_DIRECTION || task == JB_TASK_MOVE_IN_DIRECTION || task == JB_TASK_MOVE_TO_POSE) {
    jb_sincos((T)0.5 * yaw, &sa, &ca);
    tg.qw = ca; tg.qz = sa;
  }
  s.pos[0] = g[G_ROOTPOS0]; s.pos[1] = g[G_ROOTPOS0 + 1]; s.pos[2] = g[G_ROOTPOS0 + 2];
  s.quat[0] = 1; s.quat[1] = 0; s.quat[2] = 0; s.quat[3] = 0;
  if (random_pose) {
    T nrm = (T)1 / jb_sqrt(ax * ax + ay * ay + (T)1);
    jb_sincos((T)0.5 * rot, &sa, &ca);
    s.quat[0] = ca; s.quat[1] = sa * ax * nrm; s.quat[2] = sa * ay * nrm; s.quat[3] = sa * nrm;
  }
  s.qm = 0; s.vm = 0; s.rev = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) { s.vlin[i] = 0; s.w[i] = 0; }
}

// The uniforms come from Philox4x32-10 streams keyed by (seed, global env id, episode): the reference's two numpy
// generators cannot be reproduced bit for bit, the distributions are.
template <typename T, typename S>
JB_HD void jb_reset_env(const T* g, S& s, JbTarget<T>& tg, int task, int random_pose,
                        uint64_t seed, uint32_t env_gid, uint32_t episode) {
  JbPhilox r0 = jb_philox(env_gid, episode, 0u, 0x4A42u, (uint32_t)seed, (uint32_t)(seed >> 32));
  JbPhilox r1 = jb_philox(env_gid, episode, 1u, 0x4A42u, (uint32_t)seed, (uint32_t)(seed >> 32));
  const T u[6] = {(T)jb_u01(r0.v[0]), (T)jb_u01(r0.v[1]), (T)jb_u01(r0.v[2]), (T)jb_u01(r0.v[3]),
                  (T)jb_u01(r1.v[0]), (T)jb_u01(r1.v[1])};
  jb_reset_from_uniforms<T, S>(g, s, tg, task, random_pose, u);
}

// ------------------------------------------------------------------------------------------------ heuristic policies
template <typename T> JB_HD T jb_h_face(T angle_norm) {
  T v = (T)3 * angle_norm / (T)JB_PI;
  return (T)0.9 * jb_max(jb_min(v, (T)1), (T)-1);
}
template <typename T> JB_HD T jb_h_forward(T motor_angle, T motor_vel, T offset) {
  const T kick = (T)(JB_PI / 4), speed = (T)0.3;
  if (motor_angle < offset - kick) return speed;
  if (motor_angle > offset + kick) return -speed;
  return motor_vel > (T)0 ? speed : -speed;
}
// optimal_orientation_to_move: returns the adjusted angle, writes the motor-angle offset for the kick controller
template <typename T> JB_HD T jb_h_orient(T a, T* offset) {
  const T d45 = (T)(JB_PI / 4), d90 = (T)(JB_PI / 2);
  *offset = 0;
  if (a >= -d45 && a <= d45) return a;
  if (a > d45 && a <= (T)JB_PI) { *offset = d90; return jb_abs(jb_abs(a) - d90); }
  if (a >= -(T)JB_PI && a < -d45) { *offset = -d90; return -jb_abs(jb_abs(a) - d90); }
  return a;
}
// action from a normalised observation vector (layout of jb_obs_reward)
template <typename T> JB_HD T jb_heuristic(int task, const T* obs) {
  const T thr = (T)(20.0 * JB_PI / 180.0);
  T ma = obs[13], mv = obs[14];
  if (task == JB_TASK_MOVE_FROM_ORIGIN) return jb_h_forward(ma, mv, (T)0);
  if (task == JB_TASK_FACE_DIRECTION) return jb_h_face(obs[15]);
  if (task == JB_TASK_MOVE_IN_DIRECTION) {
    T off, a = jb_h_orient(obs[15], &off);
    return jb_abs(a) > thr ? jb_h_face(a) : jb_h_forward(ma, mv, off);
  }
  T dx = obs[15], dy = obs[16];
  T a = jb_atan2(dx, -dy);
  if (task == JB_TASK_MOVE_TO_POSITION) {
    T off;
    a = jb_h_orient(a, &off);
    return jb_abs(a) > thr ? jb_h_face(a) : jb_h_forward(ma, mv, off);
  }
  if (jb_abs(a) > thr) return jb_h_face(a);
  if (jb_sqrt(dx * dx + dy * dy) > (T)0.01) return jb_h_forward(ma, mv, (T)0);
  return jb_h_face(obs[18]);
}

#endif  // JB_TASK_CUH
