// Self-clearance monitor (sm_100a).
//
// MuJoCo tests every pair of geoms whose bodies are not parent and child: besides the 22 floor pairs
// of the step kernel that is leg <-> leg, leg <-> rotor and lower leg <-> core.  The step kernel generates none of
// them, on the argument that the Jitterbug's geometry keeps them apart in every state it reaches; this kernel MEASURES
// that claim on the live batch instead of assuming it.  One thread per env rebuilds the root-frame pose of every moving
// collision geom from the stored joint angles (same product-of-exponentials kinematics as jb_quad.cuh), covers each by
// a conservative primitive -- cylinders by capsules (a cylinder plus end caps), boxes by their bounding spheres -- and
// takes the smallest distance over all primitive pairs of non-filtered body pairs.  The two flat ellipsoids (a bounding
// sphere would swallow the legs next to them) are kept exact and bounded from below in their own unit-sphere
// coordinates: with x' = S^-1 R'(x - c) a point at distance d from the ellipsoid satisfies d >= s_min (|x'| - 1), so
// clearance(capsule) >= s_min (dist(0, segment') - 1) - r, positive exactly when the segment is outside the ellipsoid
// by more than ~r; the better of this bound and the bounding-sphere one is used.  Because the covers only grow the geoms (and the ellipsoid bound only shrinks the distance), a
// positive result proves that MuJoCo's narrow phase would have found no contact there.
//
// jb_sim_step launches it every JB_MONITOR_PERIOD-th control step (a sampled monitor, ~0.1 % of the step time);
// counters: [5] env-steps checked, [6] env-steps with a clearance below zero, [7] the smallest clearance seen, in nm.
#include "jb_sim_common.cuh"
#include "jb_math.cuh"

namespace {
struct Prim { float a[3], b[3], r; int body; };      // segment a-b swept by a sphere of radius r (a == b: a sphere)
struct Ell { float c[3], R[9], s[3]; int body; };    // centre, rotation (local -> root frame), semi-axes
__device__ float seg_seg_dist2(const float* p1, const float* q1, const float* p2, const float* q2);

// lower bound of the distance between the capsule p and the ellipsoid e (see the head of this file)
__device__ float ell_clearance(const Ell& e, const Prim& p) {
  float a[3], b[3], t[3], u[3];
  sub3(t, p.a, e.c); mulMTV(a, e.R, t);
  sub3(t, p.b, e.c); mulMTV(b, e.R, t);
#pragma unroll
  for (int i = 0; i < 3; i++) { a[i] /= e.s[i]; b[i] /= e.s[i]; }
  sub3(u, b, a);
  const float uu = dot3(u, u);
  const float tt = uu > 1e-20f ? fminf(fmaxf(-dot3(a, u) / uu, 0.f), 1.f) : 0.f;
  const float x[3] = {a[0] + tt * u[0], a[1] + tt * u[1], a[2] + tt * u[2]};
  const float smin = fminf(e.s[0], fminf(e.s[1], e.s[2])), smax = fmaxf(e.s[0], fmaxf(e.s[1], e.s[2]));
  // second valid bound: the bounding sphere (tight for a very flat ellipsoid seen from far away, where s_min is tiny)
  const float zero[3] = {0.f, 0.f, 0.f};
  float ta[3], tb[3];
  sub3(ta, p.a, e.c); sub3(tb, p.b, e.c);
  const float far = sqrtf(seg_seg_dist2(ta, tb, zero, zero)) - smax;
  return fmaxf(smin * (sqrtf(dot3(x, x)) - 1.f), far) - p.r;
}

// bodies: 0 core, 1 + 2k upper leg k, 2 + 2k lower leg k, 9 rotor.  MuJoCo filters parent-child pairs.
__device__ __forceinline__ bool filtered(int b1, int b2) {
  if (b1 > b2) { const int t = b1; b1 = b2; b2 = t; }
  if (b1 == b2) return true;
  if (b1 == 0) return b2 == 9 || (b2 & 1) == 1;       // core - rotor, core - upper leg
  return (b1 & 1) == 1 && b2 == b1 + 1;               // upper leg - its lower leg
}

// squared distance between segments p1-q1 and p2-q2 (Ericson, Real-Time Collision Detection 5.1.9)
__device__ float seg_seg_dist2(const float* p1, const float* q1, const float* p2, const float* q2) {
  float d1[3], d2[3], r[3];
  sub3(d1, q1, p1); sub3(d2, q2, p2); sub3(r, p1, p2);
  const float a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
  float s, t;
  const float eps = 1e-12f;
  if (a <= eps && e <= eps) { s = t = 0.f; }
  else if (a <= eps) { s = 0.f; t = fminf(fmaxf(f / e, 0.f), 1.f); }
  else {
    const float c = dot3(d1, r);
    if (e <= eps) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
    else {
      const float b = dot3(d1, d2), den = a * e - b * b;
      s = den > eps ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;
      t = (b * s + f) / e;
      if (t < 0.f) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
      else if (t > 1.f) { t = 1.f; s = fminf(fmaxf((b - c) / a, 0.f), 1.f); }
    }
  }
  float d[3] = {r[0] + d1[0] * s - d2[0] * t, r[1] + d1[1] * s - d2[1] * t, r[2] + d1[2] * s - d2[2] * t};
  return dot3(d, d);
}

__device__ __forceinline__ void capsule(Prim& p, const float* c, const float* ax, float hl, float r, int body) {
#pragma unroll
  for (int i = 0; i < 3; i++) { p.a[i] = c[i] - hl * ax[i]; p.b[i] = c[i] + hl * ax[i]; }
  p.r = r; p.body = body;
}
__device__ __forceinline__ void sphere(Prim& p, const float* c, float r, int body) {
#pragma unroll
  for (int i = 0; i < 3; i++) { p.a[i] = c[i]; p.b[i] = c[i]; }
  p.r = r; p.body = body;
}
// x = p + R (h - p0): a home-pose point h of a body hinged at p (home anchor p0), rotated by R
__device__ __forceinline__ void place(float* x, const float* p, const float* R, const float* h, const float* p0) {
  float r[3], t[3];
  sub3(r, h, p0); mulMV(t, R, r); add3(x, p, t);
}
}  // namespace

__global__ void __launch_bounds__(128) jb_self_clearance_kernel(const __grid_constant__ JbDev d, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= d.N) return;
  const int N = d.N;
  const float* P = d.packed + (d.per_env ? (size_t)e * JBP_SIZE : 0);
  Prim pr[20];
  Ell el[2];
  int n = 0;
  // core body (root frame == body frame): two boxes, screw cylinder, screw ellipsoid
  for (int l = 0; l < 2; l++) {
    const float* X = P + JBP_X + l;
    const float c[3] = {X[4 * X_BOX_C], X[4 * (X_BOX_C + 1)], X[4 * (X_BOX_C + 2)]};
    const float h[3] = {X[4 * X_BOX_H], X[4 * (X_BOX_H + 1)], X[4 * (X_BOX_H + 2)]};
    sphere(pr[n++], c, sqrtf(dot3(h, h)), 0);
  }
  {
    const float* X = P + JBP_X + 2;
    const float c[3] = {X[4 * X_CYL_C], X[4 * (X_CYL_C + 1)], X[4 * (X_CYL_C + 2)]};
    const float ax[3] = {X[4 * X_CYL_AX], X[4 * (X_CYL_AX + 1)], X[4 * (X_CYL_AX + 2)]};
    capsule(pr[n++], c, ax, X[4 * X_CYL_HL], X[4 * X_CYL_R], 0);
#pragma unroll
    for (int i = 0; i < 3; i++) { el[0].c[i] = X[4 * (X_ELL_C + i)]; el[0].s[i] = X[4 * (X_ELL_S + i)]; }
#pragma unroll
    for (int i = 0; i < 9; i++) el[0].R[i] = X[4 * (X_ELL_R + i)];
    el[0].body = 0;
  }
  // legs
  for (int k = 0; k < 4; k++) {
    const float* L = P + JBP_L + k;
    const float qa = d.qpos[(7 + 2 * k) * N + e], qb = d.qpos[(8 + 2 * k) * N + e];
    float a1[3], p1[3], a2h[3], p2h[3], R1[9], R2[9], R12[9], p2[3], c[3], h[3], ax[3], axh[3], sa, ca, sb, cb;
#pragma unroll
    for (int i = 0; i < 3; i++) { a1[i] = L[4 * (L_A1 + i)]; p1[i] = L[4 * (L_P1 + i)]; a2h[i] = L[4 * (L_A2 + i)]; p2h[i] = L[4 * (L_P2 + i)]; }
    sincosf(qa, &sa, &ca); sincosf(qb, &sb, &cb);
    rodrigues(R1, a1, sa, ca); rodrigues(R2, a2h, sb, cb); mulMM(R12, R1, R2);
    place(p2, p1, R1, p2h, p1);
#pragma unroll
    for (int i = 0; i < 3; i++) { h[i] = L[4 * (L_UC_C + i)]; axh[i] = L[4 * (L_UC_AX + i)]; }
    place(c, p1, R1, h, p1); mulMV(ax, R1, axh);
    capsule(pr[n++], c, ax, L[4 * L_UC_HL], L[4 * L_UC_R], 1 + 2 * k);
#pragma unroll
    for (int i = 0; i < 3; i++) h[i] = L[4 * (L_KN_C + i)];
    place(c, p1, R1, h, p1);
    sphere(pr[n++], c, L[4 * L_KN_R], 1 + 2 * k);
#pragma unroll
    for (int i = 0; i < 3; i++) { h[i] = L[4 * (L_LC_C + i)]; axh[i] = L[4 * (L_LC_AX + i)]; }
    place(c, p2, R12, h, p2h); mulMV(ax, R12, axh);
    capsule(pr[n++], c, ax, L[4 * L_LC_HL], L[4 * L_LC_R], 2 + 2 * k);
#pragma unroll
    for (int i = 0; i < 3; i++) h[i] = L[4 * (L_FT_C + i)];
    place(c, p2, R12, h, p2h);
    sphere(pr[n++], c, L[4 * L_FT_R], 2 + 2 * k);
  }
  // rotor: thread cylinder + mass ellipsoid
  {
    const float* X = P + JBP_X + 3;
    const float pm[3] = {P[G_PM], P[G_PM + 1], P[G_PM + 2]}, am[3] = {P[G_AM], P[G_AM + 1], P[G_AM + 2]};
    float Rm[9], sm, cmn, c[3], ax[3];
    sincosf(d.qpos[15 * N + e], &sm, &cmn);
    rodrigues(Rm, am, sm, cmn);
    const float h[3] = {X[4 * X_CYL_C], X[4 * (X_CYL_C + 1)], X[4 * (X_CYL_C + 2)]};
    const float axh[3] = {X[4 * X_CYL_AX], X[4 * (X_CYL_AX + 1)], X[4 * (X_CYL_AX + 2)]};
    place(c, pm, Rm, h, pm); mulMV(ax, Rm, axh);
    capsule(pr[n++], c, ax, X[4 * X_CYL_HL], X[4 * X_CYL_R], 9);
    const float he[3] = {X[4 * X_ELL_C], X[4 * (X_ELL_C + 1)], X[4 * (X_ELL_C + 2)]};
    float Rh[9];
    place(el[1].c, pm, Rm, he, pm);
#pragma unroll
    for (int i = 0; i < 9; i++) Rh[i] = X[4 * (X_ELL_R + i)];
    mulMM(el[1].R, Rm, Rh);
#pragma unroll
    for (int i = 0; i < 3; i++) el[1].s[i] = X[4 * (X_ELL_S + i)];
    el[1].body = 9;
  }
  float best = 1e30f;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      if (filtered(pr[i].body, pr[j].body)) continue;
      const float dist = sqrtf(seg_seg_dist2(pr[i].a, pr[i].b, pr[j].a, pr[j].b)) - pr[i].r - pr[j].r;
      best = fminf(best, dist);
    }
  for (int k = 0; k < 2; k++)                          // (the two ellipsoids sit on parent and child: no pair of their own)
    for (int i = 0; i < n; i++)
      if (!filtered(el[k].body, pr[i].body)) best = fminf(best, ell_clearance(el[k], pr[i]));
  if (out) out[e] = best;
  // counters, one set of atomics per warp
  const unsigned m = __activemask();
  const int nm = (int)fminf(fmaxf(best * 1e9f, -1e9f), 2e9f);               // nanometres
  const int wmin = __reduce_min_sync(m, nm);
  const int below = __reduce_add_sync(m, best < 0.f ? 1 : 0);
  if ((threadIdx.x & 31) == __ffs(m) - 1) {
    atomicAdd(&d.counters[5], (unsigned long long)__popc(m));
    if (below) atomicAdd(&d.counters[6], (unsigned long long)below);
    atomicMin(reinterpret_cast<long long*>(&d.counters[7]), (long long)wmin);
  }
}

cudaError_t jb_launch_self_clearance(const JbDev& d, float* out_dev, cudaStream_t st) {
  jb_self_clearance_kernel<<<(d.N + 127) / 128, 128, 0, st>>>(d, out_dev);
  return cudaGetLastError();
}
