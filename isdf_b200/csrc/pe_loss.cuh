// Device functions shared by the CUDA-core and tensor-core paths:
//   icosahedral positional encoding (reference embedding.py:25-111) and the per-sample loss
//   with its closed-form adjoints (reference loss.py:13-22,122-205; trainer.py:814-836).
#pragma once
#include "common.cuh"

// 21 icosahedron directions, row d = (x,y,z)   (embedding.py:40-62)
__constant__ float c_ico[ISDFB_NDIRS][3] = {
    {0.8506508f, 0.f, 0.5257311f},   {0.809017f, 0.5f, 0.309017f},   {0.5257311f, 0.8506508f, 0.f},
    {1.f, 0.f, 0.f},                 {0.809017f, 0.5f, -0.309017f},  {0.8506508f, 0.f, -0.5257311f},
    {0.309017f, 0.809017f, -0.5f},   {0.f, 0.5257311f, -0.8506508f}, {0.5f, 0.309017f, -0.809017f},
    {0.f, 1.f, 0.f},                 {-0.5257311f, 0.8506508f, 0.f}, {-0.309017f, 0.809017f, -0.5f},
    {0.f, 0.5257311f, 0.8506508f},   {-0.309017f, 0.809017f, 0.5f},  {0.309017f, 0.809017f, 0.5f},
    {0.5f, 0.309017f, 0.809017f},    {0.5f, -0.309017f, 0.809017f},  {0.f, 0.f, 1.f},
    {-0.5f, 0.309017f, 0.809017f},   {-0.809017f, 0.5f, 0.309017f},  {-0.809017f, 0.5f, -0.309017f}};

#define ISDFB_HALF_PI_F 1.57079637050628662109375f   // fp32(0.5*pi), what torch adds (embedding.py:108)

// x' = scale * (R x + t)    (embedding.py:12-22, transform.py:287-304)
__device__ __forceinline__ void pe_scale_input(const PEParams& pe, float x, float y, float z, float xs[3]) {
  if (pe.has_transform) {
    float a = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pe.R[0], x), __fmul_rn(pe.R[1], y)), __fmul_rn(pe.R[2], z)), pe.t[0]);
    float b = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pe.R[3], x), __fmul_rn(pe.R[4], y)), __fmul_rn(pe.R[5], z)), pe.t[1]);
    float c = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pe.R[6], x), __fmul_rn(pe.R[7], y)), __fmul_rn(pe.R[8], z)), pe.t[2]);
    x = a; y = b; z = c;
  }
  xs[0] = __fmul_rn(x, pe.scale);
  xs[1] = __fmul_rn(y, pe.scale);
  xs[2] = __fmul_rn(z, pe.scale);
}

__device__ __forceinline__ float pe_project(const float xs[3], int d) {
  return fmaf(xs[2], c_ico[d][2], fmaf(xs[1], c_ico[d][1], xs[0] * c_ico[d][0]));
}

// feature k of the embedding (k < E), layout [xs(3), sin(xb)(21F), sin(xb+pi/2)(21F)], xb index d*F+f
__device__ __forceinline__ float pe_feature(const PEParams& pe, const float xs[3], int k) {
  if (k < 3) return xs[k];
  const int F = pe.n_freqs;
  int q = k - 3;
  const int half = ISDFB_NDIRS * F;
  bool shifted = q >= half;
  if (shifted) q -= half;
  int d = q / F, f = q - d * F;
  float xb = pe_project(xs, d) * (float)(1 << f);
  return shifted ? sinf(__fadd_rn(xb, ISDFB_HALF_PI_F)) : sinf(xb);
}

// ---------------------------------------------------------------------------------------
struct LossPoint {
  float l_sdf, l_grad, l_eik, total;
  float sbar;       // d mean(total) / d sdf
  float gbar[3];    // d mean(total) / d g
};

// u = target direction for the normal term: surface normal for sample 0, -dir_W otherwise.
__device__ __forceinline__ LossPoint loss_point(const isdfb_loss_cfg& c, float sdf, const float g[3],
                                                float bnd, const float u[3]) {
  LossPoint o;
  const float inv_n = c.inv_count_dev ? __ldg(c.inv_count_dev) : c.inv_count;
  // free-space / truncation SDF loss (loss.py:122-164)
  float relu_t = fmaxf(sdf - bnd, 0.f);
  float ex = expf(-5.0f * sdf);
  float exp_t = ex - 1.0f;
  float d_relu = (sdf > bnd) ? 1.f : 0.f;
  float d_exp = -5.0f * ex;
  float d_fs = (relu_t > exp_t) ? d_relu : ((relu_t < exp_t) ? d_exp : 0.5f * (d_relu + d_exp));
  bool free_sp = bnd > c.trunc_distance;
  float raw = free_sp ? fmaxf(relu_t, exp_t) : (sdf - bnd);
  float d_raw = free_sp ? d_fs : 1.f;
  float w = free_sp ? 1.f : c.trunc_weight;
  float l, outer;
  if (c.loss_type == 1) {
    l = fabsf(raw);
    outer = (raw > 0.f) ? 1.f : ((raw < 0.f) ? -1.f : 0.f);
  } else {
    l = raw * raw;
    outer = 2.f * raw;
  }
  o.l_sdf = l * w;
  o.sbar = outer * d_raw * w * inv_n;

  float ng = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
  float gb0 = 0.f, gb1 = 0.f, gb2 = 0.f;
  o.l_eik = 0.f;
  if (c.eik_weight != 0.f) {       // trainer.py:814-816, loss.py:196-200
    bool on = !(bnd < c.eik_apply_dist);
    float dev = ng - 1.f;
    o.l_eik = on ? fabsf(dev) * c.eik_weight : 0.f;
    float sg = (dev > 0.f) ? 1.f : ((dev < 0.f) ? -1.f : 0.f);
    float coef = on ? c.eik_weight * sg / fmaxf(ng, 1e-30f) : 0.f;
    gb0 = coef * g[0]; gb1 = coef * g[1]; gb2 = coef * g[2];
  }
  o.l_grad = 0.f;
  if (c.grad_weight != 0.f) {      // trainer.py:818-830 (CosineSimilarity eps 1e-6, per-vector clamp)
    float nu = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-6f);
    float ngc = fmaxf(ng, 1e-6f);
    float dot = u[0] * g[0] + u[1] * g[1] + u[2] * g[2];
    float inv = 1.f / (nu * ngc);
    float cosv = dot * inv;
    o.l_grad = 1.f - cosv;
    if (c.orien_loss) {
      o.l_grad = (o.l_grad > 1.f) ? 1.f : 0.f;
    } else {
      float k2 = (ng > 1e-6f) ? dot * inv / (ngc * ngc) : 0.f;
      gb0 -= c.grad_weight * (u[0] * inv - k2 * g[0]);
      gb1 -= c.grad_weight * (u[1] * inv - k2 * g[1]);
      gb2 -= c.grad_weight * (u[2] * inv - k2 * g[2]);
    }
  }
  o.total = o.l_sdf + c.grad_weight * o.l_grad + o.l_eik;
  o.gbar[0] = gb0 * inv_n; o.gbar[1] = gb1 * inv_n; o.gbar[2] = gb2 * inv_n;
  return o;
}

// Bound and normal-term target of sample (r, j), global sample index pg = r*S + j.
//   'ray'  (c.bounds_dev == NULL): bnd = ||d_C|| (depth - z), u = -R_WC d_C      (loss.py:13-22, 48-53)
//   'pc'   (c.bounds_dev != NULL): bnd / u precomputed by isdfb_bounds_pc        (loss.py:56-89); a NaN
//          direction (sample on top of its closest surface point) becomes the ray's normal (trainer.py:823-824)
// Sample 0 always targets the surface normal when normals are given (trainer.py:819-820).
__device__ __forceinline__ void loss_bound_target(const isdfb_loss_cfg& c, int64_t pg, int64_t r, int j,
                                                  const float* __restrict__ dirs_C, const float* __restrict__ depth,
                                                  const float* __restrict__ z_vals, const float* __restrict__ T_WC,
                                                  const float* __restrict__ normals, float& bnd, float u[3]) {
  const float dc[3] = {dirs_C[r * 3], dirs_C[r * 3 + 1], dirs_C[r * 3 + 2]};
  if (c.bounds_dev) {
    bnd = c.bounds_dev[pg];
  } else {
    const float nrm = sqrtf(dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]);
    bnd = nrm * (depth[r] - z_vals[pg]);
  }
  if (j == 0 && normals) {
    u[0] = normals[r * 3]; u[1] = normals[r * 3 + 1]; u[2] = normals[r * 3 + 2];
  } else if (c.grad_vec_dev) {
    u[0] = c.grad_vec_dev[pg * 3]; u[1] = c.grad_vec_dev[pg * 3 + 1]; u[2] = c.grad_vec_dev[pg * 3 + 2];
    if (u[0] != u[0] && normals) { u[0] = normals[r * 3]; u[1] = normals[r * 3 + 1]; u[2] = normals[r * 3 + 2]; }
  } else {
    const float* Tm = T_WC + r * 16;
    u[0] = -(Tm[0] * dc[0] + Tm[1] * dc[1] + Tm[2] * dc[2]);
    u[1] = -(Tm[4] * dc[0] + Tm[5] * dc[1] + Tm[6] * dc[2]);
    u[2] = -(Tm[8] * dc[0] + Tm[9] * dc[1] + Tm[10] * dc[2]);
  }
}

// softplus(beta=100, threshold=20) and its first/second derivative factors (fc_map.py:54)
__device__ __forceinline__ void softplus100(float z, float& h, float& sig) {
  float bz = 100.f * z;
  if (bz > 20.f) { h = z; sig = 1.f; return; }
  float t = expf(bz);
  h = log1pf(t) * 0.01f;
  sig = t / (1.f + t);
}
__device__ __forceinline__ float sigma_prime(float sig) {   // beta * sig * (1 - sig); 0 in the linear regime
  return (sig >= 1.f) ? 0.f : 100.f * sig * (1.f - sig);
}
