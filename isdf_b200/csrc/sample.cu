// K1 ray / depth sampling and K5 per-frame loss histogram (HBM-bound gather / scatter kernels).
//   reference: isdf/modules/sample.py:24-178, isdf/geometry/transform.py:13-41,
//              isdf/modules/loss.py:208-240
// Algorithmic bytes per ray: read 4 (depth) + 12 (normal) + 24 (3 x i64 index) + 64 (pose, L2-hot)
// + 4 (S-1) randoms; write 16 S (pc + z) + 12 + 64.  See DESIGN.md.
#include "common.cuh"

__global__ void gather_rays_kernel(const float* __restrict__ depth, const float* __restrict__ normals,
                                   const int64_t* __restrict__ frame_map, int normals_use_map,
                                   const int64_t* __restrict__ ib, const int64_t* __restrict__ ih,
                                   const int64_t* __restrict__ iw, int64_t n_rays, int H, int W,
                                   float* __restrict__ depth_out, float* __restrict__ normal_out,
                                   uint8_t* __restrict__ valid_out) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  int64_t b = ib[r], h = ih[r], w = iw[r];
  int64_t f = frame_map ? frame_map[b] : b;
  float d = depth[(f * H + h) * W + w];
  bool ok = d != 0.f;
  if (normals) {
    int64_t fn = normals_use_map ? f : b;     // reference quirk Q1: normals are NOT window-indexed
    const float* np = normals + ((fn * H + h) * W + w) * 3;
    float nx = np[0], ny = np[1], nz = np[2];
    ok = ok && !isnan(nx);
    normal_out[r * 3] = nx; normal_out[r * 3 + 1] = ny; normal_out[r * 3 + 2] = nz;
  }
  depth_out[r] = d;
  valid_out[r] = ok ? 1 : 0;
}

// one thread per (ray, sample)
__global__ void sample_rays_kernel(const float* __restrict__ T_WC, const int64_t* __restrict__ frame_map,
                                   const int64_t* __restrict__ ib, const int64_t* __restrict__ ih,
                                   const int64_t* __restrict__ iw, const float* __restrict__ dirs_in,
                                   const float* __restrict__ depth_s, const float* __restrict__ far_in,
                                   const float* __restrict__ near_in,
                                   const float* __restrict__ u_strat, const float* __restrict__ n_near,
                                   const float* __restrict__ lin, int64_t n_rays, int n_strat, int n_surf,
                                   isdfb_camera cam, float min_depth, float dist_behind,
                                   float* __restrict__ pc, float* __restrict__ z_vals,
                                   float* __restrict__ dirs_C, float* __restrict__ T_out) {
  const int S = n_strat + n_surf;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rays * S) return;
  int64_t r = i / S;
  int j = (int)(i - r * S);
  int64_t b = ib ? ib[r] : r;            // ib == NULL: T_WC is already per ray
  int64_t f = frame_map ? frame_map[b] : b;
  const float* T = T_WC + f * 16;
  float dx, dy, dz;
  if (dirs_in) {
    dx = dirs_in[r * 3]; dy = dirs_in[r * 3 + 1]; dz = dirs_in[r * 3 + 2];
  } else {                               // camera-frame direction (transform.py:13-33, depth_type 'z')
    dx = __fdiv_rn(__fsub_rn((float)iw[r], cam.cx), cam.fx);
    dy = __fdiv_rn(__fsub_rn((float)ih[r], cam.cy), cam.fy);
    dz = 1.0f;
  }
  // world-frame direction (transform.py:36-41)
  float wx = __fadd_rn(__fadd_rn(__fmul_rn(T[0], dx), __fmul_rn(T[1], dy)), __fmul_rn(T[2], dz));
  float wy = __fadd_rn(__fadd_rn(__fmul_rn(T[4], dx), __fmul_rn(T[5], dy)), __fmul_rn(T[6], dz));
  float wz = __fadd_rn(__fadd_rn(__fmul_rn(T[8], dx), __fmul_rn(T[9], dy)), __fmul_rn(T[10], dz));
  float d = depth_s ? depth_s[r] : 0.f;
  float far = far_in ? far_in[r] : __fadd_rn(d, dist_behind);
  if (near_in) min_depth = near_in[r];                   // per-ray near limit (render passes, trainer.py:1121-1128)
  float z;
  if (j < n_surf) {
    if (j == 0) {
      z = d;                                             // on-surface sample (sample.py:159)
    } else {
      float v = __fadd_rn(d, n_near[r * (n_surf - 1) + (j - 1)]);
      z = fminf(fmaxf(v, min_depth), far);               // clamp (sample.py:167-171)
    }
  } else {
    int q = j - n_surf;                                  // stratified bin (sample.py:92-128)
    float range = __fsub_rn(far, min_depth);
    float lower = __fadd_rn(__fmul_rn(lin[q], range), min_depth);
    float bin_len = __fdiv_rn(range, (float)n_strat);
    z = __fadd_rn(lower, __fmul_rn(u_strat[r * n_strat + q], bin_len));
  }
  z_vals[i] = z;
  pc[i * 3 + 0] = __fadd_rn(T[3], __fmul_rn(wx, z));
  pc[i * 3 + 1] = __fadd_rn(T[7], __fmul_rn(wy, z));
  pc[i * 3 + 2] = __fadd_rn(T[11], __fmul_rn(wz, z));
  if (j == 0) {
    dirs_C[r * 3] = dx; dirs_C[r * 3 + 1] = dy; dirs_C[r * 3 + 2] = dz;
#pragma unroll
    for (int k = 0; k < 16; ++k) T_out[r * 16 + k] = T[k];
  }
}

// ---- K5 ------------------------------------------------------------------------------------
// Rays are ordered by frame (sample.py:18-19 builds indices_b with repeat_interleave and the
// compaction keeps order).  A ray is "live" unless a LATER ray hits the same pixel of the same
// frame (CPU index_put: last writer wins; the mask counts the pixel once).
// one warp per ray: the 32 lanes scan the later rays of the same frame for a duplicate pixel and add up
// the ray's samples; lane 0 commits.
__global__ void frame_bins_accum_kernel(const float* __restrict__ loss_mat, const uint8_t* __restrict__ ray_valid,
                                        const int64_t* __restrict__ ib, const int64_t* __restrict__ ih,
                                        const int64_t* __restrict__ iw, int64_t n_rays, int S, int H, int W,
                                        int factor, float* __restrict__ bins, float* __restrict__ cnt) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rays) return;
  if (ray_valid && !ray_valid[r]) return;
  const int64_t b = ib[r], h = ih[r], w = iw[r];
  bool shadowed = false;
  for (int64_t q0 = r + 1; q0 < n_rays; q0 += 32) {
    const int64_t q = q0 + lane;
    bool same_frame = false, hit = false;
    if (q < n_rays) {
      same_frame = ib[q] == b;
      hit = same_frame && ih[q] == h && iw[q] == w && (!ray_valid || ray_valid[q]);
    }
    if (__any_sync(0xffffffffu, hit)) { shadowed = true; break; }
    if (!__all_sync(0xffffffffu, same_frame)) break;          // rays are ordered by frame
  }
  if (shadowed) return;
  float s = 0.f;
  for (int j = lane; j < S; j += 32) s += loss_mat[r * S + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const int by = (int)(h / (H / factor)), bx = (int)(w / (W / factor));
    const int64_t cell = (b * factor + by) * factor + bx;
    atomicAdd(bins + cell, s);
    atomicAdd(cnt + cell, 1.0f);
  }
}

__global__ void frame_bins_final_kernel(const float* __restrict__ bins, const float* __restrict__ cnt,
                                        int n_frames, int factor, float* __restrict__ loss_approx,
                                        float* __restrict__ frame_avg) {
  int f = blockIdx.x;
  int cells = factor * factor;
  float s = 0.f;
  for (int c = threadIdx.x; c < cells; c += blockDim.x) {
    float n = cnt[f * cells + c];
    float v = bins[f * cells + c] / (n == 0.f ? 1.f : n);
    loss_approx[f * cells + c] = v;
    s += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    frame_avg[f] = t / (float)cells;
  }
}

// ---- N3: frame ingest ------------------------------------------------------------------------
// Per-pixel surface normals of a depth image (reference transform.py:169-196 + 215-270): back-project
// p = (z (c-cx)/fx, z (r-cy)/fy, z); among the 8 neighbour pairs (k, k+2) at pixel distance 2 pick the
// one minimising |p_k - p| + |p_{k+2} - p| (out-of-image = NaN = infinite cost, first minimum wins) and
// return the normalised cross product.  One thread per pixel; 16 fused temporaries instead of HBM tensors.
__global__ void ingest_normals_kernel(const float* __restrict__ depth, isdfb_camera cam, float* __restrict__ normals) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  if (c >= cam.W || r >= cam.H) return;
  const int offs[8][2] = {{-2, 0}, {-2, 2}, {0, 2}, {2, 2}, {2, 0}, {2, -2}, {0, -2}, {-2, -2}};   // (dy, dx)
  auto point = [&](int rr, int cc, float* o) {
    const float z = depth[(size_t)rr * cam.W + cc];
    o[0] = z * ((float)cc - cam.cx) / cam.fx;
    o[1] = z * ((float)rr - cam.cy) / cam.fy;
    o[2] = z;
  };
  float p[3];
  point(r, c, p);
  float dv[8][3], len[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int rr = r + offs[k][0], cc = c + offs[k][1];
    if (rr < 0 || rr >= cam.H || cc < 0 || cc >= cam.W) {
      dv[k][0] = dv[k][1] = dv[k][2] = __int_as_float(0x7fc00000);
      len[k] = __int_as_float(0x7f800000);
    } else {
      float q[3];
      point(rr, cc, q);
      dv[k][0] = q[0] - p[0]; dv[k][1] = q[1] - p[1]; dv[k][2] = q[2] - p[2];
      len[k] = sqrtf(dv[k][0] * dv[k][0] + dv[k][1] * dv[k][1] + dv[k][2] * dv[k][2]);
      if (isnan(len[k])) len[k] = __int_as_float(0x7f800000);
    }
  }
  int best = 0;
  float bc = len[0] + len[2];
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const float cst = len[k] + len[(k + 2) & 7];
    if (cst < bc) { bc = cst; best = k; }
  }
  float a[3], b[3];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k == best) {
#pragma unroll
      for (int t = 0; t < 3; ++t) { a[t] = dv[k][t]; b[t] = dv[(k + 2) & 7][t]; }
    }
  const float nx = a[1] * b[2] - a[2] * b[1], ny = a[2] * b[0] - a[0] * b[2], nz = a[0] * b[1] - a[1] * b[0];
  const float inv = 1.f / sqrtf(nx * nx + ny * ny + nz * nz);
  float* o = normals + ((size_t)r * cam.W + c) * 3;
  o[0] = nx * inv; o[1] = ny * inv; o[2] = nz * inv;
}

int sample_ingest_normals(isdfb_ctx* ctx, const float* depth, const isdfb_camera* cam, float* normals, cudaStream_t st) {
  dim3 blk(32, 8), grd((cam->W + 31) / 32, (cam->H + 7) / 8);
  ingest_normals_kernel<<<grd, blk, 0, st>>>(depth, *cam, normals);
  ISDFB_LAUNCHED(ctx);
  ISDFB_CUDA_OK(ctx, cudaGetLastError());
  return ISDFB_OK;
}

int sample_gather(isdfb_ctx* ctx, const float* depth, const float* normals, const int64_t* frame_map,
                  int normals_use_map, const int64_t* ib, const int64_t* ih, const int64_t* iw,
                  int64_t n_rays, const isdfb_camera* cam, float* depth_out, float* normal_out,
                  uint8_t* valid_out, cudaStream_t st) {
  if (n_rays == 0) return ISDFB_OK;
  gather_rays_kernel<<<(unsigned)((n_rays + 127) / 128), 128, 0, st>>>(
      depth, normals, frame_map, normals_use_map, ib, ih, iw, n_rays, cam->H, cam->W, depth_out, normal_out, valid_out);
  ISDFB_LAUNCHED(ctx);
  ISDFB_CUDA_OK(ctx, cudaGetLastError());
  return ISDFB_OK;
}

int sample_along(isdfb_ctx* ctx, const float* T_WC, const int64_t* frame_map, const int64_t* ib,
                 const int64_t* ih, const int64_t* iw, const float* dirs_in, const float* depth_sample,
                 const float* far_in, const float* near_in, const float* u_strat,
                 const float* n_near, const float* lin, int64_t n_rays, int n_strat, int n_surf,
                 const isdfb_camera* cam, float min_depth, float dist_behind, float* pc, float* z_vals,
                 float* dirs_C, float* T_out, cudaStream_t st) {
  if (n_rays == 0) return ISDFB_OK;
  int64_t total = n_rays * (n_strat + n_surf);
  sample_rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      T_WC, frame_map, ib, ih, iw, dirs_in, depth_sample, far_in, near_in, u_strat, n_near, lin, n_rays, n_strat, n_surf, *cam,
      min_depth, dist_behind, pc, z_vals, dirs_C, T_out);
  ISDFB_LAUNCHED(ctx);
  ISDFB_CUDA_OK(ctx, cudaGetLastError());
  return ISDFB_OK;
}

int sample_frame_bins(isdfb_ctx* ctx, float* scratch, const float* loss_mat, const uint8_t* ray_valid,
                      const int64_t* ib, const int64_t* ih, const int64_t* iw, int64_t n_rays, int S,
                      int n_frames, int H, int W, int factor, float* loss_approx, float* frame_avg,
                      cudaStream_t st) {
  int cells = n_frames * factor * factor;
  float* bins = scratch;
  float* cnt = scratch + cells;
  ISDFB_CUDA_OK(ctx, cudaMemsetAsync(scratch, 0, 2 * (size_t)cells * sizeof(float), st));
  if (n_rays > 0) {
    frame_bins_accum_kernel<<<(unsigned)((n_rays * 32 + 255) / 256), 256, 0, st>>>(
        loss_mat, ray_valid, ib, ih, iw, n_rays, S, H, W, factor, bins, cnt);
    ISDFB_LAUNCHED(ctx);
  }
  frame_bins_final_kernel<<<n_frames, 64, 0, st>>>(bins, cnt, n_frames, factor, loss_approx, frame_avg);
  ISDFB_LAUNCHED(ctx);
  ISDFB_CUDA_OK(ctx, cudaGetLastError());
  return ISDFB_OK;
}
