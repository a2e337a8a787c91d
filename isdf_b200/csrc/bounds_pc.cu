// N2: the "batch distance" bound of the reference's bounds_method = 'pc' (loss.py:56-89).
// For each of the N = R*S samples find the closest of the R surface samples (pc[r,0,:]).  All-pairs
// search: one thread per sample, the surface points stream through shared memory in tiles of 512
// (6 KB), every thread keeps (min distance, argmin) in registers.  Ties resolve to the lowest ray
// index, like torch.min on CPU.  FLOP-bound on the CUDA cores (8 flops + a sqrt per pair); the grid is
// N/256 CTAs, 211 at the default workload, >= 1 wave for every larger one.
#include "common.cuh"

#define BPC_THREADS 256
#define BPC_TILE 512

__global__ void __launch_bounds__(BPC_THREADS) bounds_pc_kernel(const float* __restrict__ pc, const float* __restrict__ z_vals,
                                                                 const float* __restrict__ depth,
                                                                 const uint8_t* __restrict__ ray_valid, int64_t n_rays, int S,
                                                                 float* __restrict__ bounds, float* __restrict__ grad_vec) {
  __shared__ float sx[BPC_TILE], sy[BPC_TILE], sz[BPC_TILE];
  const int64_t n = n_rays * S;
  const int64_t i = (int64_t)blockIdx.x * BPC_THREADS + threadIdx.x;
  const bool live = i < n;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (live) { px = pc[i * 3]; py = pc[i * 3 + 1]; pz = pc[i * 3 + 2]; }
  float best = __int_as_float(0x7f800000);   // +inf
  float bx = 0.f, by = 0.f, bz = 0.f;
  for (int64_t t0 = 0; t0 < n_rays; t0 += BPC_TILE) {
    const int m = (int)((n_rays - t0 < BPC_TILE) ? (n_rays - t0) : BPC_TILE);
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += BPC_THREADS) {
      const int64_t r = t0 + k;
      const bool ok = ray_valid ? (ray_valid[r] != 0) : true;
      const float* s = pc + r * S * 3;          // surface sample = sample 0 of ray r
      // an invalid ray offers no surface point: park it at +inf so it never wins
      sx[k] = ok ? s[0] : __int_as_float(0x7f800000);
      sy[k] = ok ? s[1] : 0.f;
      sz[k] = ok ? s[2] : 0.f;
    }
    __syncthreads();
    if (live) {
#pragma unroll 8
      for (int k = 0; k < m; ++k) {
        const float dx = __fsub_rn(px, sx[k]), dy = __fsub_rn(py, sy[k]), dz = __fsub_rn(pz, sz[k]);
        // ||diff||: sum of squares in index order, then sqrt (diff.norm(dim=-1), loss.py:60)
        const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        if (d < best) { best = d; bx = dx; by = dy; bz = dz; }
      }
    }
  }
  if (!live) return;
  const int64_t r = i / S;
  const bool valid = ray_valid ? (ray_valid[r] != 0) : true;
  if (!valid) {
    bounds[i] = 0.f;
    grad_vec[i * 3] = 0.f; grad_vec[i * 3 + 1] = 0.f; grad_vec[i * 3 + 2] = 0.f;
    return;
  }
  const bool behind = z_vals[i] > depth[r];
  bounds[i] = behind ? -best : best;
  // grad / grad.norm (0/0 -> NaN when the sample sits on its closest surface point), flipped behind the surface
  const float nn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz)));
  float gx = __fdiv_rn(bx, nn), gy = __fdiv_rn(by, nn), gz = __fdiv_rn(bz, nn);
  if (behind) { gx = -gx; gy = -gy; gz = -gz; }
  grad_vec[i * 3] = gx; grad_vec[i * 3 + 1] = gy; grad_vec[i * 3 + 2] = gz;
}

int bounds_pc_launch(isdfb_ctx* ctx, const float* pc, const float* z_vals, const float* depth, const uint8_t* ray_valid,
                     int64_t n_rays, int32_t S, float* bounds, float* grad_vec, cudaStream_t st) {
  const int64_t n = n_rays * S;
  const int64_t blocks = (n + BPC_THREADS - 1) / BPC_THREADS;
  if (blocks > 0x7fffffffLL) ISDFB_FAIL(ctx, ISDFB_ERR_CAPACITY, "isdfb_bounds_pc: too many samples (%lld)", (long long)n);
  bounds_pc_kernel<<<(unsigned)blocks, BPC_THREADS, 0, st>>>(pc, z_vals, depth, ray_valid, n_rays, S, bounds, grad_vec);
  ISDFB_LAUNCHED(ctx);
  ISDFB_CUDA_OK(ctx, cudaGetLastError());
  return ISDFB_OK;
}
