// Builds the bf16 (hi, lo) K-major operand images of every 256x256 weight unit, in both
// orientations, from the packed fp32 parameters.  Runs after isdfb_pack_weights and inside K6
// (isdfb_adamw); 4 images x 128 KB per unit, L2 resident for the whole step.
#include "tc_path.cuh"

// grid: (unit*2 + orient, 32 blocks of 256 threads) ; thread = (row n, k-group k8)
__global__ void tc_pack_kernel(const float* __restrict__ packed, TcUnitTable units, uint8_t* __restrict__ img) {
  const int uo = blockIdx.x;
  const int unit = uo >> 1, orient = uo & 1;
  const int idx = blockIdx.y * blockDim.x + threadIdx.x;      // 0 .. 256*32-1
  const int n = idx >> 5, k8 = idx & 31;
  const float* W = packed + units.u[unit].w_off;
  const int ld = units.u[unit].ld;
  const int ph = units.u[unit].perm_half;
  const int col0 = units.u[unit].col0;
  float x[8];
  if (ph > 0) {                     // input axis = embedding: gather the reference columns in internal order
#pragma unroll
    for (int i = 0; i < 8; ++i) {   // internal columns beyond the packed row (pure padding of the 512-wide halves) are zero
      const int nat = (orient == 0) ? pe_nat_col(col0 + k8 * 8 + i, ph) : pe_nat_col(col0 + n, ph);
      x[i] = (nat >= ld) ? 0.f
             : (orient == 0) ? W[(size_t)n * ld + nat]                            // B[n = out][k = in(internal)]
                             : W[(size_t)(k8 * 8 + i) * ld + nat];                // B[n = in(internal)][k = out]
    }
  } else if (orient == 0) {         // B[n = out][k = in]
    const float4 a = *reinterpret_cast<const float4*>(W + (size_t)n * ld + k8 * 8);
    const float4 b = *reinterpret_cast<const float4*>(W + (size_t)n * ld + k8 * 8 + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {                          // B[n = in][k = out]
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = W[(size_t)(k8 * 8 + i) * ld + n];
  }
  uint4 hi, lo;
  split8(x, hi, lo);
  uint8_t* base = img + (size_t)uo * 2 * TC_IMG_BYTES;
  const uint32_t off = kmajor_off_bytes(TC_H, n, k8 * 8);
  *reinterpret_cast<uint4*>(base + off) = hi;
  *reinterpret_cast<uint4*>(base + TC_IMG_BYTES + off) = lo;
}

int tc_repack(isdfb_ctx* ctx, cudaStream_t st) {
  if (ctx->cfg.precision == ISDFB_PREC_FP32) return ISDFB_OK;
  TcState* tc = reinterpret_cast<TcState*>(ctx->tc);
  if (!tc) ISDFB_FAIL(ctx, ISDFB_ERR_STATE, "tensor-core state missing");
  dim3 grid(tc->n_units * 2, 32);
  tc_pack_kernel<<<grid, 256, 0, st>>>(ctx->w_packed, tc->units, tc->w_img);
  ISDFB_LAUNCHED(ctx);
  ISDFB_CUDA_OK(ctx, cudaGetLastError());
  return ISDFB_OK;
}
