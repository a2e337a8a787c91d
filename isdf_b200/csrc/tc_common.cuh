// sm_100a primitives for the tensor-core path: mbarrier, bulk (TMA) copies, TMEM management,
// tcgen05.mma / tcgen05.ld wrappers, UMMA shared-memory and instruction descriptors, and the
// operand-image layouts shared by the chain kernel, the weight-gradient kernel and the packer.
#pragma once
#include "common.cuh"

#define TC_H 256                 // hidden width == padded embedding width supported by this path
#define TC_TILE 128              // points per tile (UMMA M)
#define TC_TILE_FLOATS (TC_H * TC_TILE)

// ---- layouts --------------------------------------------------------------------------------
// (1) K-major operand image of X[rows][K] (bf16), no swizzle: [K/8][rows][8] -> 8x8 core matrices
//     of 128 contiguous bytes;  SBO (next 8 rows) = 128 B,  LBO (next 8 K) = rows*16 B.
__host__ __device__ __forceinline__ uint32_t kmajor_off_bytes(uint32_t rows, uint32_t r, uint32_t k) {
  return (k >> 3) * rows * 16u + r * 16u + (k & 7u) * 2u;
}
// (2) per-tile fp32 side arrays ("aux"): [f/4][128 pts][4]  -> a thread (point) reads/writes 16 B,
//     a warp 512 contiguous bytes.
__host__ __device__ __forceinline__ uint32_t aux_off_floats(uint32_t f, uint32_t p) {
  return (f >> 2) * 512u + p * 4u + (f & 3u);
}
// (3) per-tile bf16 "dW layout": [p/16][f/8][16 pts][8 f] -> each 16-point slice (one UMMA K step
//     of the weight-gradient GEMM) is 8 KB contiguous and is an MN-major operand with
//     SBO (next 8 features) = 256 B, LBO (next 8 points) = 128 B.
__host__ __device__ __forceinline__ uint32_t dwl_off_bytes(uint32_t f, uint32_t p) {
  return (p >> 4) * 8192u + (f >> 3) * 256u + (p & 15u) * 16u + (f & 7u) * 2u;
}
#define TC_DWL_TILE_BYTES 65536u

// (4) INTERNAL embedding column order of the tensor-core path.  The reference lays the embedding out as
//     [x y z | sin(xb_0..) | sin(xb_0.. + pi/2)] (embedding.py:104-110), which puts the sin / cos "mates" of one
//     (direction, octave) pair `half` columns apart.  Inside the chain kernel a thread owns 8 consecutive
//     columns, so the columns are permuted to [sin_0 cos_0 sin_1 cos_1 ... | x y z | pad]: both mates, which
//     the PE Jacobian (S2) and its adjoint (S3) combine, are thread-local.  Only the weight images of the
//     two units fed by the embedding (tc_pack.cu) and the flush of their gradients (tc_dw.cu) see the
//     permutation; the packed fp32 parameters and gradients stay in the reference's order.
__host__ __device__ __forceinline__ int pe_nat_col(int c, int half) {     // internal column -> reference column
  if (half <= 0) return c;
  if (c < 2 * half) return 3 + (c >> 1) + (c & 1) * half;
  if (c < 2 * half + 3) return c - 2 * half;
  return c;
}

// ---- misc PTX ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin on the phase; a wait that lasts > ~2 s of SM clocks is a protocol bug -> trap (surfaces as a CUDA
// error on the host instead of hanging the GPU).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("isdf_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// global -> shared bulk copy (TMA, 1-D), completion signalled on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// bring a contiguous global range into L2 ahead of use (no registers, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all previously issued tcgen05.mma of this thread done -> arrive(1) on the mbarrier
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of fp32 -> 32 registers per thread (thread = lane/row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 8 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// sigma in [0,1] as unorm16 (abs error 7.6e-6; 0 and 1 exact): 8 values <-> 16 B.
// Conversions stay on the FMA/ALU pipes (magic-number rounding), not on the quarter-rate XU pipe.
__device__ __forceinline__ uint4 pack_unorm16x8(const float* s) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t a = __float_as_uint(fmaf(s[2 * i], 65535.f, 8388608.f));       // 2^23 + rint(65535 s)
    const uint32_t b = __float_as_uint(fmaf(s[2 * i + 1], 65535.f, 8388608.f));
    w[i] = __byte_perm(a, b, 0x5410);                                              // low halves of a, b
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void unpack_unorm16x8(const uint4& v, float* s) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  // (2^23 + n) as a float, then ONE fma: (2^23 + n) c - 2^23 c = round(n c) (2^23 c is exact).  c = the float just
  // below 1/65535, so 65535 c < 1 and no clamp is needed (sigma of the saturated regime decodes as 0.99999988
  // instead of 1: relative 1.2e-7 on a product, 1.2e-5 absolute on beta (1 - sigma) -- below the bf16 lo part)
  const float c = 1.525902e-05f;
  const float k = -8388608.f * c;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[2 * i] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7610)), c, k);
    s[2 * i + 1] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7632)), c, k);
  }
}
// warp-specialised register re-allocation (whole warpgroups = 4 consecutive warps): the helper warpgroup gives
// registers back, the epilogue warpgroups take them
template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// softplus(beta=100) and sigmoid(100 z), branch-free, three SFU ops (ex2 / lg2 / rcp approximations,
// relative error ~2^-22, far below the bf16x3 product error).  Stable form: t = exp(-|bz|) in (0,1],
// softplus = max(z,0) + log(1+t)/beta, sigmoid = (bz >= 0 ? 1 : t) / (1+t).  For bz > 20 this gives
// z + O(1e-11) and sigma == 1.0f, i.e. torch's threshold branch (fc_map.py:54) to fp32 rounding.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void softplus100_fast(float z, float& h, float& sig) {
  const float t = ex2_approx(fabsf(z) * -144.26950408889634f);      // exp(-|100 z|)
  const float u = 1.f + t;
  const float r = rcp_approx(u);
  sig = (z >= 0.f) ? r : t * r;
  h = fmaf(lg2_approx(u), 0.0069314718055994531f, fmaxf(z, 0.f));   // + ln(2)/100 * log2(1+t)
}

// ---- gradient accumulation -------------------------------------------------------------------------
// Single GPU: red.global.add into the packed gradient.  Data parallel with an exchange installed
// (isdfb_set_grad_exchange): `p` is a MULTICAST address and multimem.red adds the value into every rank's
// copy inside the NVSwitch -- the all-reduce of the reference-free data-parallel design is the flush itself.
__device__ __forceinline__ void grad_add(float* p, float v, int mc) {
  if (mc) asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
  else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void grad_add4(float* p, float a, float b, float c, float d, int mc) {
  if (mc) asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
  else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4   [16,30) LBO>>4   [32,46) SBO>>4   [46,48) version=1   [61,64) layout=0
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor for kind::f16: D=f32, A=B=bf16, M x N, optional MN-major operands
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- bf16 split -----------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// two fp32 -> packed bf16x2 (x0 in the low half), round-to-nearest-even, one instruction
__device__ __forceinline__ uint32_t cvt_bf16x2(float x0, float x1) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x1), "f"(x0));
  return d;
}
// eight fp32 -> 16 B of bf16 hi and 16 B of bf16 lo  (lo = bf16(x - float(hi)))
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = cvt_bf16x2(x[2 * i], x[2 * i + 1]);
    const float r0 = x[2 * i] - __uint_as_float(h[i] << 16);
    const float r1 = x[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u);
    l[i] = cvt_bf16x2(r0, r1);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ uint4 pack8_hi(const float* x) {
  return make_uint4(cvt_bf16x2(x[0], x[1]), cvt_bf16x2(x[2], x[3]), cvt_bf16x2(x[4], x[5]), cvt_bf16x2(x[6], x[7]));
}
__device__ __forceinline__ void unpack8(const uint4& v, float* x) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[2 * i] = __uint_as_float(w[i] << 16);
    x[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}

// ---- weight images (tc_pack.cu) ---------------------------------------------------------------------
// One "unit" = a 256x256 block of a layer's weight, stored as two bf16 K-major images (hi, lo) for
// each orientation:  fwd: B[n=out][k=in]  (Y = X W^T),  bwd: B[n=in][k=out]  (Y = X W).
#define TC_IMG_BYTES (TC_H * TC_H * 2)                 // 128 KB
struct TcUnit {
  int64_t w_off;      // offset of the fp32 [256][256] block in the packed parameter buffer
  int32_t ld;         // its leading dimension
  int32_t perm_half;  // > 0: the unit's input axis is the embedding -> internal column order (pe_nat_col)
  int32_t col0;       // embedding-fed units: first internal embedding column of this unit (0, or 256 for the second
                      // half when the padded embedding is 512 wide: n_embed_funcs 8 / 10 of the realsense configs)
};
#define TC_MAX_UNITS (ISDFB_MAX_HIDDEN_LAYERS + 3)   // layers + concat embedding part + second embedding halves
