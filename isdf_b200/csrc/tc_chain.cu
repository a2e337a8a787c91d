// K2/K3/K4 on tcgen05: one persistent kernel runs, for a 128-point tile, the whole chain
//   PE -> S1 (forward) -> S2 (input gradient) -> loss -> S3 -> S4   
// as a sequence of 128x256x256 UMMA products whose accumulators live in TMEM (two 256-column
// buffers, ping-pong) and whose A operand (the activations) stays in shared memory: the epilogue of
// step s reads D_s from TMEM, applies the element-wise stage (softplus / sigmoid products / loss),
// and writes the A operand of step s+1 in place, 64 columns at a time, so the MMA of step s+1 starts
// on K-chunk 0 while the epilogue is still producing chunks 1..3.  Weights stream from L2 through a
// 3-stage ring of bulk (TMA) copies of pre-packed operand images.
//
// Warp roles (192 threads): warps 0-3 = epilogue (thread = point = TMEM lane), warp 4 = MMA issuer
// (+TMEM alloc), warp 5 = weight producer.
//
// Precision: kPasses = 3 -> every product is A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with bf16 hi/lo splits
// and fp32 accumulation (~fp32 accuracy); kPasses = 1 -> single bf16 pass (fast mode).
#include "tc_chain.cuh"
#include "pe_loss.cuh"

#define EPI_THREADS 128
#define NUM_THREADS 192
#define K_CHUNK 32                       // K elements per weight-ring stage
#define N_KCHUNKS (TC_H / K_CHUNK)       // 8
#define A_IMG_BYTES (TC_TILE * TC_H * 2) // 64 KB
#define A_LBO (TC_TILE * 16)             // 2048
#define B_LBO (TC_H * 16)                // 4096
#define CHUNK_IMG_BYTES (TC_H * K_CHUNK * 2)   // 16 KB per precision part

template <int kPasses> struct ChainCfg {
  static constexpr int kStageBytes = CHUNK_IMG_BYTES * (kPasses == 3 ? 2 : 1);
  static constexpr int kStages = (kPasses == 3) ? 3 : 6;
  static constexpr int kABytes = A_IMG_BYTES * (kPasses == 3 ? 2 : 1);
  static constexpr int kSmem = kABytes + kStages * kStageBytes + 256;
};

struct ChainSmemTail {       // lives after the operand buffers
  uint64_t w_full[6], w_empty[6], a_ready[4], d_full[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// column -> PE feature decomposition
struct FeatInfo { int kind; int d; float scale; int mate; };   // kind 0: xs, 1: sin, 2: shifted-sin, 3: pad
__device__ __forceinline__ FeatInfo feat_info(int k, int F, int E) {
  FeatInfo fi;
  const int half = ISDFB_NDIRS * F;
  if (k < 3) { fi.kind = 0; fi.d = k; fi.scale = 1.f; fi.mate = k; return fi; }
  if (k >= E) { fi.kind = 3; fi.d = 0; fi.scale = 0.f; fi.mate = k; return fi; }
  int q = k - 3;
  if (q < half) { fi.kind = 1; fi.mate = k + half; } else { fi.kind = 2; q -= half; fi.mate = k - half; }
  fi.d = q / F;
  fi.scale = (float)(1 << (q - fi.d * F));
  return fi;
}

template <int kPasses>
__global__ void __launch_bounds__(NUM_THREADS, 1) tc_chain_kernel(const TcChainArgs args) {
  using Cfg = ChainCfg<kPasses>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + A_IMG_BYTES;                         // only valid when kPasses == 3
  uint8_t* w_ring = smem + Cfg::kABytes;
  ChainSmemTail* tail = reinterpret_cast<ChainSmemTail*>(smem + Cfg::kABytes + Cfg::kStages * Cfg::kStageBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_steps = args.n_steps;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(smem_u32(&tail->w_full[i]), 1); mbar_init(smem_u32(&tail->w_empty[i]), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&tail->a_ready[i]), 4);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&tail->d_full[i]), 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tail->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  const int my_tiles = (args.n_tiles > (int)blockIdx.x) ? (args.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 5) {
    // ===================== weight producer =====================
    if (elect_one()) {
      uint32_t j = 0;
      for (int it = 0; it < my_tiles; ++it) {
        for (int s = 0; s < n_steps; ++s) {
          const TcStep st = args.steps[s];
          const uint8_t* img_hi = args.w_img + ((size_t)(st.unit * 2 + st.orient) * 2 + 0) * TC_IMG_BYTES;
          const uint8_t* img_lo = img_hi + TC_IMG_BYTES;
          for (int kc = 0; kc < N_KCHUNKS; ++kc, ++j) {
            const uint32_t stage = j % Cfg::kStages, ph = (j / Cfg::kStages) & 1;
            mbar_wait(smem_u32(&tail->w_empty[stage]), ph ^ 1);
            const uint32_t bar = smem_u32(&tail->w_full[stage]);
            const uint32_t dst = smem_u32(w_ring + stage * Cfg::kStageBytes);
            mbar_arrive_expect_tx(bar, Cfg::kStageBytes);
            bulk_g2s(dst, img_hi + (size_t)kc * CHUNK_IMG_BYTES, CHUNK_IMG_BYTES, bar);
            if (kPasses == 3) bulk_g2s(dst + CHUNK_IMG_BYTES, img_lo + (size_t)kc * CHUNK_IMG_BYTES, CHUNK_IMG_BYTES, bar);
          }
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    uint32_t j = 0, n = 0;
    for (int it = 0; it < my_tiles; ++it) {
      for (int s = 0; s < n_steps; ++s, ++n) {
        const uint32_t d_tmem = tmem + (n & 1) * 256;
        for (int kc = 0; kc < N_KCHUNKS; ++kc, ++j) {
          if ((kc & 1) == 0) mbar_wait(smem_u32(&tail->a_ready[kc >> 1]), n & 1);
          const uint32_t stage = j % Cfg::kStages, ph = (j / Cfg::kStages) & 1;
          mbar_wait(smem_u32(&tail->w_full[stage]), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_base = smem_u32(w_ring + stage * Cfg::kStageBytes);
#pragma unroll
            for (int ks = 0; ks < K_CHUNK / 16; ++ks) {
              const uint32_t kg = kc * (K_CHUNK / 8) + ks * 2;
              const uint64_t ah = umma_desc(smem_u32(a_hi) + kg * A_LBO, A_LBO, 128);
              const uint64_t bh = umma_desc(b_base + ks * 2 * B_LBO, B_LBO, 128);
              tc_mma_f16(d_tmem, ah, bh, idesc, (kc | ks) != 0);
              if (kPasses == 3) {
                const uint64_t al = umma_desc(smem_u32(a_lo) + kg * A_LBO, A_LBO, 128);
                const uint64_t bl = umma_desc(b_base + CHUNK_IMG_BYTES + ks * 2 * B_LBO, B_LBO, 128);
                tc_mma_f16(d_tmem, al, bh, idesc, 1);
                tc_mma_f16(d_tmem, ah, bl, idesc, 1);
              }
            }
            tc_commit(smem_u32(&tail->w_empty[stage]));
            if (kc == N_KCHUNKS - 1) tc_commit(smem_u32(&tail->d_full[n & 1]));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue: thread = point =====================
    const int p = threadIdx.x;                                    // row / TMEM lane
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const float c_out = args.scale_output;
    const float* Wp = args.w_packed;
    const int L = args.L, ic = args.ic, F = args.pe.n_freqs, E = args.E;
    const bool train = args.mode == TC_MODE_TRAIN;
    const bool store_state = args.mode != TC_MODE_FWD;
    uint32_t n = 0;
    float lsum0 = 0.f, lsum1 = 0.f, lsum2 = 0.f, lsum3 = 0.f, sbsum = 0.f;

    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int64_t pl = (int64_t)tile * TC_TILE + p;             // point index inside the chunk
      const bool real = pl < args.n_points;
      const size_t tile_aux = (size_t)tile * TC_TILE_FLOATS;
      const size_t tile_dwl = (size_t)tile * TC_DWL_TILE_BYTES;
      auto aux_ptr = [&](int arr) { return args.aux + (size_t)arr * args.aux_stride + tile_aux; };
      auto dwl_hi = [&](int arr) { return args.dwl_hi + (size_t)arr * args.dwl_stride + tile_dwl; };
      auto dwl_lo = [&](int arr) { return args.dwl_lo + (size_t)arr * args.dwl_stride + tile_dwl; };
      // write 8 consecutive features [k0, k0+8) of this point: A operand image and/or a dW-layout copy
      auto put8 = [&](const float* x8, int k0, bool to_a, int dwl_arr) {
        uint4 hi, lo;
        split8(x8, hi, lo);
        if (to_a) {
          const uint32_t off = (uint32_t)(k0 >> 3) * A_LBO + (uint32_t)p * 16u;
          *reinterpret_cast<uint4*>(a_hi + off) = hi;
          if (kPasses == 3) *reinterpret_cast<uint4*>(a_lo + off) = lo;
        }
        if (dwl_arr >= 0) {
          const uint32_t off = dwl_off_bytes(k0, p);
          *reinterpret_cast<uint4*>(dwl_hi(dwl_arr) + off) = hi;
          if (kPasses == 3) *reinterpret_cast<uint4*>(dwl_lo(dwl_arr) + off) = lo;
        }
      };
      auto chunk_ready = [&](int chunk) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tail->a_ready[chunk]));
      };

      // ---------------- PE stage: x -> e (A operand of the first step) ----------------
      float xs[3] = {0.f, 0.f, 0.f};
      if (real) {
        const float* xp = args.x + pl * 3;
        pe_scale_input(args.pe, xp[0], xp[1], xp[2], xs);
      }
      float* e32 = aux_ptr(args.arr_e32);
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = real ? ((c * 32 + jj < E) ? pe_feature(args.pe, xs, c * 32 + jj) : 0.f) : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) put8(v + q * 8, c * 32 + q * 8, true, train ? args.arr_yh : -1);
        if (store_state) {
#pragma unroll
          for (int q = 0; q < 8; ++q) st4(e32 + aux_off_floats(c * 32 + q * 4, p), v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
        if (c & 1) chunk_ready(c >> 1);
      }

      // per-point state carried across steps
      float sdf_reg = 0.f, sbar = 0.f, u3[3] = {0.f, 0.f, 0.f};

      for (int s = 0; s < n_steps; ++s, ++n) {
        const TcStep st = args.steps[s];
        const int l = st.layer;
        const bool last_step = (s == n_steps - 1);
        const uint32_t d_tmem = tmem + (n & 1) * 256 + lane_addr;
        mbar_wait(smem_u32(&tail->d_full[n & 1]), (n >> 1) & 1);
        tc_fence_after();
        if (st.epi == EPI_RAW) {
          // A is left untouched: release the next step now (its MMA overlaps this drain of D into a
          // side array).  Arriving only after d_full guarantees every warp finished the previous phase.
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) mbar_arrive(smem_u32(&tail->a_ready[ch]));
          }
        }

        const float* bias = Wp + args.lay_b_off[l];
        const float* sig_in = aux_ptr(args.arr_sig + l);
        float* sig_out = aux_ptr(args.arr_sig + l);
        float* zb2 = aux_ptr(args.arr_zb2 + l);
        float raw_acc = 0.f;                      // S1_LAST: h . w_out
        float gx = 0.f, gy = 0.f, gz = 0.f;       // S2_END: PE back-projection accumulators

#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          float v[32];
          tmem_ld32(d_tmem + c * 32, v);
          const int k0 = c * 32;
          switch (st.epi) {
            case EPI_RAW: {
              float* dst = aux_ptr(args.arr_part + st.aux);
#pragma unroll
              for (int q = 0; q < 8; ++q) st4(dst + aux_off_floats(k0 + q * 4, p), v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
            } break;
            case EPI_S1:
            case EPI_S1_LAST: {
              const float* part = aux_ptr(args.arr_part + 0);
              const float* wout = Wp + args.wout_off;
              float* hl = aux_ptr(args.arr_hlast);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 b4 = ld4(bias + k0 + q * 4);
                float z[4] = {v[q * 4] + b4.x, v[q * 4 + 1] + b4.y, v[q * 4 + 2] + b4.z, v[q * 4 + 3] + b4.w};
                if (l == ic) {
                  float4 p4 = ld4(part + aux_off_floats(k0 + q * 4, p));
                  z[0] += p4.x; z[1] += p4.y; z[2] += p4.z; z[3] += p4.w;
                }
                float h[4], sg[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) softplus100(z[i], h[i], sg[i]);
                if (store_state) st4(sig_out + aux_off_floats(k0 + q * 4, p), sg[0], sg[1], sg[2], sg[3]);
                if (st.epi == EPI_S1) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) v[q * 4 + i] = h[i];
                } else {
                  float4 w4 = ld4(wout + k0 + q * 4);
                  const float ww[4] = {w4.x, w4.y, w4.z, w4.w};
                  if (train) st4(hl + aux_off_floats(k0 + q * 4, p), h[0], h[1], h[2], h[3]);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    raw_acc = fmaf(h[i], ww[i], raw_acc);
                    v[q * 4 + i] = c_out * ww[i] * sg[i];          // delta_{L-1} = a_{L-1} * sigma
                  }
                }
              }
              const int arr = (st.epi == EPI_S1) ? ((train && l + 1 < L) ? args.arr_yh + l + 1 : -1)
                                                 : (train ? args.arr_xd + l : -1);
              const bool to_a = !(st.epi == EPI_S1_LAST && args.mode == TC_MODE_FWD);
#pragma unroll
              for (int q = 0; q < 4; ++q) put8(v + q * 8, k0 + q * 8, to_a, arr);
            } break;
            case EPI_S2: {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 s4 = ld4(sig_in + aux_off_floats(k0 + q * 4, p));
                v[q * 4] *= s4.x; v[q * 4 + 1] *= s4.y; v[q * 4 + 2] *= s4.z; v[q * 4 + 3] *= s4.w;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) put8(v + q * 8, k0 + q * 8, true, train ? args.arr_xd + l : -1);
            } break;
            case EPI_S2_END: {
              const float* part = aux_ptr(args.arr_part + 1);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 p4 = ld4(part + aux_off_floats(k0 + q * 4, p));
                const float a[4] = {v[q * 4] + p4.x, v[q * 4 + 1] + p4.y, v[q * 4 + 2] + p4.z, v[q * 4 + 3] + p4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int k = k0 + q * 4 + i;
                  const FeatInfo fi = feat_info(k, F, E);
                  if (fi.kind == 0) {
                    if (k == 0) gx += a[i]; else if (k == 1) gy += a[i]; else gz += a[i];
                  } else if (fi.kind != 3) {
                    const float mate = e32[aux_off_floats(fi.mate, p)];
                    // d sin(xb)/d xb = "cos" = e[mate];  d sin(xb+pi/2)/d xb = -sin(xb) = -e[mate]
                    const float w = (fi.kind == 1 ? mate : -mate) * fi.scale * a[i];
                    gx = fmaf(w, c_ico[fi.d][0], gx);
                    gy = fmaf(w, c_ico[fi.d][1], gy);
                    gz = fmaf(w, c_ico[fi.d][2], gz);
                  }
                }
              }
            } break;
            case EPI_S3:
            case EPI_S3_LAST: {
              const float* part = aux_ptr(args.arr_part + 2);
              const uint8_t* dhi = dwl_hi(args.arr_xd + l);
              const uint8_t* dlo = dwl_lo(args.arr_xd + l);
              const float* wout = Wp + args.wout_off;
              const float* hl = aux_ptr(args.arr_hlast);
              float zb[32];
#pragma unroll
              for (int q8 = 0; q8 < 4; ++q8) {
                float dl[8];
                {
                  uint4 h8 = *reinterpret_cast<const uint4*>(dhi + dwl_off_bytes(k0 + q8 * 8, p));
                  unpack8(h8, dl);
                  if (kPasses == 3) {
                    float t8[8];
                    uint4 l8 = *reinterpret_cast<const uint4*>(dlo + dwl_off_bytes(k0 + q8 * 8, p));
                    unpack8(l8, t8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) dl[i] += t8[i];
                  }
                }
#pragma unroll
                for (int hq = 0; hq < 2; ++hq) {
                  const int q = q8 * 2 + hq;
                  float4 s4 = ld4(sig_in + aux_off_floats(k0 + q * 4, p));
                  const float sg[4] = {s4.x, s4.y, s4.z, s4.w};
                  float db[4] = {v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]};
                  if (l == ic) {
                    float4 p4 = ld4(part + aux_off_floats(k0 + q * 4, p));
                    db[0] += p4.x; db[1] += p4.y; db[2] += p4.z; db[3] += p4.w;
                  }
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float om = (sg[i] >= 1.f) ? 0.f : 100.f * (1.f - sg[i]);
                    zb[q * 4 + i] = db[i] * dl[hq * 4 + i] * om;      // zbar2 = dbar * delta * beta (1 - sigma)
                    v[q * 4 + i] = db[i] * sg[i];                      // abar = dbar * sigma
                  }
                }
              }
              if (st.epi == EPI_S3) {
#pragma unroll
                for (int q = 0; q < 8; ++q) st4(zb2 + aux_off_floats(k0 + q * 4, p), zb[q * 4], zb[q * 4 + 1], zb[q * 4 + 2], zb[q * 4 + 3]);
#pragma unroll
                for (int q = 0; q < 4; ++q) put8(v + q * 8, k0 + q * 8, true, (l + 1 < L) ? args.arr_ya + l + 1 : -1);
              } else {
                // v_blob = sbar * h_last + abar_last  (for d w_out);  A <- zbar_last = sbar c w_out sigma + zbar2
                float vb[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  float4 h4 = ld4(hl + aux_off_floats(k0 + q * 4, p));
                  float4 s4 = ld4(sig_in + aux_off_floats(k0 + q * 4, p));
                  float4 w4 = ld4(wout + k0 + q * 4);
                  const float hh[4] = {h4.x, h4.y, h4.z, h4.w}, sg[4] = {s4.x, s4.y, s4.z, s4.w}, ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    vb[q * 4 + i] = fmaf(sbar, hh[i], v[q * 4 + i]);
                    v[q * 4 + i] = fmaf(sbar * c_out * ww[i], sg[i], zb[q * 4 + i]);
                  }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) put8(vb + q * 8, k0 + q * 8, false, args.arr_v);
#pragma unroll
                for (int q = 0; q < 4; ++q) put8(v + q * 8, k0 + q * 8, true, args.arr_xz + l);
              }
            } break;
            case EPI_S4: {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 s4 = ld4(sig_in + aux_off_floats(k0 + q * 4, p));
                float4 z4 = ld4(zb2 + aux_off_floats(k0 + q * 4, p));
                v[q * 4] = fmaf(v[q * 4], s4.x, z4.x);
                v[q * 4 + 1] = fmaf(v[q * 4 + 1], s4.y, z4.y);
                v[q * 4 + 2] = fmaf(v[q * 4 + 2], s4.z, z4.z);
                v[q * 4 + 3] = fmaf(v[q * 4 + 3], s4.w, z4.w);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) put8(v + q * 8, k0 + q * 8, !last_step, args.arr_xz + l);
            } break;
          }
          if ((c & 1) && st.epi != EPI_RAW && st.epi != EPI_S2_END && !last_step) chunk_ready(c >> 1);
        }
        tc_fence_before();

        if (st.epi == EPI_S1_LAST) {
          float raw = raw_acc + Wp[args.bout_off];
          if (args.noise && real) raw += args.noise[pl] * args.noise_std;
          sdf_reg = raw * c_out;
          if (real) args.sdf_out[pl] = sdf_reg;
        } else if (st.epi == EPI_S2_END) {
          // g = s R^T g_xs
          float ox = gx, oy = gy, oz = gz;
          if (args.pe.has_transform) {
            ox = args.pe.R[0] * gx + args.pe.R[3] * gy + args.pe.R[6] * gz;
            oy = args.pe.R[1] * gx + args.pe.R[4] * gy + args.pe.R[7] * gz;
            oz = args.pe.R[2] * gx + args.pe.R[5] * gy + args.pe.R[8] * gz;
          }
          float g[3] = {args.pe.scale * ox, args.pe.scale * oy, args.pe.scale * oz};
          if (real && args.g_out) { args.g_out[pl * 3] = g[0]; args.g_out[pl * 3 + 1] = g[1]; args.g_out[pl * 3 + 2] = g[2]; }
          sbar = 0.f;
          float gb[3] = {0.f, 0.f, 0.f};
          if (train) {
            float tot = 0.f;
            if (real) {
              const int64_t pg = args.p0 + pl;                 // global sample index = r*S + j
              const int64_t r = pg / args.S;
              const int jx = (int)(pg - r * args.S);
              const bool valid = args.ray_valid ? (args.ray_valid[r] != 0) : true;
              if (valid) {
                const float dc[3] = {args.dirs_C[r * 3], args.dirs_C[r * 3 + 1], args.dirs_C[r * 3 + 2]};
                const float nrm = sqrtf(dc[0] * dc[0] + dc[1] * dc[1] + dc[2] * dc[2]);
                const float bnd = nrm * (args.depth[r] - args.z_vals[pg]);
                float uu[3];
                if (jx == 0 && args.normals) {
                  uu[0] = args.normals[r * 3]; uu[1] = args.normals[r * 3 + 1]; uu[2] = args.normals[r * 3 + 2];
                } else {
                  const float* T = args.T_WC + r * 16;
                  uu[0] = -(T[0] * dc[0] + T[1] * dc[1] + T[2] * dc[2]);
                  uu[1] = -(T[4] * dc[0] + T[5] * dc[1] + T[6] * dc[2]);
                  uu[2] = -(T[8] * dc[0] + T[9] * dc[1] + T[10] * dc[2]);
                }
                const LossPoint o = loss_point(args.loss, sdf_reg, g, bnd, uu);
                sbar = o.sbar; gb[0] = o.gbar[0]; gb[1] = o.gbar[1]; gb[2] = o.gbar[2];
                tot = o.total;
                lsum0 += o.l_sdf; lsum1 += o.l_grad; lsum2 += o.l_eik; lsum3 += o.total;
                sbsum += o.sbar;
              }
              args.loss_mat[pg] = tot;
            }
            // u = s R gbar ; abar_e -> A operand of S3 (second pass over the columns)
            u3[0] = gb[0]; u3[1] = gb[1]; u3[2] = gb[2];
            if (args.pe.has_transform) {
              u3[0] = args.pe.R[0] * gb[0] + args.pe.R[1] * gb[1] + args.pe.R[2] * gb[2];
              u3[1] = args.pe.R[3] * gb[0] + args.pe.R[4] * gb[1] + args.pe.R[5] * gb[2];
              u3[2] = args.pe.R[6] * gb[0] + args.pe.R[7] * gb[1] + args.pe.R[8] * gb[2];
            }
            u3[0] *= args.pe.scale; u3[1] *= args.pe.scale; u3[2] *= args.pe.scale;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              float v[32];
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) {
                const int k = c * 32 + jj;
                const FeatInfo fi = feat_info(k, F, E);
                float val = 0.f;
                if (fi.kind == 0) {
                  val = u3[k];
                } else if (fi.kind != 3) {
                  const float ud = (u3[0] * c_ico[fi.d][0] + u3[1] * c_ico[fi.d][1] + u3[2] * c_ico[fi.d][2]) * fi.scale;
                  const float mate = e32[aux_off_floats(fi.mate, p)];
                  val = (fi.kind == 1) ? ud * mate : -ud * mate;
                }
                v[jj] = val;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) put8(v + q * 8, c * 32 + q * 8, true, args.arr_ya);
              if (c & 1) chunk_ready(c >> 1);
            }
          }
        }
      }
    }
    // loss sums: warp reduce, one atomic per warp
    if (args.mode == TC_MODE_TRAIN) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lsum0 += __shfl_xor_sync(0xffffffffu, lsum0, o);
        lsum1 += __shfl_xor_sync(0xffffffffu, lsum1, o);
        lsum2 += __shfl_xor_sync(0xffffffffu, lsum2, o);
        lsum3 += __shfl_xor_sync(0xffffffffu, lsum3, o);
        sbsum += __shfl_xor_sync(0xffffffffu, sbsum, o);
      }
      if (lane == 0) {
        atomicAdd(args.loss_sums + 0, lsum0);
        atomicAdd(args.loss_sums + 1, lsum1);
        atomicAdd(args.loss_sums + 2, lsum2);
        atomicAdd(args.loss_sums + 3, lsum3);
        atomicAdd(args.g_packed + args.bout_off, c_out * sbsum);     // d b_out = c * sum sbar
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 512);
}

int tc_chain_launch(isdfb_ctx* ctx, const TcChainArgs& args, int passes, int grid, cudaStream_t st) {
  if (passes == 3) {
    ISDFB_CUDA_OK(ctx, cudaFuncSetAttribute(tc_chain_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg<3>::kSmem));
    tc_chain_kernel<3><<<grid, NUM_THREADS, ChainCfg<3>::kSmem, st>>>(args);
  } else {
    ISDFB_CUDA_OK(ctx, cudaFuncSetAttribute(tc_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg<1>::kSmem));
    tc_chain_kernel<1><<<grid, NUM_THREADS, ChainCfg<1>::kSmem, st>>>(args);
  }
  ISDFB_LAUNCHED(ctx);
  ISDFB_CUDA_OK(ctx, cudaGetLastError());
  return ISDFB_OK;
}
