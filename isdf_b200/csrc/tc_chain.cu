// K2/K3/K4 on tcgen05: one persistent kernel runs, for a 128-point tile, the whole chain
//   PE -> S1 (forward) -> S2 (input gradient) -> loss -> S3 -> S4   
// as a sequence of 128x256x256 UMMA products whose accumulators live in TMEM (two 256-column
// buffers, ping-pong) and whose A operand (the activations) stays in shared memory: the epilogue of
// step s reads D_s from TMEM, applies the element-wise stage (softplus / sigmoid products / loss),
// and writes the A operand of step s+1 in place.  Weights stream from L2 through a ring of bulk
// (TMA) copies of pre-packed operand images, one UMMA K-step (16) per stage.
//
// Warp roles (576 threads):
//   warps 0-15  epilogue: warp w owns TMEM lane quadrant w%4 (points 32(w%4)..+31) and 16 of the 64
//               columns of every K chunk of the next step's A operand; 4 warps per scheduler hide the
//               ALU/SFU/LSU latency of the element-wise math; chunk 0 of the next A operand is done
//               after a quarter of the epilogue, so the next MMA runs under the rest of it;
//   warp 16     MMA issuer (+ TMEM alloc);      warp 17   weight producer (+ L2 prefetch of side arrays).
//
// Precision: kPasses = 3 -> every product is A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with bf16 hi/lo splits
// and fp32 accumulation (~fp32 accuracy); kPasses = 1 -> single bf16 pass (fast mode).
#include "tc_chain.cuh"
#include "pe_loss.cuh"

#define EPI_WARPS 16
#define EPI_THREADS (EPI_WARPS * 32)
#define NUM_THREADS (EPI_THREADS + 64)
#define K_STEP 16                        // K elements per weight-ring stage (one UMMA K step)
#define N_KSTEPS (TC_H / K_STEP)         // 16
#define A_IMG_BYTES (TC_TILE * TC_H * 2) // 64 KB
#define A_LBO (TC_TILE * 16)             // 2048
#define B_LBO (TC_H * 16)                // 4096
#define KSTEP_IMG_BYTES (TC_H * K_STEP * 2)   // 8 KB per precision part
#define SCRATCH_BYTES 8192

template <int kPasses> struct ChainCfg {
  static constexpr int kStageBytes = KSTEP_IMG_BYTES * (kPasses == 3 ? 2 : 1);
  static constexpr int kStages = (kPasses == 3) ? 5 : 8;
  static constexpr int kABytes = A_IMG_BYTES * (kPasses == 3 ? 2 : 1);
  static constexpr int kSmem = kABytes + kStages * kStageBytes + SCRATCH_BYTES + 256;
};

struct ChainSmemTail {       // lives after the operand buffers and the scratch
  uint64_t w_full[8], w_empty[8], a_ready[4], d_full[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

template <int kPasses>
__global__ void __launch_bounds__(NUM_THREADS, 1) tc_chain_kernel(const __grid_constant__ TcChainArgs args) {
  using Cfg = ChainCfg<kPasses>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_hi = smem;
  uint8_t* a_lo = smem + A_IMG_BYTES;                         // only valid when kPasses == 3
  uint8_t* w_ring = smem + Cfg::kABytes;
  float* red = reinterpret_cast<float*>(smem + Cfg::kABytes + Cfg::kStages * Cfg::kStageBytes);   // [3][4][128]
  float* bcast = red + 3 * 4 * 128;                                                               // [128][4]
  ChainSmemTail* tail = reinterpret_cast<ChainSmemTail*>(reinterpret_cast<uint8_t*>(red) + SCRATCH_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_steps = args.n_steps;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(smem_u32(&tail->w_full[i]), 1); mbar_init(smem_u32(&tail->w_empty[i]), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&tail->a_ready[i]), EPI_WARPS);
    for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&tail->d_full[i]), 1);
    mbar_fence_init();
  }
  if (warp == EPI_WARPS) tmem_alloc(smem_u32(&tail->tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  const int my_tiles = (args.n_tiles > (int)blockIdx.x) ? (args.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == EPI_WARPS + 1) {
    // ===================== weight producer =====================
    if (elect_one()) {
      uint32_t j = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        for (int s = 0; s < n_steps; ++s) {
          const TcStep st = args.steps[s];
          if (args.prefetch && s + 1 < n_steps) {
            // pull the side arrays the NEXT step's epilogue will read from HBM into L2 while this step runs
            const TcStep nx = args.steps[s + 1];
            const float* aux_t = args.aux + (size_t)tile * TC_TILE_FLOATS;
            const size_t dwl_t = (size_t)tile * TC_DWL_TILE_BYTES;
            auto pf_aux = [&](int arr) { bulk_prefetch_l2(aux_t + (size_t)arr * args.aux_stride, TC_TILE_FLOATS * 4); };
            auto pf_sig = [&](int l_) { bulk_prefetch_l2(args.sig16 + (size_t)l_ * args.sig16_stride + dwl_t, TC_DWL_TILE_BYTES); };
            auto pf_dwl = [&](int arr) {
              bulk_prefetch_l2(args.dwl_hi + (size_t)arr * args.dwl_stride + dwl_t, TC_DWL_TILE_BYTES);
              if (kPasses == 3) bulk_prefetch_l2(args.dwl_lo + (size_t)arr * args.dwl_stride + dwl_t, TC_DWL_TILE_BYTES);
            };
            switch (nx.epi) {
              case EPI_S1: case EPI_S1_LAST: if (nx.layer == args.ic) pf_aux(args.arr_part + 0); break;
              case EPI_S2: pf_sig(nx.layer); break;
              case EPI_S2_END: pf_aux(args.arr_part + 1); pf_aux(args.arr_e32); break;
              case EPI_S3: case EPI_S3_LAST:
                pf_sig(nx.layer); pf_dwl(args.arr_xd + nx.layer);
                if (nx.layer == args.ic) pf_aux(args.arr_part + 2);
                if (nx.epi == EPI_S3_LAST) pf_aux(args.arr_hlast);
                break;
              case EPI_S4: pf_sig(nx.layer); pf_aux(args.arr_zb2 + nx.layer); break;
              default: break;
            }
          }
          const uint8_t* img_hi = args.w_img + ((size_t)(st.unit * 2 + st.orient) * 2 + 0) * TC_IMG_BYTES;
          const uint8_t* img_lo = img_hi + TC_IMG_BYTES;
          for (int ks = 0; ks < N_KSTEPS; ++ks, ++j) {
            const uint32_t stage = j % Cfg::kStages, ph = (j / Cfg::kStages) & 1;
            mbar_wait(smem_u32(&tail->w_empty[stage]), ph ^ 1);
            const uint32_t bar = smem_u32(&tail->w_full[stage]);
            const uint32_t dst = smem_u32(w_ring + stage * Cfg::kStageBytes);
            mbar_arrive_expect_tx(bar, Cfg::kStageBytes);
            bulk_g2s(dst, img_hi + (size_t)ks * KSTEP_IMG_BYTES, KSTEP_IMG_BYTES, bar);
            if (kPasses == 3) bulk_g2s(dst + KSTEP_IMG_BYTES, img_lo + (size_t)ks * KSTEP_IMG_BYTES, KSTEP_IMG_BYTES, bar);
          }
        }
      }
    }
  } else if (warp == EPI_WARPS) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
    uint32_t j = 0, n = 0;
    for (int it = 0; it < my_tiles; ++it) {
      for (int s = 0; s < n_steps; ++s, ++n) {
        const uint32_t d_tmem = tmem + (n & 1) * 256;
        for (int ks = 0; ks < N_KSTEPS; ++ks, ++j) {
          if ((ks & 3) == 0) mbar_wait(smem_u32(&tail->a_ready[ks >> 2]), n & 1);
          const uint32_t stage = j % Cfg::kStages, ph = (j / Cfg::kStages) & 1;
          mbar_wait(smem_u32(&tail->w_full[stage]), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_base = smem_u32(w_ring + stage * Cfg::kStageBytes);
            const uint64_t ah = umma_desc(smem_u32(a_hi) + ks * 2 * A_LBO, A_LBO, 128);
            const uint64_t bh = umma_desc(b_base, B_LBO, 128);
            tc_mma_f16(d_tmem, ah, bh, idesc, ks != 0);
            if (kPasses == 3) {
              const uint64_t al = umma_desc(smem_u32(a_lo) + ks * 2 * A_LBO, A_LBO, 128);
              const uint64_t bl = umma_desc(b_base + KSTEP_IMG_BYTES, B_LBO, 128);
              tc_mma_f16(d_tmem, al, bh, idesc, 1);
              tc_mma_f16(d_tmem, ah, bl, idesc, 1);
            }
            tc_commit(smem_u32(&tail->w_empty[stage]));
            if (ks == N_KSTEPS - 1) tc_commit(smem_u32(&tail->d_full[n & 1]));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue: thread = (point, 4 x 16 columns) =====================
    // Warp (q, j) owns points 32q..32q+31 and, inside every 64-column K chunk c of the next A operand,
    // columns [64c+16j, 64c+16j+16): chunk 0 is complete after a quarter of the epilogue, so the MMA of
    // the next step runs under the rest of it.  Work unit = 8 columns ("sub-piece"); the side-array
    // loads of sub-piece i+1 are issued before the stores of sub-piece i (software prefetch).
    const int q = warp & 3, jg = warp >> 2;
    const int p = q * 32 + lane;                                  // row / TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float c_out = args.scale_output;
    const float* Wp = args.w_packed;
    const int L = args.L, ic = args.ic, E = args.E;
    const int half = ISDFB_NDIRS * args.pe.n_freqs;
    const bool train = args.mode == TC_MODE_TRAIN;
    const bool store_state = args.mode != TC_MODE_FWD;
    uint32_t n = 0;
    float lsum0 = 0.f, lsum1 = 0.f, lsum2 = 0.f, lsum3 = 0.f, sbsum = 0.f;
    auto col_of = [&](int i) { return (i >> 1) * 64 + jg * 16 + (i & 1) * 8; };

    for (int it = 0; it < my_tiles; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int64_t pl = (int64_t)tile * TC_TILE + p;             // point index inside the chunk
      const bool real = pl < args.n_points;
      const size_t tile_aux = (size_t)tile * TC_TILE_FLOATS;
      const size_t tile_dwl = (size_t)tile * TC_DWL_TILE_BYTES;
      auto aux_ptr = [&](int arr) { return args.aux + (size_t)arr * args.aux_stride + tile_aux; };
      auto sig_ptr = [&](int l_) { return args.sig16 + (size_t)l_ * args.sig16_stride + tile_dwl; };
      // write 8 consecutive features [k0, k0+8) of this point: A operand image and/or a dW-layout copy
      auto put8 = [&](const float* x, int k0, bool to_a, int dwl_arr) {
        uint4 hi, lo;
        split8(x, hi, lo);
        if (to_a) {
          const uint32_t off = (uint32_t)(k0 >> 3) * A_LBO + (uint32_t)p * 16u;
          *reinterpret_cast<uint4*>(a_hi + off) = hi;
          if (kPasses == 3) *reinterpret_cast<uint4*>(a_lo + off) = lo;
        }
        if (dwl_arr >= 0) {
          const size_t off = (size_t)dwl_arr * args.dwl_stride + tile_dwl + dwl_off_bytes(k0, p);
          *reinterpret_cast<uint4*>(args.dwl_hi + off) = hi;
          if (kPasses == 3) *reinterpret_cast<uint4*>(args.dwl_lo + off) = lo;
        }
      };
      auto chunk_ready = [&](int c) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tail->a_ready[c]));
      };

      // ---------------- PE stage: x -> e (A operand of the first step) ----------------
      float xs[3] = {0.f, 0.f, 0.f};
      if (real) {
        const float* xp = args.x + pl * 3;
        pe_scale_input(args.pe, xp[0], xp[1], xp[2], xs);
      }
      float* e32 = aux_ptr(args.arr_e32);
#pragma unroll 1
      for (int i = 0; i < 8; ++i) {
        const int k0 = col_of(i);
        float v[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int k = k0 + jj;
          float val = 0.f;
          if (real && k < E) {
            if (k < 3) {
              val = (k == 0) ? xs[0] : ((k == 1) ? xs[1] : xs[2]);
            } else {
              const float xb = pe_project(xs, args.feat_d[k]) * (float)(1 << args.feat_f[k]);
              val = (k - 3 < half) ? sinf(xb) : sinf(__fadd_rn(xb, ISDFB_HALF_PI_F));
            }
          }
          v[jj] = val;
        }
        put8(v, k0, true, train ? args.arr_yh : -1);
        if (store_state) {
          st4(e32 + aux_off_floats(k0, p), v[0], v[1], v[2], v[3]);
          st4(e32 + aux_off_floats(k0 + 4, p), v[4], v[5], v[6], v[7]);
        }
        if (i & 1) chunk_ready(i >> 1);
      }

      // per-point state carried across steps
      float sdf_reg = 0.f, sbar = 0.f, u3[3] = {0.f, 0.f, 0.f};

      for (int s = 0; s < n_steps; ++s, ++n) {
        const TcStep st = args.steps[s];
        const int l = st.layer;
        const int epi = st.epi;
        const bool last_step = (s == n_steps - 1);
        const uint32_t d_tmem = tmem + (n & 1) * 256 + lane_addr;
        const bool use_sig = (epi == EPI_S2 || epi == EPI_S3 || epi == EPI_S3_LAST || epi == EPI_S4);
        const bool use_delta = (epi == EPI_S3 || epi == EPI_S3_LAST);
        const uint8_t* sigp = sig_ptr(l);
        float* zb2 = aux_ptr(args.arr_zb2 + l);
        const float* part_in = aux_ptr(args.arr_part + (epi == EPI_S2_END ? 1 : 0));
        const size_t doff = (size_t)(args.arr_xd + l) * args.dwl_stride + tile_dwl;
        const bool use_part01 = (epi == EPI_S2_END) || ((epi == EPI_S1 || epi == EPI_S1_LAST) && l == ic);
        // side-array operands of one sub-piece (12 registers), fetched one sub-piece ahead
        uint4 nx_s = make_uint4(0, 0, 0, 0), nx_b0 = nx_s, nx_b1 = nx_s;
        auto prefetch = [&](int k0) {
          if (use_sig) nx_s = *reinterpret_cast<const uint4*>(sigp + (uint32_t)(k0 >> 3) * 2048u + (uint32_t)p * 16u);
          if (use_delta) {
            const uint32_t o = dwl_off_bytes(k0, p);
            nx_b0 = *reinterpret_cast<const uint4*>(args.dwl_hi + doff + o);
            if (kPasses == 3) nx_b1 = *reinterpret_cast<const uint4*>(args.dwl_lo + doff + o);
          } else if (epi == EPI_S4) {
            nx_b0 = *reinterpret_cast<const uint4*>(zb2 + aux_off_floats(k0, p));
            nx_b1 = *reinterpret_cast<const uint4*>(zb2 + aux_off_floats(k0 + 4, p));
          } else if (use_part01) {
            nx_b0 = *reinterpret_cast<const uint4*>(part_in + aux_off_floats(k0, p));
            nx_b1 = *reinterpret_cast<const uint4*>(part_in + aux_off_floats(k0 + 4, p));
          }
        };
        prefetch(col_of(0));                      // overlaps the tail of this step's MMA
        mbar_wait(smem_u32(&tail->d_full[n & 1]), (n >> 1) & 1);
        tc_fence_after();
        if (epi == EPI_RAW) {
          // A is left untouched: release the next step now (its MMA overlaps this drain of D into a
          // side array).  Arriving only after d_full guarantees every warp finished the previous phase.
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) mbar_arrive(smem_u32(&tail->a_ready[c]));
          }
        }

        const float* bias = Wp + args.lay_b_off[l];
        float raw_acc = 0.f;                      // S1_LAST: h . w_out (this thread's 64 columns)
        float gx = 0.f, gy = 0.f, gz = 0.f;       // S2_END: PE back-projection accumulators

#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int k0 = col_of(i);
          const uint4 cs = nx_s, cb0 = nx_b0, cb1 = nx_b1;
          if (i + 1 < 8) prefetch(col_of(i + 1));
          float v[8];
          tmem_ld8(d_tmem + k0, v);
          if (epi == EPI_RAW) {
            float* dst = aux_ptr(args.arr_part + st.aux);
            st4(dst + aux_off_floats(k0, p), v[0], v[1], v[2], v[3]);
            st4(dst + aux_off_floats(k0 + 4, p), v[4], v[5], v[6], v[7]);
          } else if (epi == EPI_S1 || epi == EPI_S1_LAST) {
            const float4 ba = ld4(bias + k0), bb = ld4(bias + k0 + 4);
            float z[8] = {v[0] + ba.x, v[1] + ba.y, v[2] + ba.z, v[3] + ba.w, v[4] + bb.x, v[5] + bb.y, v[6] + bb.z, v[7] + bb.w};
            if (l == ic) {
              z[0] += __uint_as_float(cb0.x); z[1] += __uint_as_float(cb0.y); z[2] += __uint_as_float(cb0.z); z[3] += __uint_as_float(cb0.w);
              z[4] += __uint_as_float(cb1.x); z[5] += __uint_as_float(cb1.y); z[6] += __uint_as_float(cb1.z); z[7] += __uint_as_float(cb1.w);
            }
            float h[8], sg[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) softplus100_fast(z[t], h[t], sg[t]);
            if (store_state)
              *reinterpret_cast<uint4*>(args.sig16 + (size_t)l * args.sig16_stride + tile_dwl + (uint32_t)(k0 >> 3) * 2048u + (uint32_t)p * 16u) = pack_unorm16x8(sg);
            if (epi == EPI_S1) {
              put8(h, k0, true, (train && l + 1 < L) ? args.arr_yh + l + 1 : -1);
            } else {
              const float* wout = Wp + args.wout_off;
              const float4 wa = ld4(wout + k0), wb = ld4(wout + k0 + 4);
              const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
              if (train) {
                float* hl = aux_ptr(args.arr_hlast);
                st4(hl + aux_off_floats(k0, p), h[0], h[1], h[2], h[3]);
                st4(hl + aux_off_floats(k0 + 4, p), h[4], h[5], h[6], h[7]);
              }
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                raw_acc = fmaf(h[t], ww[t], raw_acc);
                // delta_{L-1} = a_{L-1} * sigma, with sigma as the later sweeps will read it back
                v[t] = c_out * ww[t] * sg[t];
              }
              put8(v, k0, args.mode != TC_MODE_FWD, train ? args.arr_xd + l : -1);
            }
          } else if (epi == EPI_S2) {
            float sg[8];
            unpack_unorm16x8(cs, sg);
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] *= sg[t];
            put8(v, k0, true, train ? args.arr_xd + l : -1);
          } else if (epi == EPI_S2_END) {
            const float a[8] = {v[0] + __uint_as_float(cb0.x), v[1] + __uint_as_float(cb0.y), v[2] + __uint_as_float(cb0.z), v[3] + __uint_as_float(cb0.w),
                                v[4] + __uint_as_float(cb1.x), v[5] + __uint_as_float(cb1.y), v[6] + __uint_as_float(cb1.z), v[7] + __uint_as_float(cb1.w)};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const int k = k0 + t;
              if (k < 3) {
                if (k == 0) gx += a[t]; else if (k == 1) gy += a[t]; else gz += a[t];
              } else if (k < E) {
                const bool first = (k - 3) < half;
                const float mate = e32[aux_off_floats(first ? k + half : k - half, p)];
                // d sin(xb)/d xb = "cos" = e[mate];  d sin(xb+pi/2)/d xb = -sin(xb) = -e[mate]
                const float w = (first ? mate : -mate) * (float)(1 << args.feat_f[k]) * a[t];
                const int d = args.feat_d[k];
                gx = fmaf(w, c_ico[d][0], gx);
                gy = fmaf(w, c_ico[d][1], gy);
                gz = fmaf(w, c_ico[d][2], gz);
              }
            }
          } else if (epi == EPI_S3 || epi == EPI_S3_LAST) {
            float sg[8], dl[8], zb[8];
            unpack_unorm16x8(cs, sg);
            unpack8(cb0, dl);
            if (kPasses == 3) {
              float t8[8];
              unpack8(cb1, t8);
#pragma unroll
              for (int t = 0; t < 8; ++t) dl[t] += t8[t];
            }
            if (l == ic) {
              const float* part = aux_ptr(args.arr_part + 2);
              const float4 pa = ld4(part + aux_off_floats(k0, p)), pb = ld4(part + aux_off_floats(k0 + 4, p));
              v[0] += pa.x; v[1] += pa.y; v[2] += pa.z; v[3] += pa.w; v[4] += pb.x; v[5] += pb.y; v[6] += pb.z; v[7] += pb.w;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float om = (sg[t] >= 1.f) ? 0.f : 100.f * (1.f - sg[t]);
              zb[t] = v[t] * dl[t] * om;           // zbar2 = dbar * delta * beta (1 - sigma)
              v[t] = v[t] * sg[t];                 // abar = dbar * sigma
            }
            if (epi == EPI_S3) {
              st4(zb2 + aux_off_floats(k0, p), zb[0], zb[1], zb[2], zb[3]);
              st4(zb2 + aux_off_floats(k0 + 4, p), zb[4], zb[5], zb[6], zb[7]);
              put8(v, k0, true, (l + 1 < L) ? args.arr_ya + l + 1 : -1);
            } else {
              // v_blob = sbar * h_last + abar_last  (for d w_out);  A <- zbar_last = sbar c w_out sigma + zbar2
              const float* wout = Wp + args.wout_off;
              const float* hl = aux_ptr(args.arr_hlast);
              const float4 ha = ld4(hl + aux_off_floats(k0, p)), hb = ld4(hl + aux_off_floats(k0 + 4, p));
              const float4 wa = ld4(wout + k0), wb = ld4(wout + k0 + 4);
              const float hh[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
              const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
              float vb[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                vb[t] = fmaf(sbar, hh[t], v[t]);
                zb[t] = fmaf(sbar * c_out * ww[t], sg[t], zb[t]);
              }
              put8(vb, k0, false, args.arr_v);
              put8(zb, k0, true, args.arr_xz + l);
            }
          } else {   // EPI_S4
            float sg[8];
            unpack_unorm16x8(cs, sg);
            const float z2[8] = {__uint_as_float(cb0.x), __uint_as_float(cb0.y), __uint_as_float(cb0.z), __uint_as_float(cb0.w),
                                 __uint_as_float(cb1.x), __uint_as_float(cb1.y), __uint_as_float(cb1.z), __uint_as_float(cb1.w)};
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = fmaf(v[t], sg[t], z2[t]);
            put8(v, k0, !last_step, args.arr_xz + l);
          }
          if ((i & 1) && epi != EPI_RAW && epi != EPI_S2_END && !last_step) chunk_ready(i >> 1);
        }
        tc_fence_before();

        if (epi == EPI_S1_LAST) {
          // out layer: combine the four column groups of every point
          red[jg * 128 + p] = raw_acc;
          named_bar_sync(1, EPI_THREADS);
          if (jg == 0) {
            float raw = red[p] + red[128 + p] + red[256 + p] + red[384 + p] + Wp[args.bout_off];
            if (args.noise && real) raw += args.noise[pl] * args.noise_std;
            sdf_reg = raw * c_out;
            if (real) args.sdf_out[pl] = sdf_reg;
          }
        } else if (epi == EPI_S2_END) {
          red[(0 * 4 + jg) * 128 + p] = gx;
          red[(1 * 4 + jg) * 128 + p] = gy;
          red[(2 * 4 + jg) * 128 + p] = gz;
          named_bar_sync(1, EPI_THREADS);
          if (jg == 0) {
            gx = red[p] + red[128 + p] + red[256 + p] + red[384 + p];
            gy = red[512 + p] + red[640 + p] + red[768 + p] + red[896 + p];
            gz = red[1024 + p] + red[1152 + p] + red[1280 + p] + red[1408 + p];
            float ox = gx, oy = gy, oz = gz;     // g = s R^T g_xs
            if (args.pe.has_transform) {
              ox = args.pe.R[0] * gx + args.pe.R[3] * gy + args.pe.R[6] * gz;
              oy = args.pe.R[1] * gx + args.pe.R[4] * gy + args.pe.R[7] * gz;
              oz = args.pe.R[2] * gx + args.pe.R[5] * gy + args.pe.R[8] * gz;
            }
            const float g[3] = {args.pe.scale * ox, args.pe.scale * oy, args.pe.scale * oz};
            if (real && args.g_out) { args.g_out[pl * 3] = g[0]; args.g_out[pl * 3 + 1] = g[1]; args.g_out[pl * 3 + 2] = g[2]; }
            float sb = 0.f, gb[3] = {0.f, 0.f, 0.f};
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
                  sb = o.sbar; gb[0] = o.gbar[0]; gb[1] = o.gbar[1]; gb[2] = o.gbar[2];
                  tot = o.total;
                  lsum0 += o.l_sdf; lsum1 += o.l_grad; lsum2 += o.l_eik; lsum3 += o.total;
                  sbsum += o.sbar;
                }
                args.loss_mat[pg] = tot;
              }
              // u = s R gbar
              float ux = gb[0], uy = gb[1], uz = gb[2];
              if (args.pe.has_transform) {
                ux = args.pe.R[0] * gb[0] + args.pe.R[1] * gb[1] + args.pe.R[2] * gb[2];
                uy = args.pe.R[3] * gb[0] + args.pe.R[4] * gb[1] + args.pe.R[5] * gb[2];
                uz = args.pe.R[6] * gb[0] + args.pe.R[7] * gb[1] + args.pe.R[8] * gb[2];
              }
              st4(bcast + p * 4, sb, ux * args.pe.scale, uy * args.pe.scale, uz * args.pe.scale);
            }
          }
          if (train) {
            named_bar_sync(1, EPI_THREADS);
            const float4 b4 = ld4(bcast + p * 4);
            sbar = b4.x; u3[0] = b4.y; u3[1] = b4.z; u3[2] = b4.w;
            // abar_e -> A operand of S3 (second pass over this thread's columns)
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int k0 = col_of(i);
              float v[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const int k = k0 + jj;
                float val = 0.f;
                if (k < 3) {
                  val = (k == 0) ? u3[0] : ((k == 1) ? u3[1] : u3[2]);
                } else if (k < E) {
                  const int d = args.feat_d[k];
                  const float ud = (u3[0] * c_ico[d][0] + u3[1] * c_ico[d][1] + u3[2] * c_ico[d][2]) * (float)(1 << args.feat_f[k]);
                  const bool first = (k - 3) < half;
                  const float mate = e32[aux_off_floats(first ? k + half : k - half, p)];
                  val = first ? ud * mate : -ud * mate;
                }
                v[jj] = val;
              }
              put8(v, k0, true, args.arr_ya);
              if (i & 1) chunk_ready(i >> 1);
            }
          }
        }
      }
    }
    // loss sums: warp reduce, one atomic per warp (column group 0 holds the per-point values)
    if (args.mode == TC_MODE_TRAIN && jg == 0) {
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
  if (warp == EPI_WARPS) tmem_dealloc(tmem, 512);
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
