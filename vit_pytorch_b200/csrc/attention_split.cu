// Single-pass attention, "split" variant: TWO warps per 32-row TMEM quadrant, each handling half of the key columns,
// so a 128-row score tile is softmax-ed by 8 warps instead of 4 (4 warps per SM sub-partition with two warpgroups).
// The per-warpgroup chain  S-MMA -> softmax -> PV-MMA -> O read-out  is latency bound (profiles/r01j_attention_timeline
// .txt); halving the softmax and read-out legs shortens it.  Row max and row sum are combined across the two column
// halves through shared memory + a named barrier.  Same data path as attention.cu (TMA 3-D loads, S/P/O in TMEM,
// V as MN-major operand); included from attention.cu's dispatcher.
#include "common.cuh"
#include "host_util.h"

namespace b200 {

struct AttnSplitParams {
  int B, N, H, KP, kv_boxes, kv_box_rows, rounds, units, I;
  int split_col;  // columns [0, split_col) belong to half 0, [split_col, KP) to half 1 (multiple of 32)
  float scale_log2e;
  __nv_bfloat16* out;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int NWG, int STAGES, int TMEM_COLS>
__global__ void __launch_bounds__((8 * NWG + 2) * 32, TMEM_COLS == 256 ? 2 : 1)
attention_split_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const AttnSplitParams p) {
  constexpr int DH = 64;
  constexpr int REGION = TMEM_COLS / NWG;
  constexpr int O_COL = REGION - DH;
  constexpr int NUM_SOFTMAX_WARPS = 8 * NWG;
  constexpr int Q_TILE_BYTES = 128 * 128;
  constexpr int TMA_WARP = NUM_SOFTMAX_WARPS, MMA_WARP = NUM_SOFTMAX_WARPS + 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_bytes = p.kv_boxes * p.kv_box_rows * 128;
  const int stage_bytes = 2 * kv_bytes + NWG * Q_TILE_BYTES;
  float* xmax = reinterpret_cast<float*>(smem + STAGES * stage_bytes);  // [NWG][2][128]
  float* xsum = xmax + NWG * 256;                                       // [NWG][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsum + NWG * 256);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* s_full = empty_bar + STAGES;
  uint64_t* p_ready = s_full + NWG;
  uint64_t* o_full = p_ready + NWG;
  uint64_t* o_free = o_full + NWG;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_free + NWG);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int t = 0; t < NWG; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_ready[t], 8);
      mbar_init(&o_full[t], 1);
      mbar_init(&o_free[t], 8);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(tmem_base_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const int round = u % p.rounds;
        const int bh = u / p.rounds;
        const int h = bh % p.H, b = bh / p.H;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sk = smem + s * stage_bytes;
        uint8_t* sv = sk + kv_bytes;
        uint8_t* sq = sv + kv_bytes;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        for (int i = 0; i < p.kv_boxes; ++i) {
          tma_load_3d(sk + i * p.kv_box_rows * 128, &tmKV, &full_bar[s], p.I + h * DH, i * p.kv_box_rows, b);
          tma_load_3d(sv + i * p.kv_box_rows * 128, &tmKV, &full_bar[s], 2 * p.I + h * DH, i * p.kv_box_rows, b);
        }
        for (int t = 0; t < NWG; ++t)
          tma_load_3d(sq + t * Q_TILE_BYTES, &tmQ, &full_bar[s], h * DH, (round * NWG + t) * 128, b);
      }
    }
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);
      int it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + s * stage_bytes);
        const uint32_t sv = sk + kv_bytes;
        const uint32_t sq = sv + kv_bytes;
        for (int t = 0; t < NWG; ++t) {
          mbar_wait(&o_free[t], (it & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_s = tmem_base + t * REGION;
          for (int n0 = 0; n0 < p.KP; n0 += 256) {
            const int nn = (p.KP - n0) < 256 ? (p.KP - n0) : 256;
            const uint32_t idesc_s = make_idesc_bf16(128, nn, 0, 0);
            const uint64_t adesc = make_smem_desc_sw128(sq + t * Q_TILE_BYTES, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sk + n0 * 128, 16, 1024);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_ss(d_s + n0, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
          }
          umma_commit(&s_full[t]);
        }
        for (int t = 0; t < NWG; ++t) {
          mbar_wait(&p_ready[t], it & 1);
          tc_fence_after();
          const uint32_t d_o = tmem_base + t * REGION + O_COL;
          const int ksteps = p.KP / 16;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t vdesc = make_smem_desc_sw128(sv + k * 2048, 1024, 1024);
            // P of column half 0 is packed at [0, split/2), P of half 1 at [split, split + (KP-split)/2): each half
            // overwrites only score columns its own warp has already consumed
            const int key0 = 16 * k;
            const int a_col = key0 < p.split_col ? (key0 >> 1) : p.split_col + ((key0 - p.split_col) >> 1);
            umma_ts(d_o, tmem_base + t * REGION + a_col, vdesc, idesc_pv, k != 0);
          }
          umma_commit(&o_full[t]);
        }
        umma_commit(&empty_bar[s]);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue: 8 warps per score tile
    const int t = warp >> 3;
    const int quad = warp & 3;
    const int half = (warp >> 2) & 1;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + t * REGION;
    const int r_in_tile = quad * 32 + lane;
    const float c = p.scale_log2e;
    const int c_lo = half ? p.split_col : 0;
    const int c_hi = half ? p.KP : p.split_col;
    float* my_max = xmax + (t * 2 + half) * 128 + r_in_tile;
    float* peer_max = xmax + (t * 2 + (half ^ 1)) * 128 + r_in_tile;
    float* my_sum = xsum + (t * 2 + half) * 128 + r_in_tile;
    float* peer_sum = xsum + (t * 2 + (half ^ 1)) * 128 + r_in_tile;
    int it = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
      const uint32_t up = it & 1;
      const int round = u % p.rounds;
      const int bh = u / p.rounds;
      const int h = bh % p.H, b = bh / p.H;
      const int qrow = (round * NWG + t) * 128 + r_in_tile;
      mbar_wait(&s_full[t], up);
      tc_fence_after();
      // ---- pass 1: max over this warp's columns
      float m0 = -INFINITY, m1 = -INFINITY;
      int c0 = c_lo;
      for (; c0 + 32 <= c_hi && c0 + 32 <= p.N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          m0 = fmaxf(m0, __uint_as_float(r[j]));
          m1 = fmaxf(m1, __uint_as_float(r[j + 1]));
        }
      }
      const int c_tail = c0;
      for (; c0 < c_hi; c0 += 16) {
        uint32_t r16[16];
        tmem_ld_32x32b_x16(t_lane + c0, r16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < p.N) m0 = fmaxf(m0, __uint_as_float(r16[j]));
      }
      *my_max = fmaxf(m0, m1);
      named_bar_sync(1 + t, 256);
      const float mc = fmaxf(fmaxf(m0, m1), *peer_max) * c;
      // ---- pass 2: exponentials, partial row sum, P -> TMEM
      float s0 = 0.f, s1 = 0.f;
      for (c0 = c_lo; c0 < c_tail; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_lane + c0, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = fast_ex2(fmaf(__uint_as_float(r[j]), c, -mc));
          const float e1 = fast_ex2(fmaf(__uint_as_float(r[j + 1]), c, -mc));
          s0 += e0;
          s1 += e1;
          pk[j >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x32b_x16(t_lane + c_lo + ((c0 - c_lo) >> 1), pk);
      }
      for (c0 = c_tail; c0 < c_hi; c0 += 16) {
        uint32_t r16[16];
        tmem_ld_32x32b_x16(t_lane + c0, r16);
        tmem_ld_wait();
        uint32_t pk8[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float e0 = (c0 + j < p.N) ? fast_ex2(fmaf(__uint_as_float(r16[j]), c, -mc)) : 0.f;
          const float e1 = (c0 + j + 1 < p.N) ? fast_ex2(fmaf(__uint_as_float(r16[j + 1]), c, -mc)) : 0.f;
          s0 += e0;
          s1 += e1;
          pk8[j >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x32b_x8(t_lane + c_lo + ((c0 - c_lo) >> 1), pk8);
      }
      tmem_st_wait();
      *my_sum = s0 + s1;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);
      named_bar_sync(1 + t, 256);          // partner's partial sum is visible
      const float inv = 1.0f / ((s0 + s1) + *peer_sum);
      // ---- epilogue: this warp's 32 of the 64 output columns
      mbar_wait(&o_full[t], up);
      tc_fence_after();
      uint32_t r0[32];
      tmem_ld_32x32b_x32(t_lane + O_COL + 32 * half, r0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      if (qrow < p.N) {
        uint4* op = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.N + qrow) * p.I + h * DH + 32 * half);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r0[8 * j]) * inv, __uint_as_float(r0[8 * j + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(r0[8 * j + 2]) * inv, __uint_as_float(r0[8 * j + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(r0[8 * j + 4]) * inv, __uint_as_float(r0[8 * j + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(r0[8 * j + 6]) * inv, __uint_as_float(r0[8 * j + 7]) * inv);
          op[j] = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int NWG, int STAGES, int TMEM_COLS>
static int launch_attention_split(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const AttnSplitParams& p,
                                  size_t smem_bytes, cudaStream_t stream) {
  auto kern = attention_split_kernel<NWG, STAGES, TMEM_COLS>;
  static size_t smem_set = 0;
  if (smem_bytes > smem_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    smem_set = smem_bytes;
  }
  const int slots = num_sms() * (TMEM_COLS == 256 ? 2 : 1);
  const int grid = p.units < slots ? p.units : slots;
  kern<<<grid, (8 * NWG + 2) * 32, smem_bytes, stream>>>(tmQ, tmKV, p);
  B200_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// variant 2: one CTA per SM, two warpgroups of 8 warps, two K/V/Q stages; variant 3: two CTAs per SM, one warpgroup
// of 8 warps each, one stage.  Both need 128 < N and KP <= 256.
int launch_attention_split_variant(int variant, const void* qkv, void* out, int B, int N, int H, float scale,
                                   cudaStream_t stream) {
  AttnSplitParams p{};
  p.B = B; p.N = N; p.H = H;
  p.I = H * 64;
  p.KP = (N + 15) / 16 * 16;
  p.kv_boxes = 1;
  p.kv_box_rows = (p.KP + 7) / 8 * 8;
  p.split_col = (p.KP / 64) * 32;
  if (p.split_col == 0) p.split_col = 32 < p.KP ? 32 : 16;
  const int nwg = variant == 2 ? 2 : 1;
  const int q_tiles = (N + 127) / 128;
  p.rounds = (q_tiles + nwg - 1) / nwg;
  p.units = B * H * p.rounds;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  CUtensorMap tmQ, tmKV;
  const uint64_t dims[3] = {(uint64_t)3 * p.I, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)3 * p.I * 2, (uint64_t)N * 3 * p.I * 2};
  {
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap_bf16(&tmQ, qkv, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint32_t box[3] = {64, (uint32_t)p.kv_box_rows, 1};
    int rc = encode_tmap_bf16(&tmKV, qkv, 3, dims, strides, box);
    if (rc) return rc;
  }
  const int stages = variant == 2 ? 2 : 1;
  const size_t kv_bytes = (size_t)p.kv_box_rows * 128;
  const size_t smem_bytes = stages * (2 * kv_bytes + (size_t)nwg * 128 * 128) + nwg * 2048 + (2 * stages + 4 * nwg) * 8 +
                            16 + 1024;
  if (variant == 2) return launch_attention_split<2, 2, 512>(tmQ, tmKV, p, smem_bytes, stream);
  return launch_attention_split<1, 1, 256>(tmQ, tmKV, p, smem_bytes, stream);
}

}  // namespace b200
