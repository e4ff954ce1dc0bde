// Multi-head softmax attention for short sequences (N <= 512 keys: one pass over the keys, no online rescale),
// reading q/k/v tiles straight out of the packed [B*N, 3*H*dh] QKV buffer with TMA (no head-split copies) and
// writing the merged-heads [B*N, H*dh] layout directly.  Replaces vit.py:55-63 / simple_vit.py:54-61:
//     q,k,v = split heads;  dots = q k^T * scale;  attn = softmax(dots);  out = attn v;  merge heads
//
// Per work unit (image b, head h, round of NWG query tiles of 128 rows):
//   TMA warp   : K[KP,64], V[KP,64], Q[128,64] x NWG  -> 128B-swizzled smem (3-D tensor map => rows >= N are zero)
//   MMA thread : S_t = Q_t K^T           tcgen05.mma 128 x KP x 64      -> TMEM region t (fp32, KP columns)
//   softmax WG : thread == query row (tcgen05.ld 32x32b): row max, p = exp2((s-max)*scale*log2e), row sum,
//                P as bf16 back into TMEM (aliasing S, FA4 style)  [or into swizzled smem: PSMEM variant]
//   MMA thread : O_t = P_t V             tcgen05.mma 128 x 64 x KP, A from TMEM, B = V as MN-major smem operand
//   softmax WG : O / rowsum -> bf16 -> out[b, row, h*64 : h*64+64]
//
// TMEM map of one region (512/NWG columns): S at [0,KP); P (packed bf16 pairs) at [0,KP/2); O at [REGION-64, REGION).
#include "common.cuh"
#include "host_util.h"

namespace b200 {

constexpr int ATT_DH = 64;

struct AttnParams {
  int B, N, H;
  int KP;         // keys padded to a multiple of 16
  int kv_boxes;   // number of TMA boxes per K (and per V)
  int kv_box_rows;
  int rounds;     // query-tile rounds per (b, h)
  int units;      // B * H * rounds
  int I;          // H * dh
  float scale_log2e;
  __nv_bfloat16* out;
  unsigned v_lbo, v_sbo;  // V (MN-major) descriptor strides, bytes
};

__host__ __device__ inline int att_kv_bytes(int kv_boxes, int kv_box_rows) { return kv_boxes * kv_box_rows * 128; }

template <int NWG, int STAGES, bool PSMEM>
__global__ void __launch_bounds__((4 * NWG + 2) * 32, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                 const AttnParams p) {
  constexpr int REGION = 512 / NWG;
  constexpr int O_COL = REGION - ATT_DH;
  constexpr int NUM_SOFTMAX_WARPS = 4 * NWG;
  constexpr int Q_TILE_BYTES = 128 * 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_bytes = att_kv_bytes(p.kv_boxes, p.kv_box_rows);  // multiple of 1024
  const int stage_bytes = 2 * kv_bytes + NWG * Q_TILE_BYTES;
  const int p_chunks = (p.KP + 63) / 64;
  const int p_tile_bytes = PSMEM ? p_chunks * 128 * 128 : 0;
  uint8_t* p_smem = smem + STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + NWG * p_tile_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* s_full = empty_bar + STAGES;
  uint64_t* p_ready = s_full + NWG;
  uint64_t* o_full = p_ready + NWG;
  uint64_t* o_free = o_full + NWG;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_free + NWG);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TMA_WARP = NUM_SOFTMAX_WARPS;
  constexpr int MMA_WARP = NUM_SOFTMAX_WARPS + 1;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int t = 0; t < NWG; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_ready[t], 4);
      mbar_init(&o_full[t], 1);
      mbar_init(&o_free[t], 4);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) {
    tmem_alloc(tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == TMA_WARP) {
    // ---------------------------------------------------------------- producer
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
          tma_load_3d(sk + i * p.kv_box_rows * 128, &tmKV, &full_bar[s], p.I + h * ATT_DH, i * p.kv_box_rows, b);
          tma_load_3d(sv + i * p.kv_box_rows * 128, &tmKV, &full_bar[s], 2 * p.I + h * ATT_DH, i * p.kv_box_rows, b);
        }
        for (int t = 0; t < NWG; ++t) {
          const int qt = round * NWG + t;  // may be >= q_tiles: box fully out of bounds -> zeros
          tma_load_3d(sq + t * Q_TILE_BYTES, &tmQ, &full_bar[s], h * ATT_DH, qt * 128, b);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_pv = make_idesc_bf16(128, ATT_DH, 0, 1);  // B = V is MN-major
      int it = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const uint32_t up = it & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + s * stage_bytes);
        const uint32_t sv = sk + kv_bytes;
        const uint32_t sq = sv + kv_bytes;
        // S_t = Q_t K^T
        for (int t = 0; t < NWG; ++t) {
          mbar_wait(&o_free[t], up ^ 1);  // region t drained by the previous unit's epilogue
          tc_fence_after();
          const uint32_t d_s = tmem_base + t * REGION;
          for (int n0 = 0; n0 < p.KP; n0 += 256) {
            const int nn = (p.KP - n0) < 256 ? (p.KP - n0) : 256;
            const uint32_t idesc_s = make_idesc_bf16(128, nn, 0, 0);
            const uint64_t adesc = make_smem_desc_sw128(sq + t * Q_TILE_BYTES, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sk + n0 * 128, 16, 1024);
#pragma unroll
            for (int k = 0; k < ATT_DH / 16; ++k) umma_ss(d_s + n0, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
          }
          umma_commit(&s_full[t]);
        }
        // O_t = P_t V
        for (int t = 0; t < NWG; ++t) {
          mbar_wait(&p_ready[t], up);
          tc_fence_after();
          const uint32_t d_o = tmem_base + t * REGION + O_COL;
          const int ksteps = p.KP / 16;
          for (int k = 0; k < ksteps; ++k) {
            // 16 keys = two 8-row groups of V = 2048 B
            const uint64_t vdesc = make_smem_desc_sw128(sv + k * 2048, p.v_lbo, p.v_sbo);
            if (PSMEM) {
              // P tile in smem: K-major, 64-key chunks of [128 rows x 128 B]
              const uint32_t pa = smem_u32(p_smem + t * p_tile_bytes) + (k >> 2) * (128 * 128) + (k & 3) * 32;
              umma_ss(d_o, make_smem_desc_sw128(pa, 16, 1024), vdesc, idesc_pv, k != 0);
            } else {
              umma_ts(d_o, tmem_base + t * REGION + k * 8, vdesc, idesc_pv, k != 0);
            }
          }
          umma_commit(&o_full[t]);
        }
        umma_commit(&empty_bar[s]);  // K/V/Q of this stage no longer needed once everything above completed
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue warpgroups
    const int t = warp >> 2;
    const int quad = warp & 3;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + t * REGION;
    const int r_in_tile = quad * 32 + lane;
    const float c = p.scale_log2e;
    int it = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++it) {
      const uint32_t up = it & 1;
      const int round = u % p.rounds;
      const int bh = u / p.rounds;
      const int h = bh % p.H, b = bh / p.H;
      const int qrow = (round * NWG + t) * 128 + r_in_tile;

      // warps whose 32 query rows all lie beyond N skip the arithmetic but keep the barrier protocol in lockstep
      const bool warp_active = (round * NWG + t) * 128 + quad * 32 < p.N;

      mbar_wait(&s_full[t], up);
      tc_fence_after();
      float sum = 1.f;
      if (warp_active) {
        const int n_chunks = (p.KP + 31) >> 5;  // 32-column chunks; the last one may be 16 wide
        // chunk loader: x32, or x16 for a 16-wide tail (upper half then holds stale values that the col < N mask drops)
        auto load_chunk = [&](uint32_t (&r)[32], int ci) {
          const int c0 = ci << 5;
          if (c0 + 32 <= p.KP) {
            tmem_ld_32x32b_x32(t_lane + c0, r);
          } else {
            uint32_t lo[16];
            tmem_ld_32x32b_x16(t_lane + c0, lo);
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = lo[j];
          }
        };
        uint32_t ra[32], rb[32];
        // ---- pass 1: row max over the valid keys (software pipelined: next chunk's tcgen05.ld in flight)
        float mx = -INFINITY;
        auto max_chunk = [&](const uint32_t (&r)[32], int ci) {
          const int c0 = ci << 5;
          if (c0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < p.N) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        };
        load_chunk(ra, 0);
        for (int ci = 0; ci < n_chunks; ci += 2) {
          tmem_ld_wait();
          if (ci + 1 < n_chunks) load_chunk(rb, ci + 1);
          max_chunk(ra, ci);
          if (ci + 1 < n_chunks) {
            tmem_ld_wait();
            if (ci + 2 < n_chunks) load_chunk(ra, ci + 2);
            max_chunk(rb, ci + 1);
          }
        }
        const float mc = mx * c;
        // ---- pass 2: p = exp2(s*c - max*c), row sum, P (bf16 pairs) -> TMEM over S, or swizzled smem
        sum = 0.f;
        auto exp_chunk = [&](const uint32_t (&r)[32], int ci) {
          const int c0 = ci << 5;
          float pv[32];
          if (c0 + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pv[j] = fast_ex2(fmaf(__uint_as_float(r[j]), c, -mc));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              pv[j] = (c0 + j < p.N) ? fast_ex2(fmaf(__uint_as_float(r[j]), c, -mc)) : 0.f;
          }
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum += pv[2 * j] + pv[2 * j + 1];
            pk[j] = pack_bf16x2(pv[2 * j], pv[2 * j + 1]);
          }
          const bool full = c0 + 32 <= p.KP;
          if (PSMEM) {
            uint8_t* prow = p_smem + t * p_tile_bytes + (c0 >> 6) * (128 * 128) + (r_in_tile >> 3) * 1024 +
                            (r_in_tile & 7) * 128;
            const int jj = (c0 & 63) >> 3;
            const int sw7 = r_in_tile & 7;
            *reinterpret_cast<uint4*>(prow + (((jj) ^ sw7) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(prow + (((jj + 1) ^ sw7) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if (full) {
              *reinterpret_cast<uint4*>(prow + (((jj + 2) ^ sw7) << 4)) = make_uint4(pk[8], pk[9], pk[10], pk[11]);
              *reinterpret_cast<uint4*>(prow + (((jj + 3) ^ sw7) << 4)) = make_uint4(pk[12], pk[13], pk[14], pk[15]);
            }
          } else if (full) {
            tmem_st_32x32b_x16(t_lane + (c0 >> 1), pk);
          } else {
            uint32_t pk8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pk8[j] = pk[j];
            tmem_st_32x32b_x8(t_lane + (c0 >> 1), pk8);
          }
        };
        load_chunk(ra, 0);
        for (int ci = 0; ci < n_chunks; ci += 2) {
          tmem_ld_wait();
          if (ci + 1 < n_chunks) load_chunk(rb, ci + 1);
          exp_chunk(ra, ci);
          if (ci + 1 < n_chunks) {
            tmem_ld_wait();
            if (ci + 2 < n_chunks) load_chunk(ra, ci + 2);
            exp_chunk(rb, ci + 1);
          }
        }
        if (PSMEM) {
          fence_proxy_async_smem();
        } else {
          tmem_st_wait();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);

      // epilogue: O / sum -> bf16 -> global
      const float inv = 1.0f / sum;
      mbar_wait(&o_full[t], up);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      if (warp_active) {
        tmem_ld_32x32b_x32(t_lane + O_COL, r0);
        tmem_ld_32x32b_x32(t_lane + O_COL + 32, r1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      if (warp_active && qrow < p.N) {
        __nv_bfloat16* op = p.out + ((size_t)b * p.N + qrow) * p.I + h * ATT_DH;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(r0[j]) * inv, __uint_as_float(r0[j + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(r0[j + 2]) * inv, __uint_as_float(r0[j + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(r0[j + 4]) * inv, __uint_as_float(r0[j + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(r0[j + 6]) * inv, __uint_as_float(r0[j + 7]) * inv);
          *reinterpret_cast<uint4*>(op + j) = pk;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(r1[j]) * inv, __uint_as_float(r1[j + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(r1[j + 2]) * inv, __uint_as_float(r1[j + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(r1[j + 4]) * inv, __uint_as_float(r1[j + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(r1[j + 6]) * inv, __uint_as_float(r1[j + 7]) * inv);
          *reinterpret_cast<uint4*>(op + 32 + j) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// debug / experiment knobs (b200vit_debug_set)
static int g_attn_psmem = 0;     // 1: stage P through shared memory instead of TMEM
static int g_attn_v_lbo = 1024;  // V descriptor leading-dim byte offset
static int g_attn_v_sbo = 1024;  // V descriptor stride-dim byte offset

template <int NWG, int STAGES, bool PSMEM>
static int launch_attention(const CUtensorMap& tmQ, const CUtensorMap& tmKV, const AttnParams& p, size_t smem_bytes,
                            cudaStream_t stream) {
  auto kern = attention_kernel<NWG, STAGES, PSMEM>;
  static size_t smem_set = 0;
  if (smem_bytes > smem_set) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    smem_set = smem_bytes;
  }
  const int grid = p.units < num_sms() ? p.units : num_sms();
  kern<<<grid, (4 * NWG + 2) * 32, smem_bytes, stream>>>(tmQ, tmKV, p);
  B200_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200vit_debug_set(int key, int value) {
  switch (key) {
    case 1: g_attn_psmem = value; return 0;
    case 2: g_attn_v_lbo = value; return 0;
    case 3: g_attn_v_sbo = value; return 0;
    case 4: gemm_force_version(value); return 0;
    default: return B200VIT_ERR_INVALID;
  }
}

extern "C" int b200vit_attention(const void* qkv, void* out, int B, int N, int H, int dh, float scale, void* stream) {
  B200_CHECK_ARG(qkv && out, "attention: null pointer");
  B200_CHECK_ARG(B > 0 && N > 0 && H > 0, "attention: bad shape B=%d N=%d H=%d", B, N, H);
  B200_CHECK_ARG(dh == ATT_DH, "attention: dim_head=%d not supported by this build (only 64)", dh);
  B200_CHECK_ARG(N <= 512, "attention: N=%d > 512 needs the (unbuilt) online-softmax path", N);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "attention: pointers must be 16-byte aligned");
  AttnParams p{};
  p.B = B; p.N = N; p.H = H;
  p.I = H * dh;
  p.KP = (N + 15) / 16 * 16;
  const int nwg = (N > 128 && p.KP <= 256) ? 2 : 1;
  p.kv_boxes = (p.KP + 255) / 256;
  p.kv_box_rows = ((p.KP + p.kv_boxes - 1) / p.kv_boxes + 7) / 8 * 8;
  const int q_tiles = (N + 127) / 128;
  p.rounds = (q_tiles + nwg - 1) / nwg;
  p.units = B * H * p.rounds;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.v_lbo = (unsigned)g_attn_v_lbo;
  p.v_sbo = (unsigned)g_attn_v_sbo;

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
  const bool psmem = g_attn_psmem != 0;
  const size_t kv_bytes = (size_t)att_kv_bytes(p.kv_boxes, p.kv_box_rows);
  const size_t stage_bytes = 2 * kv_bytes + (size_t)nwg * 128 * 128;
  const size_t p_bytes = psmem ? (size_t)nwg * ((p.KP + 63) / 64) * 128 * 128 : 0;
  auto smem_for = [&](int st) { return st * stage_bytes + p_bytes + (2 * st + 4 * nwg) * 8 + 16 + 1024; };
  // two K/V/Q stages when they fit (prefetch of the next unit), else one
  const int stages = (!psmem && smem_for(2) <= 227 * 1024) ? 2 : 1;
  const size_t smem_bytes = smem_for(stages);
  B200_CHECK_ARG(smem_bytes <= 227 * 1024, "attention: N=%d needs %zu bytes of shared memory", N, smem_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (nwg == 2) {
    if (psmem) return launch_attention<2, 1, true>(tmQ, tmKV, p, smem_bytes, st);
    if (stages == 1) return launch_attention<2, 1, false>(tmQ, tmKV, p, smem_bytes, st);
    return launch_attention<2, 2, false>(tmQ, tmKV, p, smem_bytes, st);
  }
  if (psmem) return launch_attention<1, 1, true>(tmQ, tmKV, p, smem_bytes, st);
  if (stages == 1) return launch_attention<1, 1, false>(tmQ, tmKV, p, smem_bytes, st);
  return launch_attention<1, 2, false>(tmQ, tmKV, p, smem_bytes, st);
}
