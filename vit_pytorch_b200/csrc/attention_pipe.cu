// Software-pipelined softmax attention for short sequences (N <= 224 keys, dim_head 64): the ViT-B/16 / L/16 shape
// (N = 197 / 196).  Same contract as attention.cu (packed QKV in, merged heads out; vit.py:55-63), different schedule.
//
// Why: with one score tile per softmax warpgroup (attention.cu) every (image, head, query tile) pays the full chain
//   S-MMA round trip -> softmax -> PV round trip -> O read   (5.3 us, of which 2.3 us are arithmetic; profiles/r01j),
// because the next Q K^T cannot be issued before O -- which aliases the score region -- has been read.  Here ONE CTA
// per SM owns all 512 TMEM columns:
//
//     [0, KP)      score region 0   S (fp32) -> P (packed bf16, columns [0, KP/2))
//     [KP, 2KP)    score region 1
//     [2KP, +64)   O accumulator (single slot)
//
// and the MMA thread runs two tiles ahead of the softmax warpgroup.  tcgen05.mma instructions of one CTA execute in
// issue order, so S(j+2) is issued right behind PV(j) into the region PV(j) reads its P from -- no wait in between.
// While the warpgroup is in softmax(j+1), PV(j) and S(j+2) complete; O(j) is read after softmax(j+1), PV(j+1) after
// that read.  No MMA round trip is exposed to the warpgroup any more: its loop is  softmax, O read-out, softmax, ...
//
// Work unit = (image b, head h): K and V are loaded ONCE per unit (attention.cu loaded them per query tile); the unit's
// ceil(N/128) query tiles are consecutive pipeline tiles.  K/V ring of 3 stages + Q ring of 3 tiles: at the target
// rate the kernel moves ~5.7 TB/s, so one whole unit has to be in flight from HBM while another is computed on.
//
// Softmax parallelism: ONE warp per SM sub-partition cannot hide its own instruction latencies (measured: 25 % issue
// utilisation, MUFU 28 % busy, 3.2 us per tile; profiles/r02a).  So a score tile is worked on by 16 warps: warp w owns
// TMEM lanes 32 (w % 4) ... +31 (hardware rule) = 32 query rows, and the four warps of a row quadrant split the KEY
// columns in 16-column chunks (part = w / 4).  Row max and row sum are combined through shared memory with one
// named barrier per quadrant and tile; the O read-out is split the same way (16 of the 64 columns per warp).
//
// Roles: warps 0-15 softmax / epilogue, warp 16 TMA producer, warp 17 MMA issuer + TMEM allocator.
#include "common.cuh"
#include "host_util.h"

namespace b200 {

namespace ap {
constexpr int DH = 64;
constexpr int KV_STAGES = 3;
constexpr int Q_SLOTS = 3;
constexpr int Q_TILE_BYTES = 128 * 128;
constexpr int SM_WARPS = 16;            // softmax warps: 4 row quadrants x 4 column parts
constexpr int THREADS = (SM_WARPS + 2) * 32;
constexpr int MAX_KP = 224;  // 2 * KP + 64 <= 512 TMEM columns
}  // namespace ap

struct AttnPipeParams {
  int B, N, H;
  int KP;           // keys padded to a multiple of 16
  int kv_rows;      // rows of the K / V TMA box (KP rounded up to 8)
  int kv_bytes;     // bytes of one K (or V) slab = kv_rows * 128, multiple of 1024
  int nq;           // query tiles per unit = ceil(N / 128)
  int units;        // B * H
  int I;            // H * dh
  float scale_log2e;
  __nv_bfloat16* out;
  unsigned v_lbo, v_sbo;
};

// max of three (one FMNMX3 on sm_100 instead of two FMNMX)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__global__ void __launch_bounds__(ap::THREADS, 1)
attention_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const AttnPipeParams p) {
  using namespace ap;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_stage_bytes = 2 * p.kv_bytes;
  uint8_t* q_smem = smem + KV_STAGES * kv_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_smem + Q_SLOTS * Q_TILE_BYTES);
  uint64_t* kv_full = bars;                  // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;  // [KV_STAGES]
  uint64_t* q_full = kv_empty + KV_STAGES;   // [Q_SLOTS]
  uint64_t* q_empty = q_full + Q_SLOTS;      // [Q_SLOTS]
  uint64_t* s_full = q_empty + Q_SLOTS;      // [2]
  uint64_t* p_ready = s_full + 2;            // [2]
  uint64_t* o_full = p_ready + 2;            // [1]
  uint64_t* o_free = o_full + 1;             // [1]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_free + 1);
  float* pmax = reinterpret_cast<float*>(tmem_base_smem + 4);  // [2 tile parities][4 parts][128 rows]
  float* psum = pmax + 2 * 4 * 128;                            // [2][4][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TMA_WARP = SM_WARPS, MMA_WARP = SM_WARPS + 1;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < Q_SLOTS; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
    }
    for (int r = 0; r < 2; ++r) {
      mbar_init(&s_full[r], 1);
      mbar_init(&p_ready[r], SM_WARPS);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, SM_WARPS);
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

  // everything above may overlap the tail of the QKV GEMM (programmatic dependent launch)
  pdl_wait();
  pdl_launch_dependents();

  const int n_units = blockIdx.x < p.units ? (p.units - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int n_tiles = n_units * p.nq;
  const uint32_t O_COL = 2 * p.KP;

  if (warp == TMA_WARP) {
    // ---------------------------------------------------------------- producer
    if (lane == 0) {
      int j = 0;
      for (int i = 0; i < n_units; ++i) {
        const int u = blockIdx.x + i * gridDim.x;
        const int h = u % p.H, b = u / p.H;
        const int s = i % KV_STAGES;
        mbar_wait(&kv_empty[s], ((i / KV_STAGES) & 1) ^ 1);
        uint8_t* sk = smem + s * kv_stage_bytes;
        mbar_arrive_expect_tx(&kv_full[s], 2 * p.kv_bytes);
        tma_load_3d(sk, &tmKV, &kv_full[s], p.I + h * DH, 0, b);
        tma_load_3d(sk + p.kv_bytes, &tmKV, &kv_full[s], 2 * p.I + h * DH, 0, b);
        for (int t = 0; t < p.nq; ++t, ++j) {
          const int qs = j % Q_SLOTS;
          mbar_wait(&q_empty[qs], ((j / Q_SLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[qs], Q_TILE_BYTES);
          tma_load_3d(q_smem + qs * Q_TILE_BYTES, &tmQ, &q_full[qs], h * DH, t * 128, b);  // rows >= N: zeros
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, p.KP, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);  // B = V is MN-major
      const int ksteps = p.KP / 16;
      // S(j) = Q_j K^T into score region j & 1
      auto issue_s = [&](int j) {
        const int i = j / p.nq, t = j - i * p.nq;
        const int s = i % KV_STAGES, qs = j % Q_SLOTS;
        if (t == 0) mbar_wait(&kv_full[s], (i / KV_STAGES) & 1);
        mbar_wait(&q_full[qs], (j / Q_SLOTS) & 1);
        tc_fence_after();
        const uint32_t sk = smem_u32(smem + s * kv_stage_bytes);
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(q_smem + qs * Q_TILE_BYTES), 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(sk, 16, 1024);
        const uint32_t d_s = tmem_base + (j & 1) * p.KP;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ss(d_s, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[j & 1]);
        umma_commit(&q_empty[qs]);
      };
      // O(j) = P_j V  (A = P from TMEM, B = V as MN-major smem operand: 16 keys = two 8-row groups = 2048 B)
      auto issue_pv = [&](int j) {
        const int i = j / p.nq, t = j - i * p.nq;
        const int s = i % KV_STAGES;
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
        if (j > 0) mbar_wait(o_free, (j - 1) & 1);
        tc_fence_after();
        const uint32_t sv = smem_u32(smem + s * kv_stage_bytes) + p.kv_bytes;
        const uint32_t a_p = tmem_base + (j & 1) * p.KP;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t vdesc = make_smem_desc_sw128(sv + k * 2048, p.v_lbo, p.v_sbo);
          umma_ts(tmem_base + O_COL, a_p + k * 8, vdesc, idesc_pv, k != 0);
        }
        umma_commit(o_full);
        if (t == p.nq - 1) umma_commit(&kv_empty[s]);  // every MMA reading this K/V stage has been issued
      };
      if (n_tiles > 0) issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int j = 0; j < n_tiles; ++j) {
        issue_pv(j);
        if (j + 2 < n_tiles) issue_s(j + 2);  // right behind PV(j): overwrites the P it has just consumed (in-order)
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax / epilogue warps
    const int quad = warp & 3;   // TMEM lane quadrant = 32 query rows
    const int part = warp >> 2;  // which share of the key columns (and of the 64 output columns)
    const int r_in_tile = quad * 32 + lane;
    const float c = p.scale_log2e;
    const int nch = p.KP >> 4;  // 16-column chunks, dealt to the four parts as evenly as possible
    // (the extra chunks go to the LAST parts: a warp's 4th chunk, which is re-read from TMEM instead of being kept in
    //  registers, then always lies in the upper half of the score columns, which P never overwrites -- see below)
    const int extra_from = 4 - (nch & 3);
    const int ch0 = part * (nch >> 2) + max(part - extra_from, 0);
    const int ch1 = ch0 + (nch >> 2) + (part >= extra_from ? 1 : 0);
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    auto quad_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory"); };

    // O(jj)[:, 16 part .. +16] / rowsum -> bf16 -> out[b, row, h*64 + 16 part ...]
    auto epilogue = [&](int jj, bool active) {
      const int i = jj / p.nq, t = jj - i * p.nq;
      const int u = blockIdx.x + i * gridDim.x;
      const int h = u % p.H, b = u / p.H;
      const int qrow = t * 128 + r_in_tile;
      mbar_wait(o_full, jj & 1);
      tc_fence_after();
      uint32_t r[16];
      if (active) {
        tmem_ld_32x32b_x16(tmem_base + lane_off + O_COL + part * 16, r);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (active && qrow < p.N) {
        const float* ps = psum + (jj & 1) * 512 + r_in_tile;  // the quadrant barrier of the next tile ordered these
        const float inv = 1.0f / ((ps[0] + ps[128]) + (ps[256] + ps[384]));
        uint4* op = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.N + qrow) * p.I + h * DH + part * 16);
#pragma unroll
        for (int q = 0; q < 2; ++q)
          op[q] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * q]) * inv, __uint_as_float(r[8 * q + 1]) * inv),
                             pack_bf16x2(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv),
                             pack_bf16x2(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv),
                             pack_bf16x2(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv));
      }
    };

    bool active_prev = false;
    for (int j = 0; j < n_tiles; ++j) {
      const int t = j % p.nq;
      // quadrants whose 32 query rows all lie beyond N skip the arithmetic but keep every barrier in lockstep
      const bool active = t * 128 + quad * 32 < p.N;
      const uint32_t t_lane = tmem_base + lane_off + (j & 1) * p.KP;
      float* my_max = pmax + (j & 1) * 512 + part * 128 + r_in_tile;
      float* my_sum = psum + (j & 1) * 512 + part * 128 + r_in_tile;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      // The warp's share of the score row is read from TMEM ONCE and kept in registers (3 chunks x 16 fp32).  That is
      // a correctness matter: P is written over S (columns [8 ch, 8 ch + 8) for chunk ch), i.e. into columns other
      // warps of the quadrant own -- the quadrant barrier below, passed only after every warp's loads have completed,
      // makes that safe.  A 4th chunk (KP = 208 / 224 only) is re-read after the barrier instead: it sits at columns
      // >= 144, beyond the [0, KP/2) range P occupies, so nobody overwrites it.
      uint32_t sv[3][16];
      const int my_n = ch1 - ch0;
      const bool tail = my_n == 4;
      const int tail_c0 = (ch0 + 3) * 16;
      // key columns >= N (zero-filled K rows) become -inf right after the load: max ignores them, exp2 gives 0
      auto mask16 = [&](uint32_t (&r)[16], int c0) {
        if (c0 + 16 > p.N) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (c0 + q >= p.N) r[q] = 0xff800000u;
        }
      };
      auto max16 = [&](const uint32_t (&r)[16], float& m0, float& m1) {
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
          m0 = max3(m0, __uint_as_float(r[q]), __uint_as_float(r[q + 1]));
          m1 = max3(m1, __uint_as_float(r[q + 2]), __uint_as_float(r[q + 3]));
        }
      };
      if (active) {
        // ---------------- row max over this warp's columns
        float m0 = -INFINITY, m1 = -INFINITY;
        if (tail) {  // the 4th chunk first, through the registers of chunk 0
          tmem_ld_32x32b_x16(t_lane + tail_c0, sv[0]);
          tmem_ld_wait();
          mask16(sv[0], tail_c0);
          max16(sv[0], m0, m1);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k < my_n) tmem_ld_32x32b_x16(t_lane + (ch0 + k) * 16, sv[k]);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k < my_n) {
            mask16(sv[k], (ch0 + k) * 16);
            max16(sv[k], m0, m1);
          }
        }
        *my_max = fmaxf(m0, m1);
      }
      tc_fence_before();
      quad_barrier();  // partial maxima (and the previous tile's partial sums) visible; every S load has completed
      tc_fence_after();
      float sum = 0.f;
      if (active) {
        const float* pm = pmax + (j & 1) * 512 + r_in_tile;
        const float mc = fmaxf(fmaxf(pm[0], pm[128]), fmaxf(pm[256], pm[384])) * c;
        // ---------------- p = exp2(s*c - max*c), partial row sum, P (bf16 pairs) -> TMEM over S
        const f32x2 c2v = f2_make(c, c), nmc2v = f2_make(-mc, -mc);
        f32x2 acc0 = f2_make(0.f, 0.f), acc1 = f2_make(0.f, 0.f);
        auto exp16 = [&](const uint32_t (&r)[16], int c0) {
          uint32_t pk[8];
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            float x0, x1, x2, x3;
            f2_get(f2_fma(f2_make(__uint_as_float(r[q]), __uint_as_float(r[q + 1])), c2v, nmc2v), x0, x1);
            f2_get(f2_fma(f2_make(__uint_as_float(r[q + 2]), __uint_as_float(r[q + 3])), c2v, nmc2v), x2, x3);
            const float e0 = fast_ex2(x0), e1 = fast_ex2(x1), e2 = fast_ex2(x2), e3 = fast_ex2(x3);
            acc0 = f2_add(acc0, f2_make(e0, e1));
            acc1 = f2_add(acc1, f2_make(e2, e3));
            pk[q >> 1] = pack_bf16x2(e0, e1);
            pk[(q >> 1) + 1] = pack_bf16x2(e2, e3);
          }
          tmem_st_32x32b_x8(t_lane + (c0 >> 1), pk);
        };
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k < my_n) exp16(sv[k], (ch0 + k) * 16);
        if (tail) {  // (re-read: columns >= 144 are never overwritten by P)
          tmem_ld_32x32b_x16(t_lane + tail_c0, sv[0]);
          tmem_ld_wait();
          mask16(sv[0], tail_c0);
          exp16(sv[0], tail_c0);
        }
        float s0, s1, s2, s3;
        f2_get(acc0, s0, s1);
        f2_get(acc1, s2, s3);
        sum = (s0 + s1) + (s2 + s3);
        tmem_st_wait();
      }
      *my_sum = sum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[j & 1]);

      // O of the PREVIOUS tile: its PV ran while this tile's softmax was computed
      if (j > 0) epilogue(j - 1, active_prev);
      active_prev = active;
    }
    if (n_tiles > 0) {
      quad_barrier();  // the last tile's partial sums
      epilogue(n_tiles - 1, active_prev);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// eligible: dh = 64, 2 * KP + 64 <= 512
bool attention_pipe_eligible(int N, int dh) { return dh == ap::DH && (N + 15) / 16 * 16 <= ap::MAX_KP; }

int launch_attention_pipe(const void* qkv, void* out, int B, int N, int H, float scale, unsigned v_lbo, unsigned v_sbo,
                          cudaStream_t stream) {
  using namespace ap;
  AttnPipeParams p{};
  p.B = B; p.N = N; p.H = H;
  p.I = H * DH;
  p.KP = (N + 15) / 16 * 16;
  p.kv_rows = (p.KP + 7) / 8 * 8;
  p.kv_bytes = p.kv_rows * 128;
  p.nq = (N + 127) / 128;
  p.units = B * H;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.v_lbo = v_lbo;
  p.v_sbo = v_sbo;
  B200_CHECK_ARG(p.kv_bytes % 1024 == 0, "attention: K/V slab of %d bytes is not 1024-byte aligned", p.kv_bytes);

  CUtensorMap tmQ, tmKV;
  const uint64_t dims[3] = {(uint64_t)3 * p.I, (uint64_t)N, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)3 * p.I * 2, (uint64_t)N * 3 * p.I * 2};
  {
    const uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap_bf16(&tmQ, qkv, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint32_t box[3] = {64, (uint32_t)p.kv_rows, 1};
    int rc = encode_tmap_bf16(&tmKV, qkv, 3, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem_bytes = (size_t)KV_STAGES * 2 * p.kv_bytes + (size_t)Q_SLOTS * Q_TILE_BYTES +
                            (2 * KV_STAGES + 2 * Q_SLOTS + 6) * 8 + 16 + 2 * 2 * 4 * 128 * sizeof(float) + 1024;
  B200_CHECK_ARG(smem_bytes <= 227 * 1024, "attention: N=%d needs %zu bytes of shared memory", N, smem_bytes);
  B200_ENSURE_SMEM(attention_pipe_kernel, smem_bytes);
  const int grid = p.units < num_sms() ? p.units : num_sms();
  B200_CHECK_CUDA(launch_kernel(attention_pipe_kernel, dim3(grid), dim3(THREADS), smem_bytes, stream, /*pdl=*/true,
                                tmQ, tmKV, p));
  count_launch();
  return 0;
}

}  // namespace b200
