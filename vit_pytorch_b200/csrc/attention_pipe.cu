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
// columns in units of 8 (part = w / 4; 56 / 56 / 48 / 48 columns at N = 197), each share read from TMEM once and kept
// in registers.  Row max and row sum are combined through shared memory (one named barrier per quadrant and tile).
// The O read-out belongs to a warpgroup of its own, so that it overlaps the next tile's softmax (r02d timeline: as
// part of the softmax warps' loop it cost 0.7 of 2.9 us per tile).
//
// Roles: warps 0-15 softmax (4 warpgroups = 4 column parts), 16-19 epilogue (one warp per row quadrant),
//        20 TMA producer, 21 MMA issuer + TMEM allocator, 22-23 idle (they complete the fifth warpgroup).
#include "common.cuh"
#include "host_util.h"

namespace b200 {

namespace ap {
constexpr int DH = 64;
constexpr int KV_STAGES = 3;
constexpr int Q_SLOTS = 3;
constexpr int Q_TILE_BYTES = 128 * 128;
constexpr int SM_WARPS = 16;            // softmax warps: 4 row quadrants x 4 column parts
constexpr int EPI_WARPS = 4;            // O read-out: one warp per row quadrant
constexpr int THREADS = (SM_WARPS + EPI_WARPS + 4) * 32;
constexpr int SUM_SLOTS = 4;            // ring of row-sum hand-over buffers (softmax -> epilogue warps)
constexpr int MAX_KP = 224;  // 2 * KP + 64 <= 512 TMEM columns
}  // namespace ap

struct AttnPipeParams {
  int B, N, H;
  int KP;           // keys padded to a multiple of 16
  int kv_rows;      // rows of the K / V TMA box (KP rounded up to 8)
  int kv_bytes;     // bytes of one K (or V) slab = kv_rows * 128, multiple of 1024
  int kv_stages;    // K/V ring depth: 3, or 2 when three stages do not fit shared memory (KP = 224)
  int nq;           // query tiles per unit = ceil(N / 128)
  int units;        // B * H
  int I;            // H * dh
  float scale_log2e;
  __nv_bfloat16* out;
  unsigned v_lbo, v_sbo;
  long long* trace;  // TRACE instantiation only: [tile][16] %globaltimer stamps of CTA 0 (tools/attn_pipe_trace.py)
};

// max of three (one FMNMX3 on sm_100 instead of two FMNMX)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// EMUL: every second pair of exponentials is evaluated on the FMA pipe (exp2_emul2) instead of MUFU -- the exponential
// phase is MUFU-throughput bound (r02j timeline: 1.3 us per tile against 0.85 us of MUFU issue).
template <bool TRACE, bool EMUL>
__global__ void __launch_bounds__(ap::THREADS, 1)
attention_pipe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                      const AttnPipeParams p) {
  using namespace ap;
  auto stamp = [&](int j, int slot) {
    if (TRACE && blockIdx.x == 0 && j < 64) p.trace[j * 16 + slot] = (long long)globaltimer_ns();
  };
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_stage_bytes = 2 * p.kv_bytes;
  uint8_t* q_smem = smem + p.kv_stages * kv_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_smem + Q_SLOTS * Q_TILE_BYTES);
  uint64_t* kv_full = bars;                  // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES;  // [KV_STAGES]
  uint64_t* q_full = kv_empty + KV_STAGES;   // [Q_SLOTS]
  uint64_t* q_empty = q_full + Q_SLOTS;      // [Q_SLOTS]
  uint64_t* s_full = q_empty + Q_SLOTS;      // [2]
  uint64_t* p_ready = s_full + 2;            // [2]
  uint64_t* o_full = p_ready + 2;            // [1]
  uint64_t* o_free = o_full + 1;             // [1]
  uint64_t* sum_ready = o_free + 1;          // [SUM_SLOTS]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(sum_ready + SUM_SLOTS);
  float* pmax = reinterpret_cast<float*>(tmem_base_smem + 4);  // [2 tile parities][4 parts][128 rows]
  float* psum = pmax + 2 * 4 * 128;                            // [SUM_SLOTS][4][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TMA_WARP = SM_WARPS + EPI_WARPS, MMA_WARP = TMA_WARP + 1;

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
    mbar_init(o_free, EPI_WARPS);
    for (int r = 0; r < SUM_SLOTS; ++r) mbar_init(&sum_ready[r], SM_WARPS);
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
        const int s = i % p.kv_stages;
        mbar_wait(&kv_empty[s], ((i / p.kv_stages) & 1) ^ 1);
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
        const int s = i % p.kv_stages, qs = j % Q_SLOTS;
        if (t == 0) mbar_wait(&kv_full[s], (i / p.kv_stages) & 1);
        mbar_wait(&q_full[qs], (j / Q_SLOTS) & 1);
        tc_fence_after();
        stamp(j, 0);  // S(j): operands present, issue begins
        const uint32_t sk = smem_u32(smem + s * kv_stage_bytes);
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(q_smem + qs * Q_TILE_BYTES), 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(sk, 16, 1024);
        const uint32_t d_s = tmem_base + (j & 1) * p.KP;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ss(d_s, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        umma_commit(&s_full[j & 1]);
        umma_commit(&q_empty[qs]);
        stamp(j, 1);  // S(j) issued
      };
      // O(j) = P_j V  (A = P from TMEM, B = V as MN-major smem operand: 16 keys = two 8-row groups = 2048 B)
      auto issue_pv = [&](int j) {
        const int i = j / p.nq, t = j - i * p.nq;
        const int s = i % p.kv_stages;
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
        stamp(j, 2);  // PV(j): P ready seen
        if (j > 0) mbar_wait(o_free, (j - 1) & 1);
        tc_fence_after();
        stamp(j, 3);  // PV(j): O slot free, issue begins
        const uint32_t sv = smem_u32(smem + s * kv_stage_bytes) + p.kv_bytes;
        const uint32_t a_p = tmem_base + (j & 1) * p.KP;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t vdesc = make_smem_desc_sw128(sv + k * 2048, p.v_lbo, p.v_sbo);
          umma_ts(tmem_base + O_COL, a_p + k * 8, vdesc, idesc_pv, k != 0);
        }
        umma_commit(o_full);
        if (t == p.nq - 1) umma_commit(&kv_empty[s]);  // every MMA reading this K/V stage has been issued
        stamp(j, 4);  // PV(j) issued
      };
      if (n_tiles > 0) issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int j = 0; j < n_tiles; ++j) {
        issue_pv(j);
        if (j + 2 < n_tiles) issue_s(j + 2);  // right behind PV(j): overwrites the P it has just consumed (in-order)
      }
    }
  } else if (warp >= SM_WARPS && warp < SM_WARPS + EPI_WARPS) {
    // ---------------------------------------------------------------- epilogue warps: O(j) / rowsum -> bf16 -> out
    const int quad = warp & 3;
    const int r_in_tile = quad * 32 + lane;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + O_COL;
    for (int j = 0; j < n_tiles; ++j) {
      const int i = j / p.nq, t = j - i * p.nq;
      const int u = blockIdx.x + i * gridDim.x;
      const int h = u % p.H, b = u / p.H;
      const int qrow = t * 128 + r_in_tile;
      const bool active = t * 128 + quad * 32 < p.N;
      mbar_wait(&sum_ready[j % SUM_SLOTS], (j / SUM_SLOTS) & 1);  // the 16 partial row sums of this tile are written
      const float* ps = psum + (j % SUM_SLOTS) * 512 + r_in_tile;
      const float inv = 1.0f / ((ps[0] + ps[128]) + (ps[256] + ps[384]));
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      if (lane == 0 && quad == 0) stamp(j, 12);  // O(j) seen
      uint32_t r0[32], r1[32];
      if (active) {
        tmem_ld_32x32b_x32(t_o, r0);
        tmem_ld_32x32b_x32(t_o + 32, r1);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (active && qrow < p.N) {
        uint4* op = reinterpret_cast<uint4*>(p.out + ((size_t)b * p.N + qrow) * p.I + h * DH);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          op[q] = make_uint4(pack_bf16x2(__uint_as_float(r0[8 * q]) * inv, __uint_as_float(r0[8 * q + 1]) * inv),
                             pack_bf16x2(__uint_as_float(r0[8 * q + 2]) * inv, __uint_as_float(r0[8 * q + 3]) * inv),
                             pack_bf16x2(__uint_as_float(r0[8 * q + 4]) * inv, __uint_as_float(r0[8 * q + 5]) * inv),
                             pack_bf16x2(__uint_as_float(r0[8 * q + 6]) * inv, __uint_as_float(r0[8 * q + 7]) * inv));
#pragma unroll
        for (int q = 0; q < 4; ++q)
          op[4 + q] = make_uint4(pack_bf16x2(__uint_as_float(r1[8 * q]) * inv, __uint_as_float(r1[8 * q + 1]) * inv),
                                 pack_bf16x2(__uint_as_float(r1[8 * q + 2]) * inv, __uint_as_float(r1[8 * q + 3]) * inv),
                                 pack_bf16x2(__uint_as_float(r1[8 * q + 4]) * inv, __uint_as_float(r1[8 * q + 5]) * inv),
                                 pack_bf16x2(__uint_as_float(r1[8 * q + 6]) * inv, __uint_as_float(r1[8 * q + 7]) * inv));
      }
      if (lane == 0 && quad == 0) stamp(j, 13);  // epilogue(j) done
    }
  } else if (warp < SM_WARPS) {
    // ---------------------------------------------------------------- softmax warps
    const int quad = warp & 3;   // TMEM lane quadrant = 32 query rows
    const int part = warp >> 2;  // which share of the key columns
    const int r_in_tile = quad * 32 + lane;
    const float c = p.scale_log2e;
    // key columns in units of 8, dealt to the four parts as evenly as possible: [u0, u1) units = [c_lo, c_hi) columns
    const int nun = p.KP >> 3;
    const int u0 = part * (nun >> 2) + min(part, nun & 3);
    const int u1 = u0 + (nun >> 2) + (part < (nun & 3) ? 1 : 0);
    // The share [8 u0, 8 u1) is read as naturally aligned groups (tcgen05.ld column addresses are kept multiples of
    // the group width): an 8-column group in front if the share starts at an odd unit, then up to three 16-column
    // groups, then an 8-column group at the end if one unit is left.  (With KP a multiple of 16 a share never needs
    // both 8-column groups, so one register array serves either.)
    const bool lead8 = (u0 & 1) != 0 && u1 > u0;
    const int c16 = (u0 + (lead8 ? 1 : 0)) * 8;            // first 16-column group
    const int n16 = (u1 - u0 - (lead8 ? 1 : 0)) >> 1;      // <= 3
    const bool trail8 = ((u1 - u0 - (lead8 ? 1 : 0)) & 1) != 0;
    const bool has8 = lead8 || trail8;
    const int c8 = lead8 ? u0 * 8 : c16 + n16 * 16;        // the 8-column group, if any
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    auto quad_barrier = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + quad) : "memory"); };

    // The warp's share of the score row is read from TMEM ONCE and kept in registers (<= 56 fp32).  That is a
    // correctness matter too: P is written over S (4 packed columns per 8 score columns, at column c/2), i.e. into
    // columns other warps of the quadrant own -- the quadrant barrier, passed only after every warp's loads have
    // completed, makes that safe.
    uint32_t sv[3][16];
    uint32_t s8[8];
    // wait for S(j) and issue (not await) the TMEM loads of this warp's share
    auto fetch_scores = [&](int j) {
      if (warp == 0 && lane == 0) stamp(j, 8);   // softmax(j): waiting for S
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (warp == 0 && lane == 0) stamp(j, 9);   // S(j) seen
      if ((j % p.nq) * 128 + quad * 32 < p.N) {
        const uint32_t tl = tmem_base + lane_off + (j & 1) * p.KP;
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k < n16) tmem_ld_32x32b_x16(tl + c16 + k * 16, sv[k]);
        if (has8) tmem_ld_32x32b_x8(tl + c8, s8);
      }
    };
    for (int j = 0; j < n_tiles; ++j) {
      fetch_scores(j);
      const int t = j % p.nq;
      // quadrants whose 32 query rows all lie beyond N skip the arithmetic but keep every barrier in lockstep
      const bool active = t * 128 + quad * 32 < p.N;
      const uint32_t t_lane = tmem_base + lane_off + (j & 1) * p.KP;
      float* my_max = pmax + (j & 1) * 512 + part * 128 + r_in_tile;
      float* my_sum = psum + (j % SUM_SLOTS) * 512 + part * 128 + r_in_tile;
      // (Issuing the NEXT tile's loads before waiting for this tile's tcgen05.st -- software pipelining across
      //  tiles -- was measured: 235 -> 321 us per layer, the extra live registers spill under the 80-register cap.)
      float sum = 0.f;
      if (active) {
        tmem_ld_wait();
        if (lane == 0 && warp == 0) stamp(j, 5);    // warp 0: S in registers
        if (lane == 0 && warp == 13) stamp(j, 14);  // warp 13 (quadrant 1, part 3): S in registers
        // key columns >= N (zero-filled K rows) become -inf: max ignores them, exp2 gives 0
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k < n16) {
            const int c0 = c16 + k * 16;
            if (c0 + 16 > p.N) {
#pragma unroll
              for (int q = 0; q < 16; ++q)
                if (c0 + q >= p.N) sv[k][q] = 0xff800000u;
            }
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
              m0 = max3(m0, __uint_as_float(sv[k][q]), __uint_as_float(sv[k][q + 1]));
              m1 = max3(m1, __uint_as_float(sv[k][q + 2]), __uint_as_float(sv[k][q + 3]));
            }
          }
        }
        if (has8) {
          const int c0 = c8;
          if (c0 + 8 > p.N) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + q >= p.N) s8[q] = 0xff800000u;
          }
#pragma unroll
          for (int q = 0; q < 8; q += 4) {
            m0 = max3(m0, __uint_as_float(s8[q]), __uint_as_float(s8[q + 1]));
            m1 = max3(m1, __uint_as_float(s8[q + 2]), __uint_as_float(s8[q + 3]));
          }
        }
        *my_max = fmaxf(m0, m1);
      }
      tc_fence_before();
      quad_barrier();  // partial maxima visible; every S load of the quadrant has completed
      tc_fence_after();
      if (warp == 0 && lane == 0) stamp(j, 10);  // past the quadrant barrier
      if (warp == 13 && lane == 0) stamp(j, 15);
      if (active) {
        const float* pm = pmax + (j & 1) * 512 + r_in_tile;
        const float mc = fmaxf(fmaxf(pm[0], pm[128]), fmaxf(pm[256], pm[384])) * c;
        // ---------------- p = exp2(s*c - max*c), partial row sum, P (bf16 pairs) -> TMEM over S
        const f32x2 c2v = f2_make(c, c), nmc2v = f2_make(-mc, -mc);
        f32x2 acc0 = f2_make(0.f, 0.f), acc1 = f2_make(0.f, 0.f);
#define AP_EXP4(R, Q, PK, I)                                                                               \
  {                                                                                                        \
    float x0, x1, x2, x3;                                                                                  \
    f2_get(f2_fma(f2_make(__uint_as_float(R[(Q)]), __uint_as_float(R[(Q) + 1])), c2v, nmc2v), x0, x1);     \
    f2_get(f2_fma(f2_make(__uint_as_float(R[(Q) + 2]), __uint_as_float(R[(Q) + 3])), c2v, nmc2v), x2, x3); \
    const float e0 = fast_ex2(x0), e1 = fast_ex2(x1);                                                      \
    float e2, e3;                                                                                          \
    if (EMUL) {                                                                                            \
      exp2_emul2(f2_make(x2, x3), e2, e3);                                                                 \
    } else {                                                                                               \
      e2 = fast_ex2(x2);                                                                                   \
      e3 = fast_ex2(x3);                                                                                   \
    }                                                                                                      \
    acc0 = f2_add(acc0, f2_make(e0, e1));                                                                  \
    acc1 = f2_add(acc1, f2_make(e2, e3));                                                                  \
    PK[(I)] = pack_bf16x2(e0, e1);                                                                         \
    PK[(I) + 1] = pack_bf16x2(e2, e3);                                                                     \
  }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k < n16) {
            uint32_t pk[8];
#pragma unroll
            for (int q = 0; q < 16; q += 4) AP_EXP4(sv[k], q, pk, q >> 1)
            tmem_st_32x32b_x8(t_lane + ((c16 + k * 16) >> 1), pk);
          }
        }
        if (has8) {
          uint32_t pk4[4];
          AP_EXP4(s8, 0, pk4, 0)
          AP_EXP4(s8, 4, pk4, 2)
          tmem_st_32x32b_x4(t_lane + (c8 >> 1), pk4);
        }
#undef AP_EXP4
        float s0, s1, s2, s3;
        f2_get(acc0, s0, s1);
        f2_get(acc1, s2, s3);
        sum = (s0 + s1) + (s2 + s3);
        if (lane == 0 && warp == 0) stamp(j, 6);    // warp 0: exponentials computed, P stores issued
      }
      *my_sum = sum;
      if (active) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&p_ready[j & 1]);
        mbar_arrive(&sum_ready[j % SUM_SLOTS]);
      }
      if (warp == 0 && lane == 0) stamp(j, 11);  // P(j) handed over
      if (warp == 13 && lane == 0) stamp(j, 7);
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

static std::atomic<long long*> g_pipe_trace{nullptr};
void attention_pipe_set_trace(long long* buf) { g_pipe_trace = buf; }
static std::atomic<int> g_pipe_emul{0};  // test hook 13: 0 = all exponentials on MUFU (default), 1 = half on the FMA pipe
void attention_pipe_set_emul(int v) { g_pipe_emul = v; }

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
  p.trace = g_pipe_trace.load();
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
  auto smem_for = [&](int stages) {
    return (size_t)stages * 2 * p.kv_bytes + (size_t)Q_SLOTS * Q_TILE_BYTES +
           (2 * KV_STAGES + 2 * Q_SLOTS + 6 + SUM_SLOTS) * 8 + 16 + (2 + SUM_SLOTS) * 4 * 128 * sizeof(float) + 1024;
  };
  p.kv_stages = smem_for(KV_STAGES) <= 227 * 1024 ? KV_STAGES : KV_STAGES - 1;
  const size_t smem_bytes = smem_for(p.kv_stages);
  B200_CHECK_ARG(smem_bytes <= 227 * 1024, "attention: N=%d needs %zu bytes of shared memory", N, smem_bytes);
  const int grid = p.units < num_sms() ? p.units : num_sms();
  const bool emul = g_pipe_emul.load() != 0;
  if (p.trace) {  // timing experiment (test hook): separately compiled instantiation
    B200_ENSURE_SMEM((attention_pipe_kernel<true, true>), smem_bytes);
    B200_CHECK_CUDA(launch_kernel(attention_pipe_kernel<true, true>, dim3(grid), dim3(THREADS), smem_bytes, stream,
                                  false, tmQ, tmKV, p));
  } else if (emul) {
    B200_ENSURE_SMEM((attention_pipe_kernel<false, true>), smem_bytes);
    B200_CHECK_CUDA(launch_kernel(attention_pipe_kernel<false, true>, dim3(grid), dim3(THREADS), smem_bytes, stream,
                                  /*pdl=*/true, tmQ, tmKV, p));
  } else {
    B200_ENSURE_SMEM((attention_pipe_kernel<false, false>), smem_bytes);
    B200_CHECK_CUDA(launch_kernel(attention_pipe_kernel<false, false>, dim3(grid), dim3(THREADS), smem_bytes, stream,
                                  /*pdl=*/true, tmQ, tmKV, p));
  }
  count_launch();
  return 0;
}

}  // namespace b200
