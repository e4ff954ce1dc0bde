// GEMM v2: CTA-pair (cta_group::2) tcgen05 GEMM with a shared-memory staged, TMA-stored epilogue.
//
//   out[M,N] = epilogue(A[M,K] * W[N,K]^T)          cluster of 2 CTAs = one 256 x 256 output tile
//
//   * each CTA of the pair owns 128 rows of A and HALF of the W tile (128 of the 256 N-rows); the pair's tensor cores
//     run one 256 x 256 x 16 MMA (issued by the leader CTA) that reads both halves, so every SM stages only
//     32 KB per 64-wide k-block (16 KB A + 16 KB W) instead of 48 KB  ->  4-5 stages + 64-96 KB of epilogue staging.
//   * accumulators: 2 x 256 fp32 columns of TMEM in each CTA (its own 128 rows), double buffered against the epilogue.
//   * epilogue: thread == accumulator row (tcgen05.ld 32x32b), results written 16 B at a time into 128B-swizzled
//     staging boxes (bank-conflict free) and shipped with TMA stores (coalesced, clipped at the matrix edges).
//       MODE_BF16 (QKV, FC1): 16 epilogue warps (4 per SM sub-partition: the GELU epilogue is latency/issue bound),
//                             one 64-column bf16 box per warp per tile.
//       MODE_F32  (patch, out-proj, FC2): 8 warps, 32-column fp32 boxes; the fp32 residual box is TMA-LOADED into the
//                             staging buffer, updated in place in shared memory and stored back.
//       MODE_DUAL (LN-fold producer): MODE_F32 + a bf16 copy of the new rows + their (sum, sum^2).
//
// Barrier protocol (every barrier exists at the same smem offset in both CTAs):
//   full[s]       leader only : 1 arrive (leader producer, expect_tx = bytes of BOTH CTAs) + TMA complete_tx from both
//   empty[s]      both        : tcgen05.commit multicast from the leader's MMA thread
//   tmem_full[a]  both        : tcgen05.commit multicast
//   tmem_empty[a] leader only : 2 x EPI_WARPS arrives (remote arrive from the peer CTA's epilogue warps)
#include "common.cuh"
#include "host_util.h"

namespace b200 {

namespace g2 {
constexpr int MODE_BF16 = 0, MODE_F32 = 1, MODE_DUAL = 2;
constexpr int BLOCK_M = 128;       // rows per CTA (256 per pair)
constexpr int BLOCK_N = 256;       // columns per tile (each CTA stages 128 W rows)
constexpr int BLOCK_K = 64;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB
constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;  // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int EPI_BUF_BYTES = 4096;                   // one 32-row x 128-byte swizzled box
// EW = epilogue warps per CTA.  MODE_BF16: 16.  fp32 modes: 8, or 4 for long-K problems (FC2: the 48 k-blocks of a
// tile leave the epilogue ample time, and halving its staging boxes buys a fifth operand stage).
template <int MODE, int EW>
struct Cfg {
  static constexpr int EPI_WARPS = EW;
  static constexpr int STAGES = MODE == MODE_DUAL ? (EW == 4 ? 5 : 4) : (MODE == MODE_F32 && EW == 4 ? 6 : 5);
  static constexpr int BUFS_PER_WARP = MODE == MODE_BF16 ? 1 : (MODE == MODE_DUAL ? 3 : 2);
  static constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int COLS_PER_WARP = BLOCK_N / (EPI_WARPS / 4);
  static constexpr int EPI_BYTES = EPI_WARPS * BUFS_PER_WARP * EPI_BUF_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES + EPI_BYTES;
  // full[S] empty[S] tmem_full[2] tmem_empty[2] resid_full[EPI_WARPS][2] out_ready[EPI_WARPS][2] bbuf_free[EPI_WARPS]
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 5 * EPI_WARPS;
  static constexpr int DYN_BYTES = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
};
}  // namespace g2

struct Gemm2Params {
  int M, N, K;
  int num_m_pairs, num_n_tiles, num_k_blocks;
  int flags;
  const float* bias;
  const float* ln_sums;  // [M][ln_parts][2]
  int ln_parts;
  int stats_parts;
  float ln_inv_dim, ln_eps;
  const float* col_s;
  float* stats_out;  // MODE_DUAL: [M][stats_parts][2] partial (sum, sum of squares) of the bf16-rounded output rows
  const float* head_gamma;  // EPI_HEADNORM: fp32 [norm_cols] scale of the normalised leading heads
  int norm_cols;            //               columns [0, norm_cols) are normalised per 64-wide head
  float head_eps;           // EPI_HEADLN:   epsilon of the per-head LayerNorm (instead of the RMS norm)
};

// LN-fold, bias and GELU on W consecutive accumulator columns starting at col0 (W = 16 or 32), on fp32 PAIRS (FFMA2):
//   LN-fold + bias:  y = acc * rstd + (bias - rstd*mu * s)        2 packed FMAs per pair
template <int W>
__device__ __forceinline__ void epilogue_math(float (&v)[W], int col0, int flags, float mu, float rstd,
                                              const Gemm2Params& p) {
  if (flags & (B200VIT_EPI_LNFOLD | B200VIT_EPI_BIAS)) {
    const bool fold = (flags & B200VIT_EPI_LNFOLD) != 0;
    const bool has_bias = (flags & B200VIT_EPI_BIAS) != 0;
    const float k = -rstd * mu;
    const f32x2 k2 = f2_make(k, k), r2 = f2_make(fold ? rstd : 1.0f, fold ? rstd : 1.0f);
#pragma unroll
    for (int j = 0; j < W; j += 4) {
      // columns at or beyond N are clipped by the TMA store: read any valid address for them instead of branching
      const int cj = min(col0 + j, p.N - 4);
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (fold) s4 = __ldg(reinterpret_cast<const float4*>(p.col_s + cj));
      if (has_bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cj));
      f32x2 c01 = f2_make(b4.x, b4.y), c23 = f2_make(b4.z, b4.w);
      if (fold) {
        c01 = f2_fma(k2, f2_make(s4.x, s4.y), c01);
        c23 = f2_fma(k2, f2_make(s4.z, s4.w), c23);
      }
      f2_get(f2_fma(f2_make(v[j], v[j + 1]), r2, c01), v[j], v[j + 1]);
      f2_get(f2_fma(f2_make(v[j + 2], v[j + 3]), r2, c23), v[j + 2], v[j + 3]);
    }
  }
  if (flags & B200VIT_EPI_GELU) {
#pragma unroll
    for (int j = 0; j < W; j += 2) gelu_erf2(v[j], v[j + 1]);
  }
}

// per-row LayerNorm statistics from up to 8 partial (sum, sum^2) pairs: all loads are issued before the first add
// (independent L2 round trips instead of a dependent chain), summation order is fixed => deterministic
__device__ __forceinline__ void load_row_stats(const Gemm2Params& p, int row, float& mu, float& rstd) {
  float2 part[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    part[i] = (i < p.ln_parts) ? __ldg(reinterpret_cast<const float2*>(p.ln_sums + 2 * ((size_t)row * p.ln_parts + i)))
                               : make_float2(0.f, 0.f);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1 += part[i].x;
    s2 += part[i].y;
  }
  for (int i = 8; i < p.ln_parts; ++i) {  // (more than 8 partials: N of the producer > 1024)
    const float2 ss = __ldg(reinterpret_cast<const float2*>(p.ln_sums + 2 * ((size_t)row * p.ln_parts + i)));
    s1 += ss.x;
    s2 += ss.y;
  }
  mu = s1 * p.ln_inv_dim;
  rstd = rsqrtf(fmaxf(s2 * p.ln_inv_dim - mu * mu, 0.f) + p.ln_eps);
}
// pull the bias / column-sum slices a warp is about to use into L1 (non-blocking)
__device__ __forceinline__ void prefetch_cols(const Gemm2Params& p, int flags, int col_base, int ncols, int lane) {
  const int c = col_base + lane * 32;  // one 128-byte line per lane
  if (lane * 32 < ncols && c < p.N) {
    if (flags & B200VIT_EPI_BIAS) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bias + c));
    if (flags & B200VIT_EPI_LNFOLD) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.col_s + c));
  }
}

// CTF: the epilogue flags as a compile-time constant (the inner loops then carry no flag tests), or -1 = read p.flags
template <int MODE, int CTF, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::Cfg<MODE, EW>::NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmResid,
             const __grid_constant__ CUtensorMap tmOutB, const Gemm2Params p) {
  using namespace g2;
  using C = Cfg<MODE, EW>;
  constexpr int STAGES = C::STAGES;
  constexpr int EPI_WARPS = C::EPI_WARPS;
  constexpr int COLS_PER_WARP = C::COLS_PER_WARP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* resid_full = tmem_empty + 2;             // [EPI_WARPS][2] staging box b holds its input / is free
  uint64_t* out_ready = resid_full + 2 * EPI_WARPS;  // [EPI_WARPS][2] staging box b holds finished output
  uint64_t* bbuf_free = out_ready + 2 * EPI_WARPS;   // [EPI_WARPS]    bf16 copy box has been stored (MODE_DUAL)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bbuf_free + EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    if (p.flags & B200VIT_EPI_RESIDUAL) tma_prefetch_desc(&tmResid);
    if (MODE == MODE_DUAL) tma_prefetch_desc(&tmOutB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 2 * EPI_WARPS);
    }
    for (int i = 0; i < 2 * EPI_WARPS; ++i) {
      mbar_init(&resid_full[i], 1);
      mbar_init(&out_ready[i], 1);
    }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&bbuf_free[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cg2(tmem_base_smem, 512);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();     // CTA-local hand-off of the TMEM base address written by tcgen05.alloc
  cluster_sync_all();  // peer barriers initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap
  // the tail of the previous kernel in the stream; from here on its results are read and ours are written.
  pdl_wait();
  pdl_launch_dependents();

  const int num_tiles = p.num_m_pairs * p.num_n_tiles;
  constexpr int nst = STAGES;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_pair = tile / p.num_n_tiles;
        const int n_blk = tile % p.num_n_tiles;
        const int m0 = m_pair * (2 * BLOCK_M) + rank * BLOCK_M;
        const int n0 = n_blk * BLOCK_N + rank * (BLOCK_N / 2);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
          tma_load_2d_cg2(sa, &tmA, full_leader, kb * BLOCK_K, m0);
          tma_load_2d_cg2(sb, &tmB, full_leader, kb * BLOCK_K, n0);
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sa + A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_ss_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_mc(&empty_bar[stage], 0x3);
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_mc(&tmem_full[acc], 0x3);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ epilogue TMA lanes (MODE_F32 / MODE_DUAL)
    // Lane e moves the staging boxes of epilogue warp e: it stores finished boxes, waits until the store engine has
    // read them and refills the buffer with the residual box two steps ahead (or just marks it free).  The epilogue
    // warps themselves never block on a TMA store.
    if constexpr (MODE != MODE_BF16) {
      if (lane < EPI_WARPS) {
        const int e = lane;
        const int quad = e & 3;
        const int col_off = (e >> 2) * COLS_PER_WARP;
        constexpr int BOXES = COLS_PER_WARP / 32;
        const bool has_resid = (p.flags & B200VIT_EPI_RESIDUAL) != 0;
        uint8_t* buf0 = epi_smem + e * C::BUFS_PER_WARP * EPI_BUF_BYTES;
        uint8_t* bbuf = buf0 + 2 * EPI_BUF_BYTES;
        const int my_tiles = cluster_id < num_tiles ? (num_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
        const uint32_t total = static_cast<uint32_t>(my_tiles) * BOXES;
        auto coords = [&](uint32_t g, int& cc, int& cr) {
          const int tile = cluster_id + static_cast<int>(g / BOXES) * num_clusters;
          const int m_pair = tile / p.num_n_tiles;
          const int n_blk = tile % p.num_n_tiles;
          cr = m_pair * (2 * BLOCK_M) + rank * BLOCK_M + quad * 32;
          cc = n_blk * BLOCK_N + col_off + static_cast<int>(g % BOXES) * 32;
        };
        auto refill = [&](uint32_t g) {  // make buffer g & 1 ready for box g
          uint64_t* bar = &resid_full[2 * e + (g & 1)];
          if (has_resid) {
            int cc, cr;
            coords(g, cc, cr);
            mbar_arrive_expect_tx(bar, EPI_BUF_BYTES);
            tma_load_2d(buf0 + (g & 1) * EPI_BUF_BYTES, &tmResid, bar, cc, cr);
          } else {
            mbar_arrive(bar);
          }
        };
        if (total > 0) refill(0);
        if (total > 1) refill(1);
        for (uint32_t g = 0; g < total; ++g) {
          mbar_wait(&out_ready[2 * e + (g & 1)], (g >> 1) & 1);
          int cc, cr;
          coords(g, cc, cr);
          tma_store_2d(&tmOut, buf0 + (g & 1) * EPI_BUF_BYTES, cc, cr);
          if (MODE == MODE_DUAL && (g & 1)) tma_store_2d(&tmOutB, bbuf, cc - 32, cr);
          tma_store_commit();
          tma_store_wait_read<0>();
          if (MODE == MODE_DUAL && (g & 1)) mbar_arrive(&bbuf_free[e]);
          if (g + 2 < total) refill(g + 2);
        }
        tma_store_wait<0>();  // all global writes complete before exit
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int e = warp - 4;
    const int quad = warp & 3;
    const int col_off = (e >> 2) * COLS_PER_WARP;
    const int flags = CTF >= 0 ? CTF : p.flags;
    const uint32_t tmem_empty_leader0 = mapa_shared(smem_u32(&tmem_empty[0]), 0);
    const uint32_t tmem_empty_leader1 = mapa_shared(smem_u32(&tmem_empty[1]), 0);
    const uint32_t sw = static_cast<uint32_t>(lane & 7);
    uint8_t* buf0 = epi_smem + e * C::BUFS_PER_WARP * EPI_BUF_BYTES;
    uint8_t* my_row0 = buf0 + lane * 128;  // this thread's 128-byte row inside a staging box
    const uint32_t my_row0_s = smem_u32(my_row0);
    int acc = 0;
    uint32_t acc_phase = 0;

    auto release_tmem = [&]() {
      // all TMEM reads of this tile are done: hand the accumulator back to the leader's MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc ? tmem_empty_leader1 : tmem_empty_leader0);
    };

    if constexpr (MODE == MODE_BF16) {
      // ============ bf16 output: 64 columns per warp = one staging box per tile, x16 loads double buffered
      int stats_m_pair = -1;
      float stats_mu = 0.f, stats_rstd = 1.f;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_pair = tile / p.num_n_tiles;
        const int n_blk = tile % p.num_n_tiles;
        const int row0 = m_pair * (2 * BLOCK_M) + rank * BLOCK_M + quad * 32;
        const int row = row0 + lane;
        const int col_base = n_blk * BLOCK_N + col_off;
        prefetch_cols(p, flags, col_base, COLS_PER_WARP, lane);
        float mu = 0.f, rstd = 1.f;
        if (flags & B200VIT_EPI_LNFOLD) {
          if (row < p.M) {
            if (m_pair != stats_m_pair) {
              load_row_stats(p, row, stats_mu, stats_rstd);
              stats_m_pair = m_pair;
            }
            mu = stats_mu;
            rstd = stats_rstd;
          }
          // the statistics of this thread's row in the NEXT tile of this cluster: into L1 while this tile is processed
          const int ntile = tile + num_clusters;
          if (ntile < num_tiles) {
            const int nrow = (ntile / p.num_n_tiles) * (2 * BLOCK_M) + rank * BLOCK_M + quad * 32 + lane;
            if (nrow < p.M)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(p.ln_sums + 2 * ((size_t)nrow * p.ln_parts)));
          }
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + col_off;
        uint32_t ra[16], rb[16];
        tmem_ld_32x32b_x16(t_row, ra);
        if (lane == 0) tma_store_wait_read<0>();  // last tile's store has finished reading the staging box
        __syncwarp();

        float ss = 0.f;  // EPI_HEADNORM: sum of squares of this row's 64 (bf16-rounded) values = one head
        float hs = 0.f;  // EPI_HEADLN: their sum
        auto emit16 = [&](const uint32_t (&r)[16], int cidx) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_math<16>(v, col_base + cidx * 16, flags, mu, rstd, p);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t w0 = pack_bf16x2(v[8 * q], v[8 * q + 1]), w1 = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
            const uint32_t w2 = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), w3 = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
            sts_v4(my_row0_s + ((static_cast<uint32_t>(cidx * 2 + q) ^ sw) << 4), w0, w1, w2, w3);
            if (flags & B200VIT_EPI_HEADNORM) {
              const uint32_t w4[4] = {w0, w1, w2, w3};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float lo = __uint_as_float(w4[i] << 16), hi = __uint_as_float(w4[i] & 0xFFFF0000u);
                ss = fmaf(lo, lo, fmaf(hi, hi, ss));
                if (flags & B200VIT_EPI_HEADLN) hs += lo + hi;
              }
            }
          }
        };
        tmem_ld_wait();
        tmem_ld_32x32b_x16(t_row + 16, rb);
        emit16(ra, 0);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(t_row + 32, ra);
        emit16(rb, 1);
        tmem_ld_wait();
        tmem_ld_32x32b_x16(t_row + 48, rb);
        emit16(ra, 2);
        tmem_ld_wait();
        release_tmem();
        emit16(rb, 3);
        if ((flags & B200VIT_EPI_HEADNORM) && col_base < p.norm_cols) {
          // this warp's 64 columns are exactly one head: every thread rescales its own row inside the staging box
          // RMS norm: v * sqrt(dh) / max(||v||, eps) * gamma;  LayerNorm (no bias): (v - mean) * rsqrt(var + eps) * gamma
          const bool hln = (flags & B200VIT_EPI_HEADLN) != 0;
          const float mean = hln ? hs * (1.0f / 64.0f) : 0.f;
          const float inv = hln ? rsqrtf(fmaxf(ss * (1.0f / 64.0f) - mean * mean, 0.f) + p.head_eps)
                                : 8.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t sp = my_row0_s + ((static_cast<uint32_t>(q) ^ sw) << 4);
            const float4 raw = lds_v4f(sp);
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.head_gamma + col_base + 8 * q));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.head_gamma + col_base + 8 * q + 4));
            const uint32_t w[4] = {__float_as_uint(raw.x), __float_as_uint(raw.y), __float_as_uint(raw.z),
                                   __float_as_uint(raw.w)};
            const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float lo = __uint_as_float(w[i] << 16), hi = __uint_as_float(w[i] & 0xFFFF0000u);
              o[i] = pack_bf16x2((lo - mean) * inv * gg[2 * i], (hi - mean) * inv * gg[2 * i + 1]);
            }
            sts_v4(sp, o[0], o[1], o[2], o[3]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, buf0, col_base, row0);
          tma_store_commit();
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    } else {
      // ============ fp32 output (optionally residual in place, optionally + bf16 copy + statistics)
      constexpr bool DUAL = MODE == MODE_DUAL;
      const bool has_resid = (flags & B200VIT_EPI_RESIDUAL) != 0;
      uint8_t* bbuf = buf0 + 2 * EPI_BUF_BYTES;  // DUAL only: bf16 copy staging box (64 columns)
      uint64_t* in_bar = resid_full + 2 * e;
      uint64_t* out_bar = out_ready + 2 * e;
      uint32_t box_seq = 0;  // running count of fp32 boxes used by this warp (buffer = box_seq & 1)
      float st_sum = 0.f, st_sq = 0.f;

      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m_pair = tile / p.num_n_tiles;
        const int n_blk = tile % p.num_n_tiles;
        const int row = m_pair * (2 * BLOCK_M) + rank * BLOCK_M + quad * 32 + lane;
        const bool row_ok = row < p.M;
        float mu = 0.f, rstd = 1.f;
        prefetch_cols(p, flags, n_blk * BLOCK_N + col_off, COLS_PER_WARP, lane);
        if ((flags & B200VIT_EPI_LNFOLD) && row_ok) load_row_stats(p, row, mu, rstd);
        if (DUAL) st_sum = st_sq = 0.f;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + col_off;

#pragma unroll 1
        for (int c = 0; c < COLS_PER_WARP; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + c, r);
          tmem_ld_wait();
          if (c + 32 == COLS_PER_WARP) release_tmem();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_math<32>(v, n_blk * BLOCK_N + col_off + c, flags, mu, rstd, p);

          const uint32_t myrow_s = my_row0_s + (box_seq & 1) * EPI_BUF_BYTES;
          mbar_wait(&in_bar[box_seq & 1], (box_seq >> 1) & 1);  // residual box landed / buffer free
          if (has_resid) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t sp = myrow_s + ((static_cast<uint32_t>(q) ^ sw) << 4);
              float4 x = lds_v4f(sp);
              x.x += v[4 * q]; x.y += v[4 * q + 1]; x.z += v[4 * q + 2]; x.w += v[4 * q + 3];
              sts_v4f(sp, x.x, x.y, x.z, x.w);
              if (DUAL) {
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
              }
            }
            if (DUAL) {
              // bf16 copy of the new residual rows (A operand of the next, LN-folded GEMM) + its row statistics
              const int half = (c >> 5) & 1;
              const uint32_t nmask = (n_blk * BLOCK_N + col_off + c < p.N) ? 0xFFFFFFFFu : 0u;
              if (half == 0) mbar_wait(&bbuf_free[e], ((box_seq >> 1) & 1) ^ 1);  // previous pair's store has read it
              const uint32_t brow_s = smem_u32(bbuf) + lane * 128;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 pk;
                pk.x = pack_bf16x2(v[8 * q], v[8 * q + 1]);
                pk.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
                pk.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
                pk.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
                sts_v4(brow_s + ((static_cast<uint32_t>(half * 4 + q) ^ sw) << 4), pk.x, pk.y, pk.z, pk.w);
                // (columns at or beyond N -- a last tile of 64 / 128 / 192 valid columns -- are clipped by the TMA
                //  stores; they must not reach the statistics either: masked to +0, which leaves the sums bit-exact)
                const uint32_t w4[4] = {pk.x & nmask, pk.y & nmask, pk.z & nmask, pk.w & nmask};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float lo = __uint_as_float(w4[i] << 16);
                  const float hi = __uint_as_float(w4[i] & 0xFFFF0000u);
                  st_sum += lo + hi;
                  st_sq = fmaf(lo, lo, fmaf(hi, hi, st_sq));
                }
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              sts_v4f(myrow_s + ((static_cast<uint32_t>(q) ^ sw) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2],
                      v[4 * q + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&out_bar[box_seq & 1]);  // the TMA lane stores it and recycles the buffer
          ++box_seq;
          if (DUAL && EPI_WARPS == 4 && c == 96) {  // end of the tile's first 128 columns: their statistics slot
            if (row_ok)
              *reinterpret_cast<float2*>(p.stats_out + 2 * ((size_t)row * p.stats_parts + n_blk * 2)) =
                  make_float2(st_sum, st_sq);
            st_sum = st_sq = 0.f;
          }
        }
        if (DUAL && row_ok) {
          // one statistics slot per 128 output columns (the same association of the partial sums whichever kernel
          // variant produced them: results do not depend on the batch size).  With 8 warps each column half of the
          // tile has its own warp; with 4 warps the first half was written inside the loop above.
          const int part = n_blk * 2 + (EPI_WARPS == 4 ? 1 : (e >> 2));
          if (part < p.stats_parts)
            *reinterpret_cast<float2*>(p.stats_out + 2 * ((size_t)row * p.stats_parts + part)) =
                make_float2(st_sum, st_sq);
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    if (MODE == MODE_BF16 && lane == 0) tma_store_wait<0>();  // all global writes of this warp complete before exit
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 512);
  }
}

static std::atomic<int> g_gemm_force{0};  // test hook: 0 auto, 1 force v1, 2 force v2 (wherever its epilogue applies)
static std::atomic<int> g_gemm_ew{0};     // test hook: 0 auto, 4 / 8 force the epilogue-warp count of the fp32 modes

int gemm2_eligible(int M, int N, int K, int64_t ldo, int flags, const void* out_bf16, const float* out_f32,
                   const float* resid) {
  (void)K;
  (void)resid;
  if (g_gemm_force.load() == 1) return 0;
  const bool dual = out_bf16 && out_f32;
  if (dual) {
    // fp32 stream (in place, residual) + bf16 copy + statistics: the LN-fold producer epilogue
    if (!(flags & B200VIT_EPI_RESIDUAL) || !(flags & B200VIT_EPI_STATS) || (flags & B200VIT_EPI_GELU)) return 0;
    if ((ldo % 8) != 0 || (N % 64) != 0) return 0;
  } else {
    if (flags & B200VIT_EPI_STATS) return 0;
    if (out_bf16 && (flags & B200VIT_EPI_RESIDUAL)) return 0;
    if (out_f32 && (flags & B200VIT_EPI_GELU)) return 0;
  }
  if (out_bf16 && (ldo % 8) != 0) return 0;  // TMA: 16-byte row pitch
  if (out_f32 && (ldo % 4) != 0) return 0;
  if ((N % 4) != 0) return 0;                // float4 bias / col_s loads
  if (g_gemm_force.load() == 2) return 1;
  return (M >= 1024 && N >= 256) ? 1 : 0;    // small problems: the single-CTA kernel has finer tiles
}

template <int MODE, int CTF, int EW>
static int launch_gemm2_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                          const CUtensorMap& tmResid, const CUtensorMap& tmOutB, const Gemm2Params& p, int clusters,
                          cudaStream_t stream) {
  using C = g2::Cfg<MODE, EW>;
  static_assert(C::DYN_BYTES <= 227 * 1024, "gemm2: shared memory budget");
  auto kern = gemm2_kernel<MODE, CTF, EW>;
  B200_ENSURE_SMEM(kern, C::DYN_BYTES);
  B200_CHECK_CUDA(launch_kernel(kern, dim3(2 * clusters), dim3(C::NUM_THREADS), C::DYN_BYTES, stream, /*pdl=*/true,
                                tmA, tmB, tmOut, tmResid, tmOutB, p));
  count_launch();
  return 0;
}

int launch_gemm2(const void* A, int64_t lda, const void* W, int64_t ldw, void* out_bf16, float* out_f32, int64_t ldo,
                 const float* bias, const float* resid, const float* ln_sums, int ln_parts, float ln_eps,
                 const float* col_s, float* stats_out, int M, int N, int K, int flags, cudaStream_t stream,
                 const float* head_gamma, int norm_cols, float head_eps) {
  using namespace g2;
  const bool dual = out_bf16 && out_f32;
  Gemm2Params p{};
  p.stats_out = stats_out;
  p.M = M; p.N = N; p.K = K;
  p.num_m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  p.num_n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.flags = flags;
  p.bias = bias;
  p.ln_sums = ln_sums;
  p.ln_parts = ln_parts;
  p.stats_parts = b200vit_stats_parts(N);
  p.ln_inv_dim = 1.0f / (float)K;
  p.ln_eps = ln_eps;
  p.col_s = col_s;
  p.head_gamma = head_gamma;
  p.norm_cols = norm_cols;
  p.head_eps = head_eps;

  CUtensorMap tmA, tmB, tmOut, tmResid, tmOutB;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)BLOCK_M};
    int rc = encode_tmap_bf16(&tmA, A, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)ldw * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)(BLOCK_N / 2)};
    int rc = encode_tmap_bf16(&tmB, W, 2, dims, strides, box);
    if (rc) return rc;
  }
  const uint64_t odims[2] = {(uint64_t)N, (uint64_t)M};
  if (out_f32) {
    const uint64_t strides[1] = {(uint64_t)ldo * 4};
    const uint32_t box[2] = {32, 32};
    int rc = encode_tmap_f32(&tmOut, out_f32, 2, odims, strides, box, true);
    if (rc) return rc;
    rc = encode_tmap_f32(&tmResid, resid ? resid : out_f32, 2, odims, strides, box, true);
    if (rc) return rc;
    tmOutB = tmOut;
    if (dual) {
      const uint64_t bstrides[1] = {(uint64_t)ldo * 2};
      const uint32_t bbox[2] = {64, 32};
      rc = encode_tmap_bf16(&tmOutB, out_bf16, 2, odims, bstrides, bbox);
      if (rc) return rc;
    }
  } else {
    const uint64_t strides[1] = {(uint64_t)ldo * 2};
    const uint32_t box[2] = {64, 32};
    int rc = encode_tmap_bf16(&tmOut, out_bf16, 2, odims, strides, box);
    if (rc) return rc;
    tmResid = tmOut;
    tmOutB = tmOut;
  }
  const int tiles = p.num_m_pairs * p.num_n_tiles;
  int clusters = num_sms() / 2;
  if (tiles < clusters) clusters = tiles;
#define B200_G2_LAUNCH_EW(MODE, CTF, EW) \
  launch_gemm2_t<MODE, CTF, EW>(tmA, tmB, tmOut, tmResid, tmOutB, p, clusters, stream)
#define B200_G2_LAUNCH(MODE, CTF) B200_G2_LAUNCH_EW(MODE, CTF, (MODE == MODE_BF16 ? 16 : 8))
  // long-K fp32 epilogues (FC2): 4 epilogue warps, one more operand stage
  const int ew_force = g_gemm_ew.load();
  const bool ew4 = ew_force == 4 || (ew_force == 0 && K >= 2048);
  constexpr int F_BIAS = B200VIT_EPI_BIAS, F_GELU = B200VIT_EPI_GELU, F_RES = B200VIT_EPI_RESIDUAL,
                F_FOLD = B200VIT_EPI_LNFOLD, F_STATS = B200VIT_EPI_STATS, F_HN = B200VIT_EPI_HEADNORM,
                F_HLN = B200VIT_EPI_HEADLN;
  // the flag combinations of a transformer block get their own instantiation, anything else the generic kernel
  if (dual) {
    if (flags == (F_BIAS | F_RES | F_STATS))
      return ew4 ? B200_G2_LAUNCH_EW(MODE_DUAL, F_BIAS | F_RES | F_STATS, 4)
                 : B200_G2_LAUNCH(MODE_DUAL, F_BIAS | F_RES | F_STATS);
    if (flags == (F_RES | F_STATS))
      return ew4 ? B200_G2_LAUNCH_EW(MODE_DUAL, F_RES | F_STATS, 4) : B200_G2_LAUNCH(MODE_DUAL, F_RES | F_STATS);
    return B200_G2_LAUNCH(MODE_DUAL, -1);
  }
  if (out_f32) return ew4 ? B200_G2_LAUNCH_EW(MODE_F32, -1, 4) : B200_G2_LAUNCH(MODE_F32, -1);
  if (flags & F_HN) {  // (only reached through b200vit_gemm_headnorm_bf16)
    if (flags == (F_HN | F_FOLD | F_BIAS)) return B200_G2_LAUNCH(MODE_BF16, F_HN | F_FOLD | F_BIAS);
    if (flags == (F_HN | F_HLN | F_FOLD | F_BIAS)) return B200_G2_LAUNCH(MODE_BF16, F_HN | F_HLN | F_FOLD | F_BIAS);
    if (flags == F_HN) return B200_G2_LAUNCH(MODE_BF16, F_HN);
    return B200_G2_LAUNCH(MODE_BF16, -1);
  }
  if (flags == (F_FOLD | F_BIAS)) return B200_G2_LAUNCH(MODE_BF16, F_FOLD | F_BIAS);
  if (flags == (F_FOLD | F_BIAS | F_GELU)) return B200_G2_LAUNCH(MODE_BF16, F_FOLD | F_BIAS | F_GELU);
  if (flags == (F_BIAS | F_GELU)) return B200_G2_LAUNCH(MODE_BF16, F_BIAS | F_GELU);
  return B200_G2_LAUNCH(MODE_BF16, -1);
#undef B200_G2_LAUNCH
#undef B200_G2_LAUNCH_EW
}

void gemm_force_version(int v) { g_gemm_force = v; }
void gemm2_force_epilogue_warps(int v) { g_gemm_ew = v; }

}  // namespace b200
