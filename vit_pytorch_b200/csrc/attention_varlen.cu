// Variable-length multi-head softmax attention over PACKED token sequences (cu_seqlens), any sequence length.
//
//   qkv[T, 3*H*64] bf16 (columns [q | k | v], head-major), sequences s = tokens [cu[s], cu[s+1])
//   out[T, H*64]   bf16 :  out_s = softmax(q_s k_s^T * scale) v_s        -- tokens of different sequences never mix
//
// This is the block-diagonal ("same image id") attention of NaViT (reference na_vit.py:335-337,161-166) in its
// mask-free varlen form, and the long-sequence path of ViT (N > 512).  THREE kernels live in this file; the default is
// attention_varlen2_kernel<.., ONLINE = true> further down (64-key blocks, three score buffers, ONE pass over the keys
// with a lazily moved reference max).  First, the original serial kernel (test hook 11 = 1).  Work unit = (sequence,
// head, 128-row query tile); keys are walked in blocks of 128:
//   phase 1 (only if the sequence has more than one key block): S_b = Q K_b^T for every block, softmax warps keep the
//           running row max -- the FINAL max is known before any exponential is taken, so
//   phase 2 needs no online rescaling: S_b again, P_b = exp2((S_b - max) * scale*log2e) as bf16 back into TMEM
//           (aliasing S_b), O += P_b V_b accumulates in TMEM across blocks (tcgen05.mma accumulate flag), row sums in
//           registers.  QK^T is computed twice; attention is exp/bandwidth bound, the tensor pipe has the headroom.
// Keys >= n are masked to P = 0; rows >= n are computed on whatever follows in the buffer and never stored.
// One softmax warpgroup per CTA, 192 TMEM columns (S|P at [0,128), O at [128,192)), 80 KB smem -> 2 CTAs per SM.
#include "common.cuh"
#include "host_util.h"

namespace b200 {

namespace av {
constexpr int DH = 64;
constexpr int KB = 128;                      // keys per block
constexpr int TILE_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr int KV_STAGES = 2;
constexpr int SMEM_DATA = TILE_BYTES * (1 + 2 * KV_STAGES);   // Q + 2 x (K, V)
constexpr int NUM_BARS = 2 + 2 * KV_STAGES + 5;               // q_full q_empty kv_full[] kv_empty[] s_full s_free p_ready pv_done o_free
constexpr int DYN_BYTES = SMEM_DATA + NUM_BARS * 8 + 16 + 1024;
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 128;
constexpr int THREADS = 6 * 32;
}  // namespace av

struct AttnVarlenParams {
  const int* cu_seqlens;   // [S+1] device
  const int* tile_prefix;  // [S+1] device: number of 128-row query tiles before sequence s
  int num_seqs, H, I;
  int units;               // total_tiles * H
  float scale_log2e;
  __nv_bfloat16* out;
  long long* trace;        // timing experiment (pipelined kernel, TRACE instantiation): [unit][16] stamps of CTA 0
};

__device__ __forceinline__ void av_locate(const AttnVarlenParams& p, int unit, int& seq, int& h, int& qt, int& row0,
                                          int& n) {
  const int tile = unit / p.H;  // heads of one query tile are neighbours: K/V of the sequence stay in L2
  h = unit % p.H;
  int lo = 0, hi = p.num_seqs;  // largest s with tile_prefix[s] <= tile
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (p.tile_prefix[mid] <= tile) lo = mid; else hi = mid;
  }
  seq = lo;
  qt = tile - p.tile_prefix[seq];
  row0 = p.cu_seqlens[seq];
  n = p.cu_seqlens[seq + 1] - row0;
}

__global__ void __launch_bounds__(av::THREADS, 2)
attention_varlen_kernel(const __grid_constant__ CUtensorMap tm, const AttnVarlenParams p) {
  using namespace av;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + TILE_BYTES;  // stage st: K at sKV + st*2*TILE_BYTES, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_DATA);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + KV_STAGES;
  uint64_t* s_full = kv_empty + KV_STAGES;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_ready = s_free + 1;
  uint64_t* pv_done = p_ready + 1;
  uint64_t* o_free = pv_done + 1;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TMA_WARP = 4, MMA_WARP = 5;

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_ready, 4);
    mbar_init(pv_done, 1);
    mbar_init(o_free, 4);
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
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t units_done = 0, kv_count = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++units_done) {
        int seq, h, qt, row0, n;
        av_locate(p, u, seq, h, qt, row0, n);
        const int nb = (n + KB - 1) / KB;
        mbar_wait(q_empty, (units_done & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, TILE_BYTES);
        tma_load_2d(sQ, &tm, q_full, h * DH, row0 + qt * 128);
        const int steps = (nb > 1 ? nb : 0) + nb;
        for (int s = 0; s < steps; ++s, ++kv_count) {
          const bool phase2 = s >= steps - nb;
          const int kb = phase2 ? s - (steps - nb) : s;
          const int st = kv_count % KV_STAGES;
          mbar_wait(&kv_empty[st], ((kv_count / KV_STAGES) & 1) ^ 1);
          uint8_t* k_dst = sKV + st * 2 * TILE_BYTES;
          mbar_arrive_expect_tx(&kv_full[st], phase2 ? 2 * TILE_BYTES : TILE_BYTES);
          tma_load_2d(k_dst, &tm, &kv_full[st], p.I + h * DH, row0 + kb * KB);
          if (phase2) tma_load_2d(k_dst + TILE_BYTES, &tm, &kv_full[st], 2 * p.I + h * DH, row0 + kb * KB);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KB, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);
      uint32_t units_done = 0, kv_count = 0, n_sfree = 0, n_pready = 0, n_pv = 0;
      bool wait_sfree = false, wait_pv = false;
      const uint32_t sq = smem_u32(sQ);
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++units_done) {
        int seq, h, qt, row0, n;
        av_locate(p, u, seq, h, qt, row0, n);
        const int nb = (n + KB - 1) / KB;
        const int steps = (nb > 1 ? nb : 0) + nb;
        mbar_wait(q_full, units_done & 1);
        for (int s = 0; s < steps; ++s, ++kv_count) {
          const bool phase2 = s >= steps - nb;
          const int kb = phase2 ? s - (steps - nb) : s;
          const int st = kv_count % KV_STAGES;
          mbar_wait(&kv_full[st], (kv_count / KV_STAGES) & 1);
          // the S|P region is free once the softmax warps have read S (phase 1) / the previous P V has completed
          if (wait_sfree) {
            mbar_wait(s_free, (n_sfree - 1) & 1);
            wait_sfree = false;
          }
          if (wait_pv) {
            mbar_wait(pv_done, (n_pv - 1) & 1);
            wait_pv = false;
          }
          tc_fence_after();
          const uint32_t sk = smem_u32(sKV + st * 2 * TILE_BYTES);
          {
            const uint64_t adesc = make_smem_desc_sw128(sq, 16, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(sk, 16, 1024);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
          }
          umma_commit(s_full);
          if (s == steps - 1) umma_commit(q_empty);  // last use of the Q tile
          if (!phase2) {
            umma_commit(&kv_empty[st]);
            ++n_sfree;
            wait_sfree = true;
          } else {
            if (kb == 0) {
              mbar_wait(o_free, (units_done & 1) ^ 1);  // previous unit's O has been read
            }
            mbar_wait(p_ready, n_pready & 1);
            ++n_pready;
            tc_fence_after();
            const uint32_t sv = sk + TILE_BYTES;
#pragma unroll
            for (int k = 0; k < KB / 16; ++k) {
              const uint64_t vdesc = make_smem_desc_sw128(sv + k * 2048, 1024, 1024);
              umma_ts(tmem_base + O_COL, tmem_base + k * 8, vdesc, idesc_pv, (kb | k) != 0);
            }
            umma_commit(&kv_empty[st]);
            umma_commit(pv_done);
            ++n_pv;
            wait_pv = true;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroup (thread == query row)
    const int quad = warp & 3;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float c = p.scale_log2e;
    uint32_t units_done = 0, n_sfull = 0, n_pv = 0;
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++units_done) {
      int seq, h, qt, row0, n;
      av_locate(p, u, seq, h, qt, row0, n);
      const int nb = (n + KB - 1) / KB;
      const int qrow = qt * 128 + quad * 32 + lane;
      float mx = -INFINITY;
      auto block_max = [&](int kb) {
#pragma unroll 1
        for (int c0 = 0; c0 < KB; c0 += 32) {
          if (kb * KB + c0 >= n) break;
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_lane + c0, r);
          tmem_ld_wait();
          const int lim = n - (kb * KB + c0);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < lim) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      };
      if (nb > 1) {
        for (int kb = 0; kb < nb; ++kb, ++n_sfull) {
          mbar_wait(s_full, n_sfull & 1);
          tc_fence_after();
          block_max(kb);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
      }
      float sum = 0.f;
      for (int kb = 0; kb < nb; ++kb, ++n_sfull) {
        mbar_wait(s_full, n_sfull & 1);
        tc_fence_after();
        if (nb == 1) block_max(0);
        const float mc = mx * c;
#pragma unroll 1
        for (int c0 = 0; c0 < KB; c0 += 32) {
          uint32_t r[32];
          uint32_t pk[16];
          const int lim = n - (kb * KB + c0);  // number of valid keys in this chunk (may be <= 0)
          if (lim > 0) {
            tmem_ld_32x32b_x32(t_lane + c0, r);
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float e0 = (j < lim) ? fast_ex2(fmaf(__uint_as_float(r[j]), c, -mc)) : 0.f;
            const float e1 = (j + 1 < lim) ? fast_ex2(fmaf(__uint_as_float(r[j + 1]), c, -mc)) : 0.f;
            sum += e0 + e1;
            pk[j >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st_32x32b_x16(t_lane + (c0 >> 1), pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        ++n_pv;
      }
      // epilogue: O / sum -> bf16 -> global
      mbar_wait(pv_done, (n_pv - 1) & 1);
      tc_fence_after();
      const float inv = 1.0f / sum;
      uint32_t ob[32];
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        uint32_t r0[32];
        tmem_ld_32x32b_x32(t_lane + O_COL + 32 * hcol, r0);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          ob[16 * hcol + j] = pack_bf16x2(__uint_as_float(r0[2 * j]) * inv, __uint_as_float(r0[2 * j + 1]) * inv);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (qrow < n) {
        uint4* op = reinterpret_cast<uint4*>(p.out + (size_t)(row0 + qrow) * p.I + h * DH);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_uint4(ob[4 * j], ob[4 * j + 1], ob[4 * j + 2], ob[4 * j + 3]);
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


// ---------------------------------------------------------------------------------------------------------------
// Pipelined variant: key blocks of 64 with THREE score buffers.  ONLINE = false (test hook 11 = 2): the same two-phase
// algorithm as above (exact max first).  ONLINE = true (the default): one pass -- exponentials are taken against a
// per-row reference max that moves only when a block exceeds it by 2^24, which rescales O in TMEM.
// In both, the three buffers mean that
// Q K^T of the next two blocks is issued before the softmax warps have finished block b and before P_b V_b has
// completed -- the MMA round trips (issue -> commit -> mbarrier -> wake-up, ~0.7 us each) that serialised every
// 128-key step of the kernel above overlap with the exponentials.  TMEM: S|P buffers at [0,64), [64,128) and
// [192,256) (P_b, bf16, in the first 32 columns of its own buffer), O at [128,192); 256 columns allocated, 2 CTAs per
// SM; 4 K/V stages of 64 keys.  Per-buffer barrier phases are tracked as bit masks (bit b = parity of buffer b).
// ---------------------------------------------------------------------------------------------------------------
namespace av2 {
constexpr int DH = 64;
constexpr int KB = 64;                       // keys per block
constexpr int Q_BYTES = 128 * 128;           // 128 rows x 64 bf16
constexpr int KV_BYTES = KB * 128;           // 64 rows x 64 bf16
constexpr int KV_STAGES = 4;
constexpr int NBUF = 3;                      // score buffers
constexpr int LOOKAHEAD = 2;                 // Q K^T is issued this many steps ahead of the softmax warps
constexpr int SMEM_DATA = Q_BYTES + 2 * KV_BYTES * KV_STAGES;
// q_full q_empty kv_full[4] kv_empty[4] s_full[3] s_free[3] p_ready[3] pv_done[3] o_full o_free
constexpr int NUM_BARS = 2 + 2 * KV_STAGES + 4 * NBUF + 2;
constexpr int UNIT_TAB = 128;                // units of this CTA located once, in parallel, at kernel start
constexpr int DYN_BYTES = SMEM_DATA + NUM_BARS * 8 + 16 + UNIT_TAB * 16 + 1024;
constexpr int TMEM_COLS = 256;
constexpr int O_COL = 128;
constexpr int THREADS = 6 * 32;
__device__ __forceinline__ uint32_t buf_col(int bf) { return bf == 2 ? 192u : static_cast<uint32_t>(bf) * KB; }
}  // namespace av2

template <bool TRACE, bool ONLINE>
__global__ void __launch_bounds__(av2::THREADS, 2)
attention_varlen2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         const AttnVarlenParams p) {
  using namespace av2;
  // %globaltimer stamp of slot `slot` of this CTA's `it`-th unit (CTA 0 only, first 64 units)
  auto stamp = [&](uint32_t it, int slot) {
    if (TRACE && blockIdx.x == 0 && it < 64) p.trace[it * 16 + slot] = (long long)globaltimer_ns();
  };
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + Q_BYTES;  // stage st: K at sKV + st*2*KV_BYTES, V right after
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_DATA);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + KV_STAGES;
  uint64_t* s_full = kv_empty + KV_STAGES;  // [NBUF]
  uint64_t* s_free = s_full + NBUF;         // [NBUF]  phase 1: the softmax warps have scanned the block
  uint64_t* p_ready = s_free + NBUF;        // [NBUF]
  uint64_t* pv_done = p_ready + NBUF;       // [NBUF]
  uint64_t* o_full = pv_done + NBUF;
  uint64_t* o_free = o_full + 1;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_free + 1);
  int4* unit_tab = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(bars) + NUM_BARS * 8 + 16);  // (seq, qt, row0, n)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int TMA_WARP = 4, MMA_WARP = 5;

  // Every role needs (sequence, query tile, first row, length) of each unit; the bisection of tile_prefix costs
  // ~0.6 us of dependent loads per unit and role, so the first UNIT_TAB units of this CTA are located here, one per
  // thread, and read back from shared memory (units beyond the table fall back to the bisection).
  for (int i = threadIdx.x; i < UNIT_TAB; i += blockDim.x) {
    const long long u = (long long)blockIdx.x + (long long)i * gridDim.x;
    if (u < p.units) {
      int seq, h, qt, row0, n;
      av_locate(p, (int)u, seq, h, qt, row0, n);
      unit_tab[i] = make_int4(seq, qt, row0, n);
    }
  }
  auto locate = [&](int u, uint32_t idx, int& seq, int& h, int& qt, int& row0, int& n) {
    if (idx < (uint32_t)UNIT_TAB) {
      const int4 e = unit_tab[idx];
      seq = e.x; qt = e.y; row0 = e.z; n = e.w;
      h = u % p.H;
    } else {
      av_locate(p, u, seq, h, qt, row0, n);
    }
  };

  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_free[b], 4);
      mbar_init(&p_ready[b], 4);
      mbar_init(&pv_done[b], 1);
    }
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
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
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t units_done = 0, kv_count = 0;
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++units_done) {
        int seq, h, qt, row0, n;
        stamp(units_done, 12);
        locate(u, units_done, seq, h, qt, row0, n);
        const int nb = (n + KB - 1) / KB;
        mbar_wait(q_empty, (units_done & 1) ^ 1);
        stamp(units_done, 13);
        mbar_arrive_expect_tx(q_full, Q_BYTES);
        tma_load_2d(sQ, &tmQ, q_full, h * DH, row0 + qt * 128);
        const int steps = (!ONLINE && nb > 1 ? nb : 0) + nb;
        for (int s = 0; s < steps; ++s, ++kv_count) {
          const bool phase2 = s >= steps - nb;
          const int kb = phase2 ? s - (steps - nb) : s;
          const int st = kv_count % KV_STAGES;
          mbar_wait(&kv_empty[st], ((kv_count / KV_STAGES) & 1) ^ 1);
          uint8_t* k_dst = sKV + st * 2 * KV_BYTES;
          mbar_arrive_expect_tx(&kv_full[st], phase2 ? 2 * KV_BYTES : KV_BYTES);
          tma_load_2d(k_dst, &tmKV, &kv_full[st], p.I + h * DH, row0 + kb * KB);
          if (phase2) tma_load_2d(k_dst + KV_BYTES, &tmKV, &kv_full[st], 2 * p.I + h * DH, row0 + kb * KB);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, KB, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, DH, 0, 1);
      const uint32_t sq = smem_u32(sQ);
      uint32_t units_done = 0;
      uint32_t sc = 0;   // global step counter: K/V stage = sc % KV_STAGES
      int bf_issue = 0;  // score buffer of the next Q K^T to issue (cycles 0, 1, 2)
      int bf_fin = 0;    // score buffer of the next step to finish (P V issue)
      // bit b of each mask = parity of the number of completions requested so far on buffer b's barrier
      uint32_t par_sfree = 0, par_pready = 0, par_pv = 0;
      uint32_t owes_sfree = 0, owes_pv = 0;  // bit b: the buffer's last user still has to signal s_free / pv_done
      for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++units_done) {
        int seq, h, qt, row0, n;
        locate(u, units_done, seq, h, qt, row0, n);
        const int nb = (n + KB - 1) / KB;
        const int steps = (!ONLINE && nb > 1 ? nb : 0) + nb;
        stamp(units_done, 8);
        mbar_wait(q_full, units_done & 1);
        stamp(units_done, 9);
        const uint32_t sc0 = sc;
        // issue Q K_b^T of step s (global step sc0 + s) into the next score buffer
        auto issue_s = [&](int s) {
          const uint32_t g = sc0 + s;
          const int bf = bf_issue, st = g % KV_STAGES;
          const uint32_t bit = 1u << bf;
          bf_issue = bf == NBUF - 1 ? 0 : bf + 1;
          const bool phase2 = s >= steps - nb;
          // (parity of the LAST requested completion = current parity bit ^ 1)
          if (owes_sfree & bit) mbar_wait(&s_free[bf], ((par_sfree >> bf) & 1) ^ 1);
          else if (owes_pv & bit) mbar_wait(&pv_done[bf], ((par_pv >> bf) & 1) ^ 1);
          owes_sfree &= ~bit;
          owes_pv &= ~bit;
          mbar_wait(&kv_full[st], (g / KV_STAGES) & 1);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(sq, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sKV + st * 2 * KV_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_ss(tmem_base + buf_col(bf), adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
          umma_commit(&s_full[bf]);
          if (s == steps - 1) umma_commit(q_empty);  // last use of the Q tile
          if (!phase2) {
            umma_commit(&kv_empty[st]);              // phase 1 stages hold K only
            par_sfree ^= bit;
            owes_sfree |= bit;
          } else {
            owes_pv |= bit;                          // until P_b V_b (issued below, later) has completed
          }
        };
        for (int s = 0; s < LOOKAHEAD && s < steps; ++s) issue_s(s);
        stamp(units_done, 10);
        for (int s = 0; s < steps; ++s) {
          if (s + LOOKAHEAD < steps) issue_s(s + LOOKAHEAD);  // ahead of the softmax warps
          const int bf = bf_fin;
          bf_fin = bf == NBUF - 1 ? 0 : bf + 1;
          const bool phase2 = s >= steps - nb;
          if (!phase2) continue;
          const uint32_t g = sc0 + s;
          const int st = g % KV_STAGES;
          const int kb = s - (steps - nb);
          if (kb == 0) mbar_wait(o_free, (units_done & 1) ^ 1);  // previous unit's O has been read
          mbar_wait(&p_ready[bf], (par_pready >> bf) & 1);
          par_pready ^= 1u << bf;
          tc_fence_after();
          const uint32_t sv = smem_u32(sKV + st * 2 * KV_BYTES + KV_BYTES);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            const uint64_t vdesc = make_smem_desc_sw128(sv + k * 2048, 1024, 1024);
            umma_ts(tmem_base + O_COL, tmem_base + buf_col(bf) + k * 8, vdesc, idesc_pv, (kb | k) != 0);
          }
          umma_commit(&kv_empty[st]);
          umma_commit(&pv_done[bf]);
          par_pv ^= 1u << bf;
          if (kb == nb - 1) umma_commit(o_full);
        }
        sc = sc0 + steps;
        stamp(units_done, 11);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroup (thread == query row)
    const int quad = warp & 3;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float c = p.scale_log2e;
    uint32_t units_done = 0;
    int bf = 0;                // score buffer of the next step (cycles 0, 1, 2, in step order like the MMA warp)
    uint32_t par_sfull = 0;    // bit b: parity of the next completion of s_full[b]
    auto next_buf = [&]() {
      par_sfull ^= 1u << bf;
      bf = bf == NBUF - 1 ? 0 : bf + 1;
    };
    // ONLINE only: pv_done[b] completes once per step on buffer b (barrier phases run on across units)
    uint32_t par_pv = 0;       // bit b: parity of the next completion of pv_done[b]
    int prev_bf = 0;
    uint32_t prev_par = 0;     // (buffer, parity) of the completion of the previous step's P V
    for (int u = blockIdx.x; u < p.units; u += gridDim.x, ++units_done) {
      const bool tr0 = TRACE && warp == 0 && lane == 0;
      if (tr0) stamp(units_done, 0);
      int seq, h, qt, row0, n;
      locate(u, units_done, seq, h, qt, row0, n);
      const int nb = (n + KB - 1) / KB;
      if (tr0) {
        stamp(units_done, 1);
        if (blockIdx.x == 0 && units_done < 64) p.trace[units_done * 16 + 7] = nb;
      }
      const int qrow = qt * 128 + quad * 32 + lane;
      float mx = -INFINITY;
      float sum = 0.f;
      if constexpr (ONLINE) {
        // ---- one pass: exponentials are taken against a REFERENCE max `mx` that is only moved when a block's max
        // exceeds it by more than 2^TAU (P is bf16 and sum / O are fp32: a stale reference costs range, not
        // precision); moving it rescales this warp's rows of O in TMEM, after the P V issued so far have completed.
        constexpr float TAU = 24.f;      // log2 units: P <= 2^24, sum and O stay far inside fp32 range
        for (int kb = 0; kb < nb; ++kb) {
          const uint32_t tb = t_lane + buf_col(bf);
          mbar_wait(&s_full[bf], (par_sfull >> bf) & 1);
          if (tr0 && kb == 0) stamp(units_done, 4);
          tc_fence_after();
          const int lim = n - kb * KB;   // valid keys of this block, >= 1 (warp-uniform)
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(tb, r0);
          if (lim > 32) tmem_ld_32x32b_x32(tb + 32, r1);
          tmem_ld_wait();
          float bm = -INFINITY;
          if (lim >= KB) {
#pragma unroll
            for (int j = 0; j < 32; ++j) bm = fmaxf(bm, fmaxf(__uint_as_float(r0[j]), __uint_as_float(r1[j])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < lim) bm = fmaxf(bm, __uint_as_float(r0[j]));
              if (j + 32 < lim) bm = fmaxf(bm, __uint_as_float(r1[j]));
            }
          }
          if (kb == 0) {
            mx = bm;
          } else {
            const bool grow = (bm - mx) * c > TAU;
            if (__any_sync(0xffffffffu, grow)) {
              const float f = grow ? fast_ex2((mx - bm) * c) : 1.f;
              if (grow) mx = bm;
              sum *= f;
              mbar_wait(&pv_done[prev_bf], prev_par);      // O holds every block before this one
              tc_fence_after();
#pragma unroll 1
              for (int c0 = 0; c0 < DH; c0 += 16) {
                uint32_t o[16];
                tmem_ld_32x32b_x16(t_lane + O_COL + c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                tmem_st_32x32b_x16(t_lane + O_COL + c0, o);
              }
            }
          }
          const float mc = mx * c;
          uint32_t pk[16];
          if (lim >= KB) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float e0 = fast_ex2(fmaf(__uint_as_float(r0[j]), c, -mc));
              const float e1 = fast_ex2(fmaf(__uint_as_float(r0[j + 1]), c, -mc));
              sum += e0 + e1;
              pk[j >> 1] = pack_bf16x2(e0, e1);
            }
            tmem_st_32x32b_x16(tb, pk);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float e0 = fast_ex2(fmaf(__uint_as_float(r1[j]), c, -mc));
              const float e1 = fast_ex2(fmaf(__uint_as_float(r1[j + 1]), c, -mc));
              sum += e0 + e1;
              pk[j >> 1] = pack_bf16x2(e0, e1);
            }
            tmem_st_32x32b_x16(tb + 16, pk);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float e0 = (j < lim) ? fast_ex2(fmaf(__uint_as_float(r0[j]), c, -mc)) : 0.f;
              const float e1 = (j + 1 < lim) ? fast_ex2(fmaf(__uint_as_float(r0[j + 1]), c, -mc)) : 0.f;
              sum += e0 + e1;
              pk[j >> 1] = pack_bf16x2(e0, e1);
            }
            tmem_st_32x32b_x16(tb, pk);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float e0 = (j + 32 < lim) ? fast_ex2(fmaf(__uint_as_float(r1[j]), c, -mc)) : 0.f;
              const float e1 = (j + 33 < lim) ? fast_ex2(fmaf(__uint_as_float(r1[j + 1]), c, -mc)) : 0.f;
              sum += e0 + e1;
              pk[j >> 1] = pack_bf16x2(e0, e1);
            }
            tmem_st_32x32b_x16(tb + 16, pk);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[bf]);
          prev_bf = bf;
          prev_par = (par_pv >> bf) & 1;
          par_pv ^= 1u << bf;
          next_buf();
        }
      } else {
      auto block_max = [&](int kb, uint32_t tb) {
#pragma unroll 1
        for (int c0 = 0; c0 < KB; c0 += 32) {
          if (kb * KB + c0 >= n) break;
          uint32_t r[32];
          tmem_ld_32x32b_x32(tb + c0, r);
          tmem_ld_wait();
          const int lim = n - (kb * KB + c0);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < lim) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
      };
      if (nb > 1) {
        for (int kb = 0; kb < nb; ++kb) {
          mbar_wait(&s_full[bf], (par_sfull >> bf) & 1);
          if (tr0 && kb == 0) stamp(units_done, 2);
          tc_fence_after();
          block_max(kb, t_lane + buf_col(bf));
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[bf]);
          next_buf();
        }
      }
      if (tr0) stamp(units_done, 3);
      for (int kb = 0; kb < nb; ++kb) {
        const uint32_t tb = t_lane + buf_col(bf);
        mbar_wait(&s_full[bf], (par_sfull >> bf) & 1);
        if (tr0 && kb == 0) stamp(units_done, 4);
        tc_fence_after();
        if (nb == 1) block_max(0, tb);
        const float mc = mx * c;
#pragma unroll 1
        for (int c0 = 0; c0 < KB; c0 += 32) {
          uint32_t r[32];
          uint32_t pk[16];
          const int lim = n - (kb * KB + c0);  // number of valid keys in this chunk (may be <= 0)
          if (lim > 0) {
            tmem_ld_32x32b_x32(tb + c0, r);
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float e0 = (j < lim) ? fast_ex2(fmaf(__uint_as_float(r[j]), c, -mc)) : 0.f;
            const float e1 = (j + 1 < lim) ? fast_ex2(fmaf(__uint_as_float(r[j + 1]), c, -mc)) : 0.f;
            sum += e0 + e1;
            pk[j >> 1] = pack_bf16x2(e0, e1);
          }
          tmem_st_32x32b_x16(tb + (c0 >> 1), pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[bf]);
        next_buf();
      }
      }
      // epilogue: O / sum -> bf16 -> global
      if (tr0) stamp(units_done, 5);
      mbar_wait(o_full, units_done & 1);
      if (tr0) stamp(units_done, 6);
      tc_fence_after();
      const float inv = 1.0f / sum;
      uint32_t ob[32];
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        uint32_t r0[32];
        tmem_ld_32x32b_x32(t_lane + O_COL + 32 * hcol, r0);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          ob[16 * hcol + j] = pack_bf16x2(__uint_as_float(r0[2 * j]) * inv, __uint_as_float(r0[2 * j + 1]) * inv);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
      if (qrow < n) {
        uint4* op = reinterpret_cast<uint4*>(p.out + (size_t)(row0 + qrow) * p.I + h * DH);
#pragma unroll
        for (int j = 0; j < 8; ++j) op[j] = make_uint4(ob[4 * j], ob[4 * j + 1], ob[4 * j + 2], ob[4 * j + 3]);
      }
      if (tr0) stamp(units_done, 14);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static long long* g_varlen_trace = nullptr;
void attention_varlen_set_trace(long long* buf) { g_varlen_trace = buf; }
// 0: pipelined 64-key blocks, one pass with a lazily moved reference max (default); 1: the serial 128-key-block kernel;
// 2: pipelined 64-key blocks, two passes (max first)
static int g_varlen_mode = 0;
void attention_varlen_set_mode(int v) { g_varlen_mode = v; }

}  // namespace b200

using namespace b200;

extern "C" int b200vit_attention_varlen(const void* qkv, void* out, const int32_t* cu_seqlens_dev,
                                        const int32_t* tile_prefix_dev, int num_seqs, int total_tokens, int total_tiles,
                                        int H, int dh, float scale, void* stream) {
  using namespace av;
  B200_CHECK_ARG(qkv && out && cu_seqlens_dev && tile_prefix_dev, "attention_varlen: null pointer");
  B200_CHECK_ARG(num_seqs > 0 && total_tokens > 0 && total_tiles > 0 && H > 0, "attention_varlen: bad shape");
  B200_CHECK_ARG(dh == DH, "attention_varlen: dim_head=%d not supported by this build (only 64)", dh);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 "attention_varlen: pointers must be 16-byte aligned");
  AttnVarlenParams p{};
  p.cu_seqlens = cu_seqlens_dev;
  p.tile_prefix = tile_prefix_dev;
  p.num_seqs = num_seqs;
  p.H = H;
  p.I = H * dh;
  p.units = total_tiles * H;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const uint64_t dims[2] = {(uint64_t)3 * p.I, (uint64_t)total_tokens};
  const uint64_t strides[1] = {(uint64_t)3 * p.I * 2};
  const int slots = 2 * num_sms();
  const int grid = p.units < slots ? p.units : slots;
  auto st = reinterpret_cast<cudaStream_t>(stream);
  if (g_varlen_mode != 1) {
    const bool online = g_varlen_mode == 0;
    CUtensorMap tmQ, tmKV;
    const uint32_t qbox[2] = {64, 128}, kvbox[2] = {64, (uint32_t)av2::KB};
    int rc = encode_tmap_bf16(&tmQ, qkv, 2, dims, strides, qbox);
    if (rc) return rc;
    rc = encode_tmap_bf16(&tmKV, qkv, 2, dims, strides, kvbox);
    if (rc) return rc;
    p.trace = g_varlen_trace;
    auto run = [&](auto kern) -> int {
      B200_ENSURE_SMEM(kern, av2::DYN_BYTES);
      kern<<<grid, av2::THREADS, av2::DYN_BYTES, st>>>(tmQ, tmKV, p);
      return 0;
    };
    if (p.trace) rc = online ? run(attention_varlen2_kernel<true, true>) : run(attention_varlen2_kernel<true, false>);
    else rc = online ? run(attention_varlen2_kernel<false, true>) : run(attention_varlen2_kernel<false, false>);
    if (rc) return rc;
  } else {
    CUtensorMap tm;
    const uint32_t box[2] = {64, 128};
    int rc = encode_tmap_bf16(&tm, qkv, 2, dims, strides, box);
    if (rc) return rc;
    B200_ENSURE_SMEM(attention_varlen_kernel, DYN_BYTES);
    attention_varlen_kernel<<<grid, THREADS, DYN_BYTES, st>>>(tm, p);
  }
  B200_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
