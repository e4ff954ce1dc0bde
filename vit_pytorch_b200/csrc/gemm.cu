// Persistent, warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] * W[N,K]^T)
//
//   warp 0        : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1        : MMA issuer    (one thread; tcgen05.mma cta_group::1 kind::f16, 128 x BLOCK_N x 16, D in TMEM)
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue      (tcgen05.ld 32x32b -> registers -> bias / LN-fold / GELU / residual -> global)
//
// Two TMEM accumulator stages (2 x BLOCK_N columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Replaces the nn.Linear call sites listed in include/b200vit.h.
#include "common.cuh"
#include "host_util.h"

namespace b200 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_NON_EPI_THREADS = 128;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = NUM_NON_EPI_THREADS + NUM_EPI_WARPS * 32;

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_k_blocks;
  int flags;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  long long ldo;
  const float* bias;
  const float* resid;
  const float* ln_sums;  // [M][ln_parts][2]
  int ln_parts;
  int stats_parts;
  float ln_inv_dim;
  float ln_eps;
  const float* col_s;  // [N]
  float* stats_out;    // [M][2]
  // Rows of A (and of the output) per tile: BLOCK_M, or -- patch mode -- the patches of `patch_ght` patch rows of one
  // image (<= 128; the rest of the 128-row MMA tile is ignored).
  int rows_per_tile;
  // Patch mode (b200vit_patch_embed_tma): A is not a matrix in memory but the NCHW image itself, read through a 5-D
  // tensor map (pixel 16 | patch column | patch row | pixel row 16 | image x channel).  k block kb = channel * 4 + g
  // covers pixel rows 4g .. 4g+3 of every patch.  With a swizzled tensor map every box row (16 pixels = 32 bytes)
  // lands on its own 128-byte shared-memory row (measured: tools/patch_tma_probe.py), so the four pixel rows of a
  // k block cannot share one operand row; they are loaded as four boxes into four 16 KB slabs, each a K-major
  // operand of which the tensor core reads the first 16 k (one tcgen05.mma k-step per slab).
  int patch;
  int patch_ght;            // patch rows per tile
  int patch_tiles_per_img;  // gh / patch_ght
  int patch_C;              // channels
};

// PATCH: the A stage holds FOUR 16 KB slabs, one per 16-wide k-step (see GemmParams::patch).
template <int BLOCK_N, int STAGES, bool PATCH = false>
struct GemmSmem {
  static constexpr int A_SLAB = BLOCK_M * BLOCK_K * 2;
  static constexpr int A_BYTES = PATCH ? 4 * A_SLAB : A_SLAB;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem base
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 4) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

template <int BLOCK_N, int STAGES, bool PATCH>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using L = GemmSmem<BLOCK_N, STAGES, PATCH>;
  constexpr int TMEM_COLS = 2 * BLOCK_N;  // 512 or 256 (power of two)
  static_assert(TMEM_COLS == 512 || TMEM_COLS == 256 || TMEM_COLS == 128, "TMEM columns must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_tiles;
        const int n_blk = tile % p.num_n_tiles;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (PATCH) {
            // im2col-free A tile: rows_per_tile patches x (4 pixel rows x 16 pixels) of channel kb / 4
            const int img = m_blk / p.patch_tiles_per_img, tin = m_blk % p.patch_tiles_per_img;
            mbar_arrive_expect_tx(&full_bar[stage], p.rows_per_tile * 128 + L::B_BYTES);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              tma_load_5d(sa + j * L::A_SLAB, &tmA, &full_bar[stage], 0, 0, tin * p.patch_ght, (kb & 3) * 4 + j,
                          img * p.patch_C + (kb >> 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          }
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t sb = sa + L::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle row: +2 in the (addr >> 4) field
            // (patch mode: the k-th 16 k of A are the first 32 bytes of the rows of slab k)
            const uint64_t ad = PATCH ? make_smem_desc_sw128(sa + k * L::A_SLAB, 16, 1024) : adesc + 2 * k;
            umma_ss(d_tmem, ad, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= NUM_NON_EPI_THREADS / 32) {
    // ------------------------------------------------------------------ epilogue
    const int e = warp - NUM_NON_EPI_THREADS / 32;  // 0..7
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    constexpr int COLS_PER_WARP = BLOCK_N / (NUM_EPI_WARPS / 4);
    const int col_off = (e >> 2) * COLS_PER_WARP;
    const int flags = p.flags;
    const bool vec_ok = (p.ldo & 7) == 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile % p.num_n_tiles;
      const int row = m_blk * p.rows_per_tile + quad * 32 + lane;
      const bool row_ok = row < p.M && quad * 32 + lane < p.rows_per_tile;
      float mu = 0.f, rstd = 1.f;
      if ((flags & B200VIT_EPI_LNFOLD) && row_ok) {
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < p.ln_parts; ++i) {
          const float2 ss = *reinterpret_cast<const float2*>(p.ln_sums + 2 * ((size_t)row * p.ln_parts + i));
          s1 += ss.x;
          s2 += ss.y;
        }
        mu = s1 * p.ln_inv_dim;
        const float var = fmaxf(s2 * p.ln_inv_dim - mu * mu, 0.f);
        rstd = rsqrtf(var + p.ln_eps);
      }
      float st_sum = 0.f, st_sq = 0.f;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + col_off;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_WARP; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c, r);
        tmem_ld_wait();
        const int col0 = n_blk * BLOCK_N + col_off + c;
        if (row_ok && col0 < p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int col = col0 + j;
            if (col >= p.N) continue;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[j + i]);
            const bool full8 = (col + 8 <= p.N);
            if (full8 && vec_ok) {
              if (flags & (B200VIT_EPI_LNFOLD | B200VIT_EPI_BIAS)) {
                // y = acc * rstd + (bias - rstd*mu * s): the same two FMAs as the CTA-pair kernel, so both kernels
                // give bit-identical results (batch-size invariance across the kernel switch at M = 1024)
                const bool fold = (flags & B200VIT_EPI_LNFOLD) != 0;
                float cv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (flags & B200VIT_EPI_BIAS) {
                  const float4 b0 = *reinterpret_cast<const float4*>(p.bias + col);
                  const float4 b1 = *reinterpret_cast<const float4*>(p.bias + col + 4);
                  cv[0] = b0.x; cv[1] = b0.y; cv[2] = b0.z; cv[3] = b0.w;
                  cv[4] = b1.x; cv[5] = b1.y; cv[6] = b1.z; cv[7] = b1.w;
                }
                if (fold) {
                  const float4 s0 = *reinterpret_cast<const float4*>(p.col_s + col);
                  const float4 s1 = *reinterpret_cast<const float4*>(p.col_s + col + 4);
                  const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                  const float k = -rstd * mu;
#pragma unroll
                  for (int i = 0; i < 8; ++i) cv[i] = fmaf(k, sv[i], cv[i]);
                }
                const float rs = fold ? rstd : 1.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], rs, cv[i]);
              }
              if (flags & B200VIT_EPI_GELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
              }
              if (flags & B200VIT_EPI_RESIDUAL) {
                const float* rp = p.resid + (size_t)row * p.ldo + col;
                const float4 r0 = *reinterpret_cast<const float4*>(rp);
                const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
                v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
              }
              if (p.out_f32) {
                float* op = p.out_f32 + (size_t)row * p.ldo + col;
                *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
              uint4 pk;
              pk.x = pack_bf16x2(v[0], v[1]);
              pk.y = pack_bf16x2(v[2], v[3]);
              pk.z = pack_bf16x2(v[4], v[5]);
              pk.w = pack_bf16x2(v[6], v[7]);
              if (p.out_bf16) *reinterpret_cast<uint4*>(p.out_bf16 + (size_t)row * p.ldo + col) = pk;
              if (flags & B200VIT_EPI_STATS) {
                const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float lo = __uint_as_float(w[i] << 16);
                  const float hi = __uint_as_float(w[i] & 0xFFFF0000u);
                  st_sum += lo + hi;
                  st_sq = fmaf(lo, lo, fmaf(hi, hi, st_sq));
                }
              }
            } else {
              // scalar tail (N or ldo not a multiple of 8)
              for (int i = 0; i < 8 && col + i < p.N; ++i) {
                float x = v[i];
                const int cc = col + i;
                if (flags & (B200VIT_EPI_LNFOLD | B200VIT_EPI_BIAS)) {
                  float cvs = (flags & B200VIT_EPI_BIAS) ? p.bias[cc] : 0.f;
                  if (flags & B200VIT_EPI_LNFOLD) cvs = fmaf(-rstd * mu, p.col_s[cc], cvs);
                  x = fmaf(x, (flags & B200VIT_EPI_LNFOLD) ? rstd : 1.0f, cvs);
                }
                if (flags & B200VIT_EPI_GELU) x = gelu_erf(x);
                if (flags & B200VIT_EPI_RESIDUAL) x += p.resid[(size_t)row * p.ldo + cc];
                if (p.out_f32) p.out_f32[(size_t)row * p.ldo + cc] = x;
                const __nv_bfloat16 xb = __float2bfloat16_rn(x);
                if (p.out_bf16) p.out_bf16[(size_t)row * p.ldo + cc] = xb;
                if (flags & B200VIT_EPI_STATS) {
                  const float xr = __bfloat162float(xb);
                  st_sum += xr;
                  st_sq = fmaf(xr, xr, st_sq);
                }
              }
            }
          }
        }
      }
      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if ((flags & B200VIT_EPI_STATS) && row_ok) {
        const int part = n_blk * (NUM_EPI_WARPS / 4) + (e >> 2);
        *reinterpret_cast<float2*>(p.stats_out + 2 * ((size_t)row * p.stats_parts + part)) =
            make_float2(st_sum, st_sq);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int STAGES, bool PATCH = false>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, cudaStream_t stream) {
  using L = GemmSmem<BLOCK_N, STAGES, PATCH>;
  static_assert(L::DYN_BYTES <= 227 * 1024, "gemm: shared memory budget");
  auto kern = gemm_bf16_kernel<BLOCK_N, STAGES, PATCH>;
  B200_ENSURE_SMEM(kern, L::DYN_BYTES);
  if (!p.patch) p.rows_per_tile = BLOCK_M;
  p.num_m_tiles = (p.M + p.rows_per_tile - 1) / p.rows_per_tile;
  p.num_n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  p.num_k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, NUM_THREADS, L::DYN_BYTES, stream>>>(tmA, tmB, p);
  B200_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace b200

extern "C" int b200vit_stats_parts(int N) {
  const int block_n = N > 128 ? 256 : 128;      // both GEMM kernels: two column halves per N tile
  return 2 * ((N + block_n - 1) / block_n);
}

extern "C" int b200vit_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out_bf16,
                                 float* out_f32, int64_t ldo, const float* bias, const float* resid,
                                 const float* ln_sums, int ln_parts, float ln_eps, const float* col_s,
                                 float* stats_out, int M, int N, int K, int flags, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(A && W, "gemm: A/W must not be null");
  B200_CHECK_ARG(out_bf16 || out_f32, "gemm: need at least one output");
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                 "gemm: A and W must be 16-byte aligned");
  B200_CHECK_ARG((lda & 7) == 0 && (ldw & 7) == 0 && lda >= K && ldw >= K,
                 "gemm: lda=%lld ldw=%lld must be multiples of 8 and >= K=%d", (long long)lda, (long long)ldw, K);
  B200_CHECK_ARG(ldo >= N, "gemm: ldo=%lld < N=%d", (long long)ldo, N);
  B200_CHECK_ARG(!(flags & B200VIT_EPI_BIAS) || bias, "gemm: EPI_BIAS without bias");
  B200_CHECK_ARG(!(flags & B200VIT_EPI_RESIDUAL) || resid, "gemm: EPI_RESIDUAL without resid");
  B200_CHECK_ARG(!(flags & B200VIT_EPI_LNFOLD) || (ln_sums && col_s && ln_parts >= 1 && ln_parts <= 64),
                 "gemm: EPI_LNFOLD needs ln_sums, col_s and 1 <= ln_parts <= 64");
  B200_CHECK_ARG(!(flags & B200VIT_EPI_STATS) || stats_out, "gemm: EPI_STATS without stats_out");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  B200_CHECK_ARG(al16(bias) && al16(resid) && al16(col_s) && al16(out_bf16) && al16(out_f32) && al16(ln_sums),
                 "gemm: epilogue pointers must be 16-byte aligned");

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (gemm2_eligible(M, N, K, ldo, flags, out_bf16, out_f32, resid))
    return launch_gemm2(A, lda, W, ldw, out_bf16, out_f32, ldo, bias, resid, ln_sums, ln_parts, ln_eps, col_s,
                        stats_out, M, N, K, flags, st);

  // K-tail: TMA zero-fills out-of-bounds columns of both operands, so any K works as long as rows are 16B multiples.
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.flags = flags;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.out_f32 = out_f32;
  p.ldo = ldo;
  p.bias = bias;
  p.resid = resid;
  p.ln_sums = ln_sums;
  p.ln_parts = ln_parts;
  p.stats_parts = b200vit_stats_parts(N);
  p.ln_inv_dim = 1.0f / (float)K;
  p.ln_eps = ln_eps;
  p.col_s = col_s;
  p.stats_out = stats_out;

  const bool wide = N > 128;
  const uint32_t block_n = wide ? 256 : 128;
  CUtensorMap tmA, tmB;
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
    const uint32_t box[2] = {(uint32_t)BLOCK_K, block_n};
    int rc = encode_tmap_bf16(&tmB, W, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (wide) return launch_gemm<256, 4>(tmA, tmB, p, st);
  return launch_gemm<128, 6>(tmA, tmB, p, st);
}

extern "C" int b200vit_rmsnorm_heads(void* buf, int64_t ld, const float* gamma, int T, int nheads, int dh, void* stream);
extern "C" int b200vit_layernorm_heads(void* buf, int64_t ld, const float* gamma, int T, int nheads, int dh, float eps,
                                       void* stream);

extern "C" int b200vit_gemm_headnorm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out_bf16,
                                          int64_t ldo, const float* bias, const float* ln_sums, int ln_parts,
                                          float ln_eps, const float* col_s, const float* head_gamma, int norm_heads,
                                          int dh, float head_eps, int M, int N, int K, int flags, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(out_bf16 && head_gamma, "gemm_headnorm: null pointer");
  B200_CHECK_ARG(dh == 64, "gemm_headnorm: dim_head=%d not supported by this build (only 64)", dh);
  B200_CHECK_ARG(norm_heads > 0 && norm_heads * 64 <= N, "gemm_headnorm: %d heads do not fit N=%d", norm_heads, N);
  B200_CHECK_ARG((flags & ~(B200VIT_EPI_BIAS | B200VIT_EPI_LNFOLD | B200VIT_EPI_HEADLN)) == 0,
                 "gemm_headnorm: unsupported flags %d", flags);
  const bool hln = (flags & B200VIT_EPI_HEADLN) != 0;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(head_gamma) & 15) == 0, "gemm_headnorm: head_gamma must be 16-byte aligned");
  // fused: the CTA-pair kernel's bf16 epilogue, where one warp holds one whole head of its 32 rows
  if (M > 0 && N > 0 && K > 0 && A && W && gemm2_eligible(M, N, K, ldo, flags, out_bf16, nullptr, nullptr) &&
      (ldo % 8) == 0) {
    B200_CHECK_ARG(!(flags & B200VIT_EPI_BIAS) || bias, "gemm_headnorm: EPI_BIAS without bias");
    B200_CHECK_ARG(!(flags & B200VIT_EPI_LNFOLD) || (ln_sums && col_s && ln_parts >= 1 && ln_parts <= 64),
                   "gemm_headnorm: EPI_LNFOLD needs ln_sums, col_s and 1 <= ln_parts <= 64");
    B200_CHECK_ARG((lda & 7) == 0 && (ldw & 7) == 0 && lda >= K && ldw >= K && ldo >= N, "gemm_headnorm: bad strides");
    return launch_gemm2(A, lda, W, ldw, out_bf16, nullptr, ldo, bias, nullptr, ln_sums, ln_parts, ln_eps, col_s,
                        nullptr, M, N, K, flags | B200VIT_EPI_HEADNORM, reinterpret_cast<cudaStream_t>(stream),
                        head_gamma, norm_heads * 64, head_eps);
  }
  int rc = b200vit_gemm_bf16(A, lda, W, ldw, out_bf16, nullptr, ldo, bias, nullptr, ln_sums, ln_parts, ln_eps, col_s,
                             nullptr, M, N, K, flags & ~B200VIT_EPI_HEADLN, stream);
  if (rc) return rc;
  if (hln) return b200vit_layernorm_heads(out_bf16, ldo, head_gamma, M, norm_heads, dh, head_eps, stream);
  return b200vit_rmsnorm_heads(out_bf16, ldo, head_gamma, M, norm_heads, dh, stream);
}


// ------------------------------------------------------------------------------------------------------------------
// im2col-free patch embedding (north star: "stages p x p x 3 pixels through TMA into shared memory and feeds a tcgen05
// GEMM"): the patch projection reads the NCHW image directly; Rearrange('b c (h p1) (w p2) -> b (h w) (p1 p2 c)') and
// the LayerNorm over the patch (vit.py:100-101) never materialise.  LayerNorm is folded exactly like everywhere else:
//   LN(x) W^T + b  =  rstd (x (gamma W)^T - mu colsum) + (W beta + b),
// x = raw bf16 pixels (the A operand), (mu, rstd) from b200vit_patch_stats.  The K order of the operand is
// (c, p1, p2) -- the image's own order -- so the caller permutes the weight's columns from the reference's (p1 p2 c).
// ------------------------------------------------------------------------------------------------------------------
extern "C" int b200vit_patch_embed_tma(const void* img, const void* w_perm, const float* bias, const float* col_s,
                                       const float* patch_stats, float ln_eps, float* out_f32, int64_t ldo, int B, int C,
                                       int H, int W, int D, void* stream) {
  using namespace b200;
  B200_CHECK_ARG(img && w_perm && bias && col_s && patch_stats && out_f32, "patch_embed_tma: null pointer");
  B200_CHECK_ARG(B > 0 && C > 0 && C <= 8 && H > 0 && W > 0 && (H % 16) == 0 && (W % 16) == 0,
                 "patch_embed_tma: needs 16 x 16 patches on an image whose sides are multiples of 16 (got %dx%d)", H, W);
  B200_CHECK_ARG(D > 0 && (D % 8) == 0 && ldo >= D, "patch_embed_tma: bad D=%d / ldo", D);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_perm) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(col_s) & 15) == 0,
                 "patch_embed_tma: pointers must be 16-byte aligned");
  const int gh = H / 16, gw = W / 16;
  B200_CHECK_ARG(gw <= 128, "patch_embed_tma: %d patches per row do not fit a 128-row tile", gw);
  int ght = 1;  // largest divisor of gh whose patch rows fit the 128-row MMA tile
  for (int d = 1; d <= gh; ++d)
    if (gh % d == 0 && d * gw <= 128) ght = d;
  const int K = C * 256;
  GemmParams p{};
  p.M = B * gh * gw; p.N = D; p.K = K;
  p.flags = B200VIT_EPI_LNFOLD | B200VIT_EPI_BIAS;
  p.out_f32 = out_f32;
  p.ldo = ldo;
  p.bias = bias;
  p.ln_sums = patch_stats;
  p.ln_parts = 1;
  p.stats_parts = b200vit_stats_parts(D);
  p.ln_inv_dim = 1.0f / (float)K;
  p.ln_eps = ln_eps;
  p.col_s = col_s;
  p.patch = 1;
  p.patch_ght = ght;
  p.patch_tiles_per_img = gh / ght;
  p.patch_C = C;
  p.rows_per_tile = ght * gw;
  CUtensorMap tmA, tmB;
  {
    // innermost first: pixel in a patch row | patch column | patch row | pixel row in the patch | image x channel
    const uint64_t dims[5] = {16, (uint64_t)gw, (uint64_t)gh, 16, (uint64_t)B * C};
    const uint64_t strides[4] = {32, (uint64_t)16 * W * 2, (uint64_t)W * 2, (uint64_t)H * W * 2};
    const uint32_t box[5] = {16, (uint32_t)gw, (uint32_t)ght, 1, 1};
    int rc = encode_tmap_bf16(&tmA, img, 5, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)D};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, 256};
    int rc = encode_tmap_bf16(&tmB, w_perm, 2, dims, strides, box);
    if (rc) return rc;
  }
  return launch_gemm<256, 2, true>(tmA, tmB, p, reinterpret_cast<cudaStream_t>(stream));
}
