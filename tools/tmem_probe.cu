// Micro-benchmarks behind the attention design: what one SM sustains on
//   (a) tcgen05.ld  TMEM -> registers, for 4 / 8 / 16 warps and x16 / x32 shapes,
//   (b) a chain of 13 accumulating tcgen05.mma 128 x 64 x 16 with A from TMEM (the P V product of one score tile), alone
//       and while other warps stream tcgen05.ld,
//   (c) the same chain with A from shared memory,
//   (d) 4 accumulating 128 x 208 x 16 MMAs (Q K^T of one tile).
// Data is garbage; only clock64() deltas of SM 0 matter.  Build: see tools/tmem_probe.sh.
#include <cstdio>
#include <vector>
#include "../vit_pytorch_b200/csrc/common.cuh"

using namespace b200;

__global__ void __launch_bounds__(1024, 1) ld_probe(long long* out, int warps, int reps, int shape) {
  __shared__ uint32_t tb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tb, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    for (int r = 0; r < reps; ++r) {
      if (shape == 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(base + ((r * 32) & 255), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      } else {
        uint32_t a[16], b[16], c[16];
        tmem_ld_32x32b_x16(base + ((r * 48) & 255), a);
        tmem_ld_32x32b_x16(base + ((r * 48 + 16) & 255), b);
        tmem_ld_32x32b_x16(base + ((r * 48 + 32) & 255), c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= a[i] ^ b[i] ^ c[i];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) out[1] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// mode 0: PV chain, A in TMEM.  1: PV chain, A in smem.  2: S (4 x N=208).  3: PV chain A in TMEM, two accumulators.
// ld_warps > 0: that many other warps stream tcgen05.ld meanwhile.
__global__ void __launch_bounds__(1024, 1) mma_probe(long long* out, int mode, int ld_warps, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tb;
  __shared__ uint64_t bar;
  __shared__ int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) { tmem_alloc(&tb, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tbase = tb;
  uint32_t acc = 0;
  if (warp == 31 && lane == 0) {
    const uint32_t sa = smem_u32(smem);            // A tile (K-major, 128 rows x 64 k, sw128): 16 KB blocks
    const uint32_t sb = smem_u32(smem + 64 * 1024);  // B tile
    const uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t idesc_pv_ss = make_idesc_bf16(128, 64, 0, 1);
    const uint32_t idesc_s = make_idesc_bf16(128, 208, 0, 0);
    long long tot = 0;
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      if (mode == 0 || mode == 3) {
        for (int k = 0; k < 13; ++k) {
          const uint64_t vdesc = make_smem_desc_sw128(sb + k * 2048, 1024, 1024);
          const uint32_t d = (mode == 3 && (k & 1)) ? tbase + 320 : tbase + 416;  // (mode 3: two separate 64-column tiles)
          umma_ts(d, tbase + k * 8, vdesc, idesc_pv, mode == 3 ? (k > 1) : (k != 0));
        }
      } else if (mode == 1) {
        for (int k = 0; k < 13; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(sa + (k >> 2) * 16384, 16, 1024) + 2 * (k & 3);
          const uint64_t vdesc = make_smem_desc_sw128(sb + k * 2048, 1024, 1024);
          umma_ss(tbase + 416, adesc, vdesc, idesc_pv_ss, k != 0);
        }
      } else {
        const uint64_t adesc = make_smem_desc_sw128(sa, 16, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(sb, 16, 1024);
        for (int k = 0; k < 4; ++k) umma_ss(tbase, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
      }
      umma_commit(&bar);
      mbar_wait(&bar, r & 1);
      tc_fence_after();
      tot += clock64() - t0;
    }
    if (blockIdx.x == 0) out[0] = tot / reps;
    stop = 1;
  } else if (warp < ld_warps) {
    const uint32_t base = tbase + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    volatile int* vstop = &stop;
    int it = 0;
    while (!*vstop) {
      uint32_t a[16], b[16], c[16];
      tmem_ld_32x32b_x16(base + ((it * 48) & 255), a);
      tmem_ld_32x32b_x16(base + ((it * 48 + 16) & 255), b);
      tmem_ld_32x32b_x16(base + ((it * 48 + 32) & 255), c);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc ^= a[i] ^ b[i] ^ c[i];
      ++it;
    }
  }
  if (acc == 0x12345678u) out[1] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  long long h[2];
  cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int shape : {16, 32}) {
    for (int warps : {4, 8, 16, 32}) {
      const int reps = 2000;
      ld_probe<<<148, 1024>>>(out, warps, reps, shape);
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * reps * (shape == 32 ? 32 : 48) * 32 * 4;
      printf("ld x%-2d %2d warps: %8lld cycles  %.1f B/clk/SM\n", shape, warps, h[0], bytes / h[0]);
    }
  }
  const char* names[] = {"PV 13x(128x64x16) A=TMEM", "PV 13x(128x64x16) A=smem", "S 4x(128x208x16)", "PV A=TMEM 2 accumulators"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int ldw : {0, 8, 16}) {
      mma_probe<<<148, 1024, 100 * 1024>>>(out, mode, ldw, 200);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      printf("%-28s with %2d ld warps: %6lld cycles issue->visible (%s)\n", names[mode], ldw, h[0], cudaGetErrorString(e));
    }
  }
  return 0;
}
