// tcgen05 probe: D[128x32] (TMEM, fp32) = A[128x32] (TMEM, fp16) * B[32x32]^T (smem, fp16, K-major, no swizzle).
// Verifies the descriptor / TMEM layout assumptions used by the DUNE tcgen05 kernel.  All waits are bounded.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc05_probe tc05_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int max_spin = 2000000) {
  for (int i = 0; i < max_spin; ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

#ifdef PROBE_F16D
constexpr uint32_t IDESC = (0u << 4) | (4u << 17) | (8u << 24);  // D=f16 (accumulators packed?), A=B=f16, K-major both, N=32, M=128
#else
constexpr uint32_t IDESC = (1u << 4) | (4u << 17) | (8u << 24);  // D=f32, A=B=f16, K-major both, N=32, M=128
#endif

__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  // K-major, SWIZZLE_NONE: core matrix = 8 rows (N) x 16 B (8 halves of K), 128 B contiguous;
  // SBO (next 8 rows of N) = 128 B, LBO (next 8 halves of K) = 512 B; version 1
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(512 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

__global__ void probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D, int* status) {
  __shared__ __align__(128) unsigned char sB[2048];  // 2 k-steps x 1024 B
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B[n][k] -> canonical layout
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    const int s = k / 16, kk = (k % 16) / 8, kr = k % 8, ng = n / 8, nr = n % 8;
    *reinterpret_cast<__half*>(sB + s * 1024 + kk * 512 + ng * 128 + nr * 16 + kr * 2) = B[n * 32 + k];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  // A row = tid: 32 halves -> 16 packed columns at column offset 32
  uint32_t a[16];
  for (int c = 0; c < 16; ++c) {
    const __half2 h = __halves2half2(A[tid * 32 + 2 * c], A[tid * 32 + 2 * c + 1]);
    a[c] = *reinterpret_cast<const uint32_t*>(&h);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(lane_base + 32),
               "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]),
               "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15])
               : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t sb = smem_u32(sB);
    for (int s = 0; s < 2; ++s) {
      const uint64_t bdesc = make_b_desc(sb + s * 1024);
      const uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
          ::"r"(tbase), "r"(tbase + 32 + 8 * s), "l"(bdesc), "r"(IDESC), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  const bool ok = mbar_wait(smem_u32(&mbar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) {
    if (tid == 0) *status = 1;
  } else {
    uint32_t d[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
          "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]),
          "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
        : "r"(lane_base)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < 32; ++n) D[tid * 32 + n] = __uint_as_float(d[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}

int main() {
  static __half hA[128 * 32], hB[32 * 32];
  static float hD[128 * 32], ref[128 * 32];
  for (int m = 0; m < 128; ++m) for (int k = 0; k < 32; ++k) hA[m * 32 + k] = __float2half((float)((m * 7 + k * 3) % 17 - 8) / 8.0f);
  for (int n = 0; n < 32; ++n) for (int k = 0; k < 32; ++k) hB[n * 32 + k] = __float2half((float)((n * 5 + k * 11) % 13 - 6) / 4.0f);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) {
    float s = 0; for (int k = 0; k < 32; ++k) s += __half2float(hA[m * 32 + k]) * __half2float(hB[n * 32 + k]);
    ref[m * 32 + n] = s;
  }
  __half *dA, *dB; float* dD; int* dS;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD)); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, sizeof(hD)); cudaMemset(dS, 0, 4);
  probe<<<1, 128>>>(dA, dB, dD, dS);
  cudaError_t e = cudaDeviceSynchronize();
  int st = -1; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  printf("cuda: %s, status %d\n", cudaGetErrorString(e), st);
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < 128 * 32; ++i) { double d = fabs(hD[i] - ref[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
  printf("max err %.3e, mismatches %d / %d\n", maxerr, bad, 128 * 32);
#ifdef PROBE_F16D
  // raw dump: every 32-bit TMEM column of rows 5 and 100 as two halves, next to the reference row
  for (int row : {5, 100}) {
    printf("row %d raw columns (lo half, hi half):\n", row);
    for (int c = 0; c < 32; ++c) {
      uint32_t w; memcpy(&w, &hD[row * 32 + c], 4);
      __half lo, hi; uint16_t l16 = w & 0xffff, h16 = w >> 16; memcpy(&lo, &l16, 2); memcpy(&hi, &h16, 2);
      printf("  col %2d: %9.4f %9.4f   ref[%2d] = %9.4f\n", c, __half2float(lo), __half2float(hi), c, ref[row * 32 + c]);
    }
  }
  return 0;
#endif
  printf("D[0][0..3] = %f %f %f %f   ref %f %f %f %f\n", hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
  printf("D[5][0..3] = %f %f %f %f   ref %f %f %f %f\n", hD[160], hD[161], hD[162], hD[163], ref[160], ref[161], ref[162], ref[163]);
  printf("D[100][28..31] = %f %f %f %f   ref %f %f %f %f\n", hD[3228], hD[3229], hD[3230], hD[3231], ref[3228], ref[3229], ref[3230], ref[3231]);
  return bad ? 2 : 0;
}
