// DUNE kernel, tcgen05 version (Blackwell 5th-gen tensor cores, accumulators in TMEM).
//
// Same contract as dune_mma_kernel.cuh / dune_kernel.cuh.  Mapping: CTA = 128 threads = 128 TMEM lanes;
// thread r owns point r of a 128-point tile for the whole network, so LayerNorm statistics, tanh, ReLU
// and the distance are thread-local (no shuffles, no fragment bookkeeping).  Per dense layer:
//   registers --split x = hi + lo (fp16)--> tcgen05.st (A operand, TMEM)         [all threads]
//   D[128 x 32] = bias (tcgen05.st) + A_lo.B_hi + A_hi.B_lo + A_hi.B_hi           [one thread, tcgen05.mma,
//                                                               B from shared memory via UMMA descriptors]
//   tcgen05.commit -> mbarrier -> tcgen05.ld (32 fp32 columns = the thread's row)  [all threads]
// Layouts / descriptors were verified in isolation with tools/tc05_probe.cu.
//
// TMEM columns per CTA (64 allocated): D [0,32)  A_hi [32,48)  A_lo [48,64).  D is pre-loaded with the bias row by
// tcgen05.st (every thread writes the same 32 floats into its lane), all MMAs then accumulate.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "dune_kernel.cuh"  // DuneParams, orderable(), flow()

namespace nb {

// Weight image built on the host (dune_tc.cu): canonical K-major / no-swizzle UMMA operand layout
// (core matrix = 8 rows x 16 B; SBO = 128 B between 8-row groups; LBO between the two 8-half K groups).
struct TcImage {
  static constexpr int kHiddenStride = 4096;         // per hidden layer: W_hi 2048 | W_lo 2048
  static constexpr int kHeadOff = 4 * kHiddenStride;  // head (N padded to 16): W_hi 1024 | W_lo 1024
  static constexpr int kFloatOff = kHeadOff + 2048;   // 18432
  // float section (offsets in floats): W0 (32x2), b0, LayerNorm gain/offset pre-multiplied by 2*log2(e), then the
  // biases of the four hidden layers and of the head (16, zero padded) -- already including the folded tanh map
  static constexpr int W0 = 0, B0 = 64, G1 = 96, BE1 = 128, G6 = 160, BE6 = 192, G11 = 224, BE11 = 256, BH = 288, BHEAD = 416,
                       kFloats = 432;
  static constexpr int kBytes = kFloatOff + kFloats * 4;  // 20160
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr uint32_t kIdescN32 = (1u << 4) | (4u << 17) | (8u << 24);  // D f32, A/B f16 K-major, N = 32, M = 128
constexpr uint32_t kIdescN16 = (1u << 4) | (2u << 17) | (8u << 24);  // N = 16

__device__ __forceinline__ uint64_t b_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 26); ++i) {
    uint32_t ok;
    // the suspend-time hint lets the hardware park the thread until the phase completes (or the time limit passes)
    // instead of returning immediately: fewer polling instructions competing for issue slots
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity), "r"(20000u)
                 : "memory");
    if (ok) return;
  }
  __trap();  // never hang the device on a protocol error
}

__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&a)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(a[0]),
               "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]),
               "r"(a[13]), "r"(a[14]), "r"(a[15])
               : "memory");
}

// every thread writes the same 32-float row (the layer's bias) into its TMEM lane: D := bias
__device__ __forceinline__ void st_bias32(uint32_t taddr, const float* __restrict__ b) {
  uint32_t v[32];
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const uint4 q = *reinterpret_cast<const uint4*>(b + 4 * j4);
    v[4 * j4] = q.x; v[4 * j4 + 1] = q.y; v[4 * j4 + 2] = q.z; v[4 * j4 + 3] = q.w;
  }
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
      "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
      "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void st_bias16(uint32_t taddr, const float* __restrict__ b) {
  uint32_t v[16];
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const uint4 q = *reinterpret_cast<const uint4*>(b + 4 * j4);
    v[4 * j4] = q.x; v[4 * j4 + 1] = q.y; v[4 * j4 + 2] = q.z; v[4 * j4 + 3] = q.w;
  }
  st16(taddr, v);
}

__device__ __forceinline__ void ld32(uint32_t taddr, float (&h)[32]) {
  uint32_t d[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
        "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]),
        "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) h[j] = __uint_as_float(d[j]);
}

__device__ __forceinline__ void ld8(uint32_t taddr, float (&h)[8]) {
  uint32_t d[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = __uint_as_float(d[j]);
}

// x = hi + lo (fp16 each), packed two values per 32-bit TMEM column (even k in the low half)
__device__ __forceinline__ void split32(const float (&h)[32], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const __half2 hh = __floats2half2_rn(h[2 * c], h[2 * c + 1]);
    const float2 f = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(h[2 * c] - f.x, h[2 * c + 1] - f.y);
    hi[c] = *reinterpret_cast<const uint32_t*>(&hh);
    lo[c] = *reinterpret_cast<const uint32_t*>(&ll);
  }
}

// r = 1/(exp2(a) + 1) with a = 2*log2(e)*y; tanh(y) = 1 - 2r.  The affine map 1 - 2r is folded into the weights and
// bias of the layer that follows (W' = -2W, b' = b + rowsum(W), built on the host), so r itself is the activation.
__device__ __forceinline__ float tanh_r(float a) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  const float d = e + 1.0f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
}

// thread-local LayerNorm (eps 1e-5, biased variance) + tanh; g / be pre-scaled by 2*log2(e)
__device__ __forceinline__ void ln_tanh32(float (&h)[32], const float* __restrict__ g, const float* __restrict__ be) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += h[j];
  const float mean = s * (1.0f / 32);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    h[j] -= mean;
    q = fmaf(h[j], h[j], q);
  }
  const float r = rsqrtf(fmaf(q, 1.0f / 32, 1e-5f));
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 gg = *reinterpret_cast<const float4*>(g + 4 * j4);
    const float4 bb = *reinterpret_cast<const float4*>(be + 4 * j4);
    h[4 * j4 + 0] = tanh_r(fmaf(h[4 * j4 + 0] * r, gg.x, bb.x));
    h[4 * j4 + 1] = tanh_r(fmaf(h[4 * j4 + 1] * r, gg.y, bb.y));
    h[4 * j4 + 2] = tanh_r(fmaf(h[4 * j4 + 2] * r, gg.z, bb.z));
    h[4 * j4 + 3] = tanh_r(fmaf(h[4 * j4 + 3] * r, gg.w, bb.w));
  }
}

}  // namespace tc

__host__ __device__ inline size_t dune_tc_smem_bytes(int N, int E, int M) {
  return TcImage::kBytes + (size_t)N * 8 + (((size_t)N * E * 4 + 7) / 8) * 8 + (size_t)4 * M * 8 + 64;
}

__global__ void __launch_bounds__(128, 5) dune_tc_kernel(const DuneParams prm, const unsigned char* __restrict__ image) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];  // the attribute aligns the dynamic segment (UMMA operands need 128 B)
  using I = TcImage;
  unsigned char* simg = smem_dyn;  // operand image
  const float* fl = reinterpret_cast<const float*>(simg + I::kFloatOff);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(simg + I::kBytes);
  float* smu = reinterpret_cast<float*>(simg + I::kBytes + (size_t)prm.N * 8);
  unsigned long long* cands = reinterpret_cast<unsigned long long*>(simg + I::kBytes + (size_t)prm.N * 8 + (((size_t)prm.N * prm.geo.E * 4 + 7) / 8) * 8);
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < I::kBytes / 16; i += 128) reinterpret_cast<uint4*>(simg)[i] = reinterpret_cast<const uint4*>(image)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // operand image -> visible to the tensor core (async proxy)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(tc::smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
  const uint32_t tD = tbase, tAhi = tbase + 32, tAlo = tbase + 48;
  const uint32_t simg_u = tc::smem_u32(simg);
  const uint32_t bar = tc::smem_u32(&mbar);
  uint32_t phase = 0;

  // one dense layer on the tensor core: h (activations) -> h (pre-activations incl. bias)
  auto dense = [&](float (&h)[32], int layer) {
    uint32_t hi[16], lo[16];
    tc::split32(h, hi, lo);
    tc::st16(trow + 32, hi);
    tc::st16(trow + 48, lo);
    tc::st_bias32(trow, fl + I::BH + 32 * layer);  // D := bias; every MMA below accumulates
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t w = simg_u + layer * I::kHiddenStride;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const uint64_t bhi = tc::b_desc(w + s * 1024, 512), blo = tc::b_desc(w + 2048 + s * 1024, 512);
        tc::mma_f16(tD, tAlo + 8 * s, bhi, tc::kIdescN32, 1u);
        tc::mma_f16(tD, tAhi + 8 * s, blo, tc::kIdescN32, 1u);
        tc::mma_f16(tD, tAhi + 8 * s, bhi, tc::kIdescN32, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tc::ld32(trow, h);
  };
  auto head = [&](float (&h)[32], float (&mu)[8]) {
    uint32_t hi[16], lo[16];
    tc::split32(h, hi, lo);
    tc::st16(trow + 32, hi);
    tc::st16(trow + 48, lo);
    tc::st_bias16(trow, fl + I::BHEAD);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t w = simg_u + I::kHeadOff;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const uint64_t bhi = tc::b_desc(w + s * 512, 256), blo = tc::b_desc(w + 1024 + s * 512, 256);
        tc::mma_f16(tD, tAlo + 8 * s, bhi, tc::kIdescN16, 1u);
        tc::mma_f16(tD, tAhi + 8 * s, blo, tc::kIdescN16, 1u);
        tc::mma_f16(tD, tAhi + 8 * s, bhi, tc::kIdescN16, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tc::ld8(trow, mu);
  };

  const int T1 = prm.T + 1, N = prm.N, M = prm.M, E = prm.geo.E;
  const int items = prm.B * T1;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / T1, t = item - b * T1;
    if (prm.active && prm.active[b] == 0) continue;  // uniform per CTA
    int n = prm.num_points ? prm.num_points[b] : N;
    n = n < 0 ? 0 : (n > N ? N : n);
    const int cnt = n < M ? n : M;
    if (t == 0 && tid == 0) {
      prm.sel_count[b] = cnt;
      if (n == 0 && prm.min_dist) prm.min_dist[b] = __int_as_float(0x7f800000);
    }
    if (n == 0) continue;

    const float* ns = prm.nom_s + (size_t)b * 3 * T1;
    const float sx = ns[t], sy = ns[T1 + t], th = ns[2 * T1 + t];
    const float cs = cosf(th), sn = sinf(th);
    const float* px = prm.points + (size_t)b * 2 * N;
    const float* py = px + N;
    const float* vx = prm.velocities ? prm.velocities + (size_t)b * 2 * N : nullptr;
    const float* vy = vx ? vx + N : nullptr;

    // ---- phase 1: 128 points per tile, one per thread ------------------------------------------------
#pragma unroll 1
    for (int tile = 0; tile * 128 < n; ++tile) {
      int i = tile * 128 + tid;
      const bool valid = i < n;
      i = valid ? i : n - 1;
      float gx = px[i], gy = py[i];
      if (vx) {
        gx = flow(gx, vx[i], prm.dt, t);
        gy = flow(gy, vy[i], prm.dt, t);
      }
      const float dx = gx - sx, dy = gy - sy;  // p0 = R^T (p_t - trans)   (pan.py:210)
      const float x0 = fmaf(cs, dx, sn * dy), y0 = fmaf(cs, dy, -(sn * dx));
      float h[32];
#pragma unroll
      for (int j2 = 0; j2 < 16; ++j2) {  // layer 0 on the FMA pipe
        const float4 w = *reinterpret_cast<const float4*>(fl + I::W0 + 4 * j2);  // W0[2j][0..1], W0[2j+1][0..1]
        const float2 bb = *reinterpret_cast<const float2*>(fl + I::B0 + 2 * j2);
        h[2 * j2] = fmaf(w.y, y0, fmaf(w.x, x0, bb.x));
        h[2 * j2 + 1] = fmaf(w.w, y0, fmaf(w.z, x0, bb.y));
      }
      tc::ln_tanh32(h, fl + I::G1, fl + I::BE1);
      dense(h, 0);
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
      dense(h, 1);
      tc::ln_tanh32(h, fl + I::G6, fl + I::BE6);
      dense(h, 2);
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
      dense(h, 3);
      tc::ln_tanh32(h, fl + I::G11, fl + I::BE11);
      float mu[8];
      head(h, mu);
      float d = 0.f;  // dist = mu^T (G p0 - h)   (dune.py:119-122)
#pragma unroll
      for (int e = 0; e < kMaxEdges; ++e) {
        if (e < E) {
          mu[e] = fmaxf(mu[e], 0.f);
          const float ge = fmaf(prm.geo.G[e][1], y0, prm.geo.G[e][0] * x0) - prm.geo.h[e];
          d = fmaf(mu[e], ge, d);
          if (valid) smu[i * E + e] = mu[e];
        }
      }
      if (valid) keys[i] = ((unsigned long long)orderable(d) << 32) | (unsigned)i;
    }
    __syncthreads();

    // ---- phase 2: top-M (ascending, ties -> lower index): per-warp REDUX rounds, then warp 0 merges ----
    unsigned long long mine = ~0ull;
    {
      unsigned long long* cand = cands + warp * M;
      for (int m = 0; m < cnt; ++m) {
        unsigned bd = 0xFFFFFFFFu, bi = 0xFFFFFFFFu;
        for (int i = warp * 32 + lane; i < n; i += 128) {  // the keys of rows this warp computed
          const uint2 k = *reinterpret_cast<const uint2*>(keys + i);
          if (k.y < bd) { bd = k.y; bi = k.x; }
        }
        const unsigned md = __reduce_min_sync(0xffffffffu, bd);
        const unsigned mi = __reduce_min_sync(0xffffffffu, bd == md ? bi : 0xFFFFFFFFu);
        if (md != 0xFFFFFFFFu && bd == md && bi == mi) keys[mi] = ~0ull;
        if (lane == 0) cand[m] = md == 0xFFFFFFFFu ? ~0ull : (((unsigned long long)md << 32) | mi);
        __syncwarp();
      }
    }
    __syncthreads();
    if (warp == 0) {
      const int total = 4 * cnt;
      for (int m = 0; m < cnt; ++m) {
        unsigned bd = 0xFFFFFFFFu, bi = 0xFFFFFFFFu;
        int bpos = -1;
        for (int c = lane; c < total; c += 32) {
          const int w = c / cnt, r = c - w * cnt;
          const uint2 k = *reinterpret_cast<const uint2*>(cands + w * M + r);
          if (k.y < bd || (k.y == bd && k.x < bi)) { bd = k.y; bi = k.x; bpos = w * M + r; }
        }
        const unsigned md = __reduce_min_sync(0xffffffffu, bd);
        const unsigned mi = __reduce_min_sync(0xffffffffu, bd == md ? bi : 0xFFFFFFFFu);
        if (bpos >= 0 && bd == md && bi == mi) cands[bpos] = ~0ull;
        if (lane == m) mine = ((unsigned long long)md << 32) | mi;
        __syncwarp();
      }
      // ---- phase 3: lane m writes the m-th closest point ---------------------------------------------
      if (lane < cnt) {
        unsigned idx = (unsigned)(mine & 0xffffffffull);
        if (idx >= (unsigned)n) idx = 0;
        uint32_t u = (uint32_t)(mine >> 32);
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        const float d = __uint_as_float(u);
        float gx = px[idx], gy = py[idx];
        if (vx) {
          gx = flow(gx, vx[idx], prm.dt, t);
          gy = flow(gy, vy[idx], prm.dt, t);
        }
        const size_t o = ((size_t)b * T1 + t) * M + lane;
        float lx = 0.f, ly = 0.f;
#pragma unroll
        for (int e = 0; e < kMaxEdges; ++e) {  // lam = ((-R) G^T) mu   (dune.py:89); constant indices keep geo in the constant bank
          if (e < E) {
            const float m_e = smu[idx * E + e];
            lx = fmaf(fmaf(sn, prm.geo.G[e][1], -cs * prm.geo.G[e][0]), m_e, lx);
            ly = fmaf(fmaf(-cs, prm.geo.G[e][1], -sn * prm.geo.G[e][0]), m_e, ly);
            prm.sel_mu[o * E + e] = m_e;
          }
        }
        prm.sel_lam[o * 2 + 0] = lx; prm.sel_lam[o * 2 + 1] = ly;
        prm.sel_pts[o * 2 + 0] = gx; prm.sel_pts[o * 2 + 1] = gy;
        prm.sel_dist[o] = d;
        if (t == 0 && lane == 0 && prm.min_dist) prm.min_dist[b] = d;  // dune.py:97-98
      }
    }
    __syncthreads();  // keys / smu / cands are reused by the next item
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}


// ------------------------------------------------------------------------------------------------------
// Ping-pong version: two 128-point tiles in flight per CTA and a dedicated MMA-issuer warp.
//   warps 0-3 (compute): thread r owns point r of tile A (TMEM slot 0) and of tile B (slot 1) and alternates
//     between them: wait D[slot] -> tcgen05.ld -> epilogue (LN/tanh/ReLU/split) -> tcgen05.st A[slot] + bias ->
//     arrive on a_ready[slot].  While it works on one slot the tensor core computes the other, so the waits are
//     (almost) always already satisfied and no block-wide barrier is needed inside the tile loop.
//   warp 4 (issuer): wait a_ready[slot] (4 arrivals, one per compute warp) -> 6 tcgen05.mma -> tcgen05.commit -> d_ready[slot].
// TMEM per CTA: 128 columns = 2 slots x {D [0,32) | A_hi [32,48) | A_lo [48,64)}.
__host__ __device__ inline size_t dune_tc2_smem_bytes(int N, int E, int M) { return dune_tc_smem_bytes(N, E, M); }

__global__ void __launch_bounds__(160, 4) dune_tc2_kernel(const DuneParams prm, const unsigned char* __restrict__ image) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  using I = TcImage;
  unsigned char* simg = smem_dyn;
  const float* fl = reinterpret_cast<const float*>(simg + I::kFloatOff);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(simg + I::kBytes);
  float* smu = reinterpret_cast<float*>(simg + I::kBytes + (size_t)prm.N * 8);
  unsigned long long* cands = reinterpret_cast<unsigned long long*>(simg + I::kBytes + (size_t)prm.N * 8 + (((size_t)prm.N * prm.geo.E * 4 + 7) / 8) * 8);
  __shared__ __align__(8) unsigned long long a_ready[2], d_ready[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool issuer = warp == 4;
  for (int i = tid; i < I::kBytes / 16; i += 160) reinterpret_cast<uint4*>(simg)[i] = reinterpret_cast<const uint4*>(image)[i];
  if (tid == 0) {
    for (int sl = 0; sl < 2; ++sl) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(tc::smem_u32(&a_ready[sl])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(&d_ready[sl])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tc::smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t simg_u = tc::smem_u32(simg);
  uint32_t ph[2] = {0u, 0u};  // compute warps: parity of d_ready[slot]; issuer: parity of a_ready[slot]

  const int T1 = prm.T + 1, N = prm.N, M = prm.M, E = prm.geo.E;
  const int items = prm.B * T1;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / T1, t = item - b * T1;
    if (prm.active && prm.active[b] == 0) continue;  // uniform per CTA
    int n = prm.num_points ? prm.num_points[b] : N;
    n = n < 0 ? 0 : (n > N ? N : n);
    const int cnt = n < M ? n : M;
    if (t == 0 && tid == 0) {
      prm.sel_count[b] = cnt;
      if (n == 0 && prm.min_dist) prm.min_dist[b] = __int_as_float(0x7f800000);
    }
    if (n == 0) continue;
    const int pairs = (n + 255) >> 8;  // two tiles of 128 points per pass

    if (issuer) {
      // ---- MMA issuer warp ---------------------------------------------------------------------------
#pragma unroll 1
      for (int pr = 0; pr < pairs; ++pr)
#pragma unroll 1
        for (int st = 0; st < 5; ++st)
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            tc::mbar_wait(tc::smem_u32(&a_ready[sl]), ph[sl]);
            ph[sl] ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
              const uint32_t tD = tbase + 64 * sl, tAhi = tD + 32, tAlo = tD + 48;
              const bool is_head = st == 4;
              const uint32_t w = simg_u + (is_head ? I::kHeadOff : st * I::kHiddenStride);
              const uint32_t lbo = is_head ? 256u : 512u, ks = is_head ? 512u : 1024u, lo_off = is_head ? 1024u : 2048u;
              const uint32_t idesc = is_head ? tc::kIdescN16 : tc::kIdescN32;
#pragma unroll
              for (int s = 0; s < 2; ++s) {
                const uint64_t bhi = tc::b_desc(w + s * ks, lbo), blo = tc::b_desc(w + lo_off + s * ks, lbo);
                tc::mma_f16(tD, tAlo + 8 * s, bhi, idesc, 1u);
                tc::mma_f16(tD, tAhi + 8 * s, blo, idesc, 1u);
                tc::mma_f16(tD, tAhi + 8 * s, bhi, idesc, 1u);
              }
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc::smem_u32(&d_ready[sl])) : "memory");
            }
            __syncwarp();
          }
    } else {
      // ---- compute warps -----------------------------------------------------------------------------
      const float* ns = prm.nom_s + (size_t)b * 3 * T1;
      const float sx = ns[t], sy = ns[T1 + t], th = ns[2 * T1 + t];
      const float cs = cosf(th), sn = sinf(th);
      const float* px = prm.points + (size_t)b * 2 * N;
      const float* py = px + N;
      const float* vx = prm.velocities ? prm.velocities + (size_t)b * 2 * N : nullptr;
      const float* vy = vx ? vx + N : nullptr;
      const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);

      // hand the activations h of one slot to the tensor core: A := split(h), D := bias, signal the issuer
      auto publish = [&](const float (&h)[32], int sl, const float* bias, bool head) {
        uint32_t hi[16], lo[16];
        tc::split32(h, hi, lo);
        tc::st16(trow + 64 * sl + 32, hi);
        tc::st16(trow + 64 * sl + 48, lo);
        if (head) tc::st_bias16(trow + 64 * sl, bias); else tc::st_bias32(trow + 64 * sl, bias);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&a_ready[sl])) : "memory");
      };
      auto acquire = [&](int sl) {
        tc::mbar_wait(tc::smem_u32(&d_ready[sl]), ph[sl]);
        ph[sl] ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      };

#pragma unroll 1
      for (int pr = 0; pr < pairs; ++pr) {
        float x0[2], y0[2];
        int pi[2];
        bool valid[2];
        // stage 0 of both slots: layer 0 + LayerNorm/tanh on the FMA / MUFU pipes
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          int i = (pr * 2 + sl) * 128 + tid;
          valid[sl] = i < n;
          i = valid[sl] ? i : n - 1;
          pi[sl] = i;
          float gx = px[i], gy = py[i];
          if (vx) {
            gx = flow(gx, vx[i], prm.dt, t);
            gy = flow(gy, vy[i], prm.dt, t);
          }
          const float dx = gx - sx, dy = gy - sy;
          x0[sl] = fmaf(cs, dx, sn * dy);
          y0[sl] = fmaf(cs, dy, -(sn * dx));
          float h[32];
#pragma unroll
          for (int j2 = 0; j2 < 16; ++j2) {
            const float4 w = *reinterpret_cast<const float4*>(fl + I::W0 + 4 * j2);
            const float2 bb = *reinterpret_cast<const float2*>(fl + I::B0 + 2 * j2);
            h[2 * j2] = fmaf(w.y, y0[sl], fmaf(w.x, x0[sl], bb.x));
            h[2 * j2 + 1] = fmaf(w.w, y0[sl], fmaf(w.z, x0[sl], bb.y));
          }
          tc::ln_tanh32(h, fl + I::G1, fl + I::BE1);
          publish(h, sl, fl + I::BH, false);
        }
        // stages 1..4: epilogue of dense layer st-1, input of dense layer st (st == 4: the head)
#pragma unroll 1
        for (int st = 1; st < 5; ++st) {
#pragma unroll
          for (int sl = 0; sl < 2; ++sl) {
            float h[32];
            acquire(sl);
            tc::ld32(trow + 64 * sl, h);
            if (st == 1 || st == 3) {
#pragma unroll
              for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
            } else if (st == 2) {
              tc::ln_tanh32(h, fl + I::G6, fl + I::BE6);
            } else {
              tc::ln_tanh32(h, fl + I::G11, fl + I::BE11);
            }
            publish(h, sl, st == 4 ? fl + I::BHEAD : fl + I::BH + 32 * st, st == 4);
          }
        }
        // head epilogue: mu = relu(.), distance, key
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          float mu[8];
          acquire(sl);
          tc::ld8(trow + 64 * sl, mu);
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < kMaxEdges; ++e) {
            if (e < E) {
              mu[e] = fmaxf(mu[e], 0.f);
              const float ge = fmaf(prm.geo.G[e][1], y0[sl], prm.geo.G[e][0] * x0[sl]) - prm.geo.h[e];
              d = fmaf(mu[e], ge, d);
              if (valid[sl]) smu[pi[sl] * E + e] = mu[e];
            }
          }
          if (valid[sl]) keys[pi[sl]] = ((unsigned long long)orderable(d) << 32) | (unsigned)pi[sl];
        }
      }
    }
    __syncthreads();

    // ---- top-M (compute warps; the issuer warp only joins the barriers) ---------------------------------
    unsigned long long mine = ~0ull;
    if (!issuer) {
      unsigned long long* cand = cands + warp * M;
      for (int m = 0; m < cnt; ++m) {
        unsigned bd = 0xFFFFFFFFu, bi = 0xFFFFFFFFu;
        for (int i = warp * 32 + lane; i < n; i += 128) {
          const uint2 k = *reinterpret_cast<const uint2*>(keys + i);
          if (k.y < bd) { bd = k.y; bi = k.x; }
        }
        const unsigned md = __reduce_min_sync(0xffffffffu, bd);
        const unsigned mi = __reduce_min_sync(0xffffffffu, bd == md ? bi : 0xFFFFFFFFu);
        if (md != 0xFFFFFFFFu && bd == md && bi == mi) keys[mi] = ~0ull;
        if (lane == 0) cand[m] = md == 0xFFFFFFFFu ? ~0ull : (((unsigned long long)md << 32) | mi);
        __syncwarp();
      }
    }
    __syncthreads();
    if (warp == 0) {
      const float* ns = prm.nom_s + (size_t)b * 3 * T1;
      const float th = ns[2 * T1 + t];
      const float cs = cosf(th), sn = sinf(th);
      const float* px = prm.points + (size_t)b * 2 * N;
      const float* py = px + N;
      const float* vx = prm.velocities ? prm.velocities + (size_t)b * 2 * N : nullptr;
      const float* vy = vx ? vx + N : nullptr;
      const int total = 4 * cnt;
      for (int m = 0; m < cnt; ++m) {
        unsigned bd = 0xFFFFFFFFu, bi = 0xFFFFFFFFu;
        int bpos = -1;
        for (int c = lane; c < total; c += 32) {
          const int w = c / cnt, r = c - w * cnt;
          const uint2 k = *reinterpret_cast<const uint2*>(cands + w * M + r);
          if (k.y < bd || (k.y == bd && k.x < bi)) { bd = k.y; bi = k.x; bpos = w * M + r; }
        }
        const unsigned md = __reduce_min_sync(0xffffffffu, bd);
        const unsigned mi = __reduce_min_sync(0xffffffffu, bd == md ? bi : 0xFFFFFFFFu);
        if (bpos >= 0 && bd == md && bi == mi) cands[bpos] = ~0ull;
        if (lane == m) mine = ((unsigned long long)md << 32) | mi;
        __syncwarp();
      }
      if (lane < cnt) {
        unsigned idx = (unsigned)(mine & 0xffffffffull);
        if (idx >= (unsigned)n) idx = 0;
        uint32_t u = (uint32_t)(mine >> 32);
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        const float d = __uint_as_float(u);
        float gx = px[idx], gy = py[idx];
        if (vx) {
          gx = flow(gx, vx[idx], prm.dt, t);
          gy = flow(gy, vy[idx], prm.dt, t);
        }
        const size_t o = ((size_t)b * T1 + t) * M + lane;
        float lx = 0.f, ly = 0.f;
#pragma unroll
        for (int e = 0; e < kMaxEdges; ++e) {
          if (e < E) {
            const float m_e = smu[idx * E + e];
            lx = fmaf(fmaf(sn, prm.geo.G[e][1], -cs * prm.geo.G[e][0]), m_e, lx);
            ly = fmaf(fmaf(-cs, prm.geo.G[e][1], -sn * prm.geo.G[e][0]), m_e, ly);
            prm.sel_mu[o * E + e] = m_e;
          }
        }
        prm.sel_lam[o * 2 + 0] = lx; prm.sel_lam[o * 2 + 1] = ly;
        prm.sel_pts[o * 2 + 0] = gx; prm.sel_pts[o * 2 + 1] = gy;
        prm.sel_dist[o] = d;
        if (t == 0 && lane == 0 && prm.min_dist) prm.min_dist[b] = d;
      }
    }
    __syncthreads();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

}  // namespace nb
