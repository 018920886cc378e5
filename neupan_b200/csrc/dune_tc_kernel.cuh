// DUNE kernels, tcgen05 versions (Blackwell 5th-gen tensor cores, accumulators and A operand in TMEM).
//
// Same contract as dune_mma_kernel.cuh / dune_kernel.cuh.  Mapping: CTA = 128 threads = 128 TMEM lanes;
// thread r owns point r of a 128-point tile for the whole network, so LayerNorm statistics, tanh, ReLU
// and the distance are thread-local (no shuffles, no fragment bookkeeping).  Per dense layer:
//   registers --split x = hi + lo (fp16)--> tcgen05.st (A operand, TMEM)              [all threads]
//   D[128 x 32] = ONES.BIASB + A_lo.B_hi + A_hi.B_lo + A_hi.B_hi                       [one elected lane, tcgen05.mma,
//                                                      B (and ONES) from shared memory via UMMA descriptors]
//   tcgen05.commit -> mbarrier -> tcgen05.ld (32 fp32 columns = the thread's row)      [all threads]
// Layouts / descriptors were verified in isolation with tools/tc05_probe.cu.
//
// dune_tcp_kernel (default): two 128-point tiles ("slots") per CTA, 128 TMEM columns = 2 x {D [0,32) | A_hi [32,48) |
//   A_lo [48,64)}; the epilogue of one slot overlaps the MMAs of the other; packed FP32 / FHFMA epilogue math.
// dune_tc_kernel (NB_DUNE_TC=1): the first version -- one slot (64 columns), bias row written into D by tcgen05.st,
//   scalar epilogue math, every layer waits for its MMAs.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "dune_kernel.cuh"  // DuneParams, orderable(), flow()

namespace nb {

// Weight image built on the host (dune_tc.cu): canonical K-major / no-swizzle UMMA operand layout
// (core matrix = 8 rows x 16 B; SBO = 128 B between 8-row groups; LBO between the two 8-half K groups).
struct TcImage {
  static constexpr int kLayerStride = 4096;  // per dense layer (4 hidden + head, all N = 32): W_hi 2048 | W_lo 2048
  static constexpr int kLayers = 5;          // the head's rows E..31 are zero
  static constexpr int kFloatOff = kLayers * kLayerStride;  // 20480
  // float section (offsets in floats): W0 (32x2, row-major), b0, LayerNorm gain/offset pre-multiplied by 2*log2(e),
  // the biases of the five dense layers (already including the folded tanh map; head zero padded to 32), and W0
  // again as two columns (W0X[j] = W0[j][0], W0Y[j] = W0[j][1]) for the packed-FP32 layer 0
  static constexpr int W0 = 0, B0 = 64, G1 = 96, BE1 = 128, G6 = 160, BE6 = 192, G11 = 224, BE11 = 256, BH = 288, W0X = 448, W0Y = 480,
                       kFloats = 512;
  // bias by tensor core: D = ONES[128 x 16] . BIASB[layer][32 x 16] with ONES(m, 0..2) = 1 and BIASB(n, 0..2) = the three
  // fp16 pieces of the bias (hi, lo, lo2: exact to fp32) -- the first MMA of a layer, replaces 8 LDS.128 + a 32-column
  // tcgen05.st per thread and layer.  Both are K-major / no-swizzle core-matrix images like the weights.
  static constexpr int kOnesOff = kFloatOff + kFloats * 4;   // 22528, 4096 B  (LBO 2048)
  static constexpr int kBiasBOff = kOnesOff + 4096;          // 26624, 5 x 1024 B (LBO 512)
  static constexpr int kBytes = kBiasBOff + kLayers * 1024;  // 31744
};

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr uint32_t kIdescN32 = (1u << 4) | (4u << 17) | (8u << 24);  // D f32, A/B f16 K-major, N = 32, M = 128

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

// A and B from shared memory (both through descriptors); used for the bias product that initialises D
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
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

// ---- packed FP32 (FFMA2 / FADD2 / FMUL2: one issue slot, two lanes) and mixed f16*f16+f32 (FHFMA) -----------------
using f2 = unsigned long long;  // two floats in an aligned register pair (low word = even feature)
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 pku(uint32_t lo, uint32_t hi) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// (a0, a1) -> hi = fp16x2(a), lo = fp16x2(a - hi): F2FP, 2 x FHFMA (a - hi*1 in one instruction), F2FP
__device__ __forceinline__ void split_pair(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  const __half2 hh = __floats2half2_rn(a0, a1);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  unsigned short h0, h1;
  asm("mov.b32 {%0,%1}, %2;" : "=h"(h0), "=h"(h1) : "r"(hi));
  const unsigned short m1 = 0xBC00;  // -1.0h
  float l0, l1;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l0) : "h"(h0), "h"(m1), "f"(a0));
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l1) : "h"(h1), "h"(m1), "f"(a1));
  const __half2 ll = __floats2half2_rn(l0, l1);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}

__device__ __forceinline__ void ld32p(uint32_t taddr, f2 (&hp)[16]) {
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
  for (int c = 0; c < 16; ++c) hp[c] = pku(d[2 * c], d[2 * c + 1]);
}

// LayerNorm (eps 1e-5, biased variance) + r = 1/(exp2(.)+1) + fp16 split, two features per instruction where the ISA
// allows it (squares, scale/offset, +1); MUFU.EX2 / MUFU.RCP stay scalar.  The inputs arrive CENTRED: the layer that
// produces them has W - colmean(W), b - mean(b) (host image), so mean(h) = 0 up to rounding and only the variance is left.
// kFast: one MUFU.RCP serves four features: with d_i = exp2(a_i) + 1,  t = 1/(d0 d1 d2 d3),  1/d0 = t d1 d2 d3, ...  (three
// packed and three scalar multiplies instead of three MUFUs; the host enables it only if the checkpoint's LayerNorm
// gains/offsets bound every a_i by 30, so that the product of four stays below 2^124)
template <bool kFast>
__device__ __forceinline__ void ln_tanh_split(f2 (&hp)[16], const float* __restrict__ g, const float* __restrict__ be, uint32_t (&hi)[16],
                                              uint32_t (&lo)[16]) {
  f2 qa = 0ull, qb = 0ull, qc = 0ull, qd = 0ull;
#pragma unroll
  for (int c = 0; c < 16; c += 4) {
    qa = fma2(hp[c], hp[c], qa);
    qb = fma2(hp[c + 1], hp[c + 1], qb);
    qc = fma2(hp[c + 2], hp[c + 2], qc);
    qd = fma2(hp[c + 3], hp[c + 3], qd);
  }
  float q0, q1;
  upk(add2(add2(qa, qb), add2(qc, qd)), q0, q1);
  float r;  // the argument is >= 1e-5: no denormal guard needed
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(q0 + q1, 1.0f / 32, 1e-5f)));
  const f2 r2 = pk(r, r), one2 = pk(1.0f, 1.0f);
#pragma unroll
  for (int c = 0; c < 16; c += 2) {
    const ulonglong2 gg = *reinterpret_cast<const ulonglong2*>(g + 2 * c);
    const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(be + 2 * c);
    float a0, a1, a2, a3, e0, e1, e2, e3;
    upk(fma2(mul2(hp[c], r2), gg.x, bb.x), a0, a1);
    upk(fma2(mul2(hp[c + 1], r2), gg.y, bb.y), a2, a3);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(a2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(a3));
    const f2 da = add2(pk(e0, e1), one2), db = add2(pk(e2, e3), one2);  // (d0, d1), (d2, d3)
    float r0, r1, r2s, r3;
    if (kFast) {
      float p0, p1, t;
      upk(mul2(da, db), p0, p1);  // (d0 d2, d1 d3)
      const float pp = p0 * p1;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(pp));
      const f2 u = pk(t * p1, t * p0);  // (1/(d0 d2), 1/(d1 d3))
      upk(mul2(u, db), r0, r1);         // 1/d0, 1/d1
      upk(mul2(u, da), r2s, r3);        // 1/d2, 1/d3
    } else {
      float d0, d1, d2, d3;
      upk(da, d0, d1);
      upk(db, d2, d3);
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r2s) : "f"(d2));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r3) : "f"(d3));
    }
    split_pair(r0, r1, hi[c], lo[c]);
    split_pair(r2s, r3, hi[c + 1], lo[c + 1]);
  }
}

// relu + split without a max: hi = fp16x2(relu(a)) rounded TOWARDS ZERO (cvt.rz.relu), so that a - hi is >= 0 for a > 0
// and equals a < 0 for a <= 0 (hi = 0); lo = fp16x2(relu(a - hi)) then drops exactly the negative inputs
__device__ __forceinline__ void relu_split(const f2 (&hp)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
  const unsigned short m1 = 0xBC00;  // -1.0h
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float a0, a1, l0, l1;
    upk(hp[c], a0, a1);
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[c]) : "f"(a1), "f"(a0));  // d = {hi half: first source, lo half: second}
    unsigned short h0, h1;
    asm("mov.b32 {%0,%1}, %2;" : "=h"(h0), "=h"(h1) : "r"(hi[c]));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l0) : "h"(h0), "h"(m1), "f"(a0));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l1) : "h"(h1), "h"(m1), "f"(a1));
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo[c]) : "f"(l1), "f"(l0));
  }
}

}  // namespace tc

__host__ __device__ inline size_t dune_tc_smem_bytes(int N, int E, int M) {
  return TcImage::kBytes + (size_t)N * 8 + (((size_t)N * E * 4 + 7) / 8) * 8 + (size_t)4 * M * 8 + 64;
}

namespace tc {

// point i of env b at step t in the robot frame of the nominal state: p0 = R^T (p_t - trans)   (pan.py:210)
struct ItemFrame {
  float sx, sy, cs, sn, dt;
  const float *px, *py, *vx, *vy;
  int t;
  __device__ __forceinline__ void world(int i, float& gx, float& gy) const {
    gx = px[i]; gy = py[i];
    if (vx) {
      gx = flow(gx, vx[i], dt, t);
      gy = flow(gy, vy[i], dt, t);
    }
  }
  __device__ __forceinline__ void local(int i, float& x0, float& y0) const {
    float gx, gy;
    world(i, gx, gy);
    const float dx = gx - sx, dy = gy - sy;
    x0 = fmaf(cs, dx, sn * dy);
    y0 = fmaf(cs, dy, -(sn * dx));
  }
};

__device__ __forceinline__ ItemFrame item_frame(const DuneParams& prm, int b, int t) {
  const int T1 = prm.T + 1;
  const float* ns = prm.nom_s + (size_t)b * 3 * T1;
  ItemFrame f;
  f.sx = ns[t]; f.sy = ns[T1 + t];
  const float th = ns[2 * T1 + t];
  f.cs = cosf(th); f.sn = sinf(th);
  f.dt = prm.dt; f.t = t;
  f.px = prm.points + (size_t)b * 2 * prm.N;
  f.py = f.px + prm.N;
  f.vx = prm.velocities ? prm.velocities + (size_t)b * 2 * prm.N : nullptr;
  f.vy = f.vx ? f.vx + prm.N : nullptr;
  return f;
}

// mu = relu(head), dist = mu^T (G p0 - h)   (dune.py:119-122); publishes mu and the sortable key of the point
__device__ __forceinline__ uint32_t finish_point(const DuneParams& prm, float (&mu)[8], float x0, float y0, int i, bool valid, int E, float* smu,
                                                 unsigned long long* keys) {
  float d = 0.f;
#pragma unroll
  for (int e = 0; e < kMaxEdges; ++e) {
    if (e < E) {
      mu[e] = fmaxf(mu[e], 0.f);
      const float ge = fmaf(prm.geo.G[e][1], y0, prm.geo.G[e][0] * x0) - prm.geo.h[e];
      d = fmaf(mu[e], ge, d);
      if (valid) smu[i * E + e] = mu[e];
    }
  }
  const uint32_t key = orderable(d);
  if (valid && keys) keys[i] = ((unsigned long long)key << 32) | (unsigned)i;
  return valid ? key : 0xFFFFFFFFu;
}

// top-M (ascending, ties -> lower index) of the item's keys and the output rows; called by the 4 compute warps of a CTA
// after a block barrier that made keys / smu visible.  Contains one block barrier.
__device__ __forceinline__ void select_and_write(const DuneParams& prm, const ItemFrame& fr, int b, int t, int n, int cnt, int warp, int lane,
                                                 unsigned long long* keys, const float* smu, unsigned long long* cands) {
  const int M = prm.M, E = prm.geo.E, T1 = prm.T + 1;
  unsigned long long mine = ~0ull;
  {
    unsigned long long* cand = cands + warp * M;
    for (int m = 0; m < cnt; ++m) {  // per-warp REDUX rounds over the keys of the rows this warp computed
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
  if (warp == 0) {  // merge the 4 candidate lists
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
    if (lane < cnt) {  // lane m writes the m-th closest point
      unsigned idx = (unsigned)(mine & 0xffffffffull);
      if (idx >= (unsigned)n) idx = 0;
      uint32_t u = (uint32_t)(mine >> 32);
      u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
      const float d = __uint_as_float(u);
      float gx, gy;
      fr.world(idx, gx, gy);
      const size_t o = ((size_t)b * T1 + t) * M + lane;
      float lx = 0.f, ly = 0.f;
#pragma unroll
      for (int e = 0; e < kMaxEdges; ++e) {  // lam = ((-R) G^T) mu   (dune.py:89); constant indices keep geo in the constant bank
        if (e < E) {
          const float m_e = smu[idx * E + e];
          lx = fmaf(fmaf(fr.sn, prm.geo.G[e][1], -fr.cs * prm.geo.G[e][0]), m_e, lx);
          ly = fmaf(fmaf(-fr.cs, prm.geo.G[e][1], -fr.sn * prm.geo.G[e][0]), m_e, ly);
          prm.sel_mu[o * E + e] = m_e;
        }
      }
      prm.sel_lam[o * 2 + 0] = lx; prm.sel_lam[o * 2 + 1] = ly;
      prm.sel_pts[o * 2 + 0] = gx; prm.sel_pts[o * 2 + 1] = gy;
      prm.sel_dist[o] = d;
      if (t == 0 && lane == 0 && prm.min_dist) prm.min_dist[b] = d;  // dune.py:97-98
    }
  }
}

// the six MMAs of one dense layer (3-pass fp16 split, K = 32 as two K = 16 steps) + commit to an mbarrier; one thread
__device__ __forceinline__ void issue_layer(uint32_t tD, uint32_t wsmem, uint32_t bar) {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const uint64_t bhi = b_desc(wsmem + s * 1024, 512), blo = b_desc(wsmem + 2048 + s * 1024, 512);
    mma_f16(tD, tD + 48 + 8 * s, bhi, kIdescN32, 1u);  // A_lo . B_hi
    mma_f16(tD, tD + 32 + 8 * s, blo, kIdescN32, 1u);  // A_hi . B_lo
    mma_f16(tD, tD + 32 + 8 * s, bhi, kIdescN32, 1u);  // A_hi . B_hi
  }
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// same, but D is initialised by the bias product ONES . BIASB[layer] instead of a tcgen05.st of the bias row.
// The descriptors are linear in the operand address (all operands are 16-byte multiples inside one 1024-aligned image and
// shared memory addresses fit the 14-bit field), so they are formed by adding constants to two precomputed low words:
//   dw = descriptor low word of the image base with LBO 512 (weights, BIASB), done = same for ONES (LBO 2048)
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version bit 46
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }
__device__ __forceinline__ void issue_layer_bias(uint32_t tD, uint32_t dw, uint32_t done, int layer, uint32_t bar) {
  using I = TcImage;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  mma_f16_ss(tD, mk_desc(done), mk_desc(dw + (I::kBiasBOff >> 4) + layer * (1024 >> 4)), kIdescN32, 0u);
  const uint32_t w = dw + layer * (I::kLayerStride >> 4);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const uint64_t bhi = mk_desc(w + s * (1024 >> 4)), blo = mk_desc(w + (2048 >> 4) + s * (1024 >> 4));
    mma_f16(tD, tD + 48 + 8 * s, bhi, kIdescN32, 1u);  // A_lo . B_hi
    mma_f16(tD, tD + 32 + 8 * s, blo, kIdescN32, 1u);  // A_hi . B_lo
    mma_f16(tD, tD + 32 + 8 * s, bhi, kIdescN32, 1u);  // A_hi . B_hi
  }
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Selection for n <= 512 with the thread's own (<= 4) distance keys in registers (thread tid owns points tid + 128 j):
// per-warp REDUX rounds produce each warp's cnt smallest in ascending order, then every candidate computes its rank among
// the 4 x cnt candidates (all four warps in parallel, no serial merge) and the lanes of rank < cnt write the output rows.
__device__ __forceinline__ void select_and_write_reg(const DuneParams& prm, const ItemFrame& fr, int b, int t, int n, int cnt, int tid, uint32_t k0,
                                                     uint32_t k1, uint32_t k2, uint32_t k3, const float* smu, unsigned long long* cands) {
  const int M = prm.M, E = prm.geo.E, T1 = prm.T + 1, warp = tid >> 5, lane = tid & 31;
  unsigned long long* cand = cands + warp * M;
  for (int m = 0; m < cnt; ++m) {
    const uint32_t bd = min(min(k0, k1), min(k2, k3));
    const uint32_t md = __reduce_min_sync(0xffffffffu, bd);
    const int j = k0 == md ? 0 : (k1 == md ? 1 : (k2 == md ? 2 : 3));
    const uint32_t bi = bd == md ? (uint32_t)(tid + 128 * j) : 0xFFFFFFFFu;
    const uint32_t mi = __reduce_min_sync(0xffffffffu, bi);
    if (md != 0xFFFFFFFFu && bi == mi) {  // the owner retires the key
      if (j == 0) k0 = 0xFFFFFFFFu;
      else if (j == 1) k1 = 0xFFFFFFFFu;
      else if (j == 2) k2 = 0xFFFFFFFFu;
      else k3 = 0xFFFFFFFFu;
    }
    if (lane == 0) cand[m] = md == 0xFFFFFFFFu ? ~0ull : (((unsigned long long)md << 32) | mi);
  }
  __syncthreads();
  if (lane < cnt) {
    const unsigned long long mine = cand[lane];
    int rank = lane;  // candidates of the own warp are sorted and distinct
    for (int w = 0; w < 4; ++w) {
      if (w == warp) continue;
      for (int r = 0; r < cnt; ++r) rank += cands[w * M + r] < mine ? 1 : 0;  // keys carry the point index: no ties
    }
    if (mine != ~0ull && rank < cnt) {
      const unsigned idx = (unsigned)(mine & 0xffffffffull);
      uint32_t u = (uint32_t)(mine >> 32);
      u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
      const float d = __uint_as_float(u);
      float gx, gy;
      fr.world(idx, gx, gy);
      const size_t o = ((size_t)b * T1 + t) * M + rank;
      float lx = 0.f, ly = 0.f;
#pragma unroll
      for (int e = 0; e < kMaxEdges; ++e) {  // lam = ((-R) G^T) mu   (dune.py:89)
        if (e < E) {
          const float m_e = smu[idx * E + e];
          lx = fmaf(fmaf(fr.sn, prm.geo.G[e][1], -fr.cs * prm.geo.G[e][0]), m_e, lx);
          ly = fmaf(fmaf(-fr.cs, prm.geo.G[e][1], -fr.sn * prm.geo.G[e][0]), m_e, ly);
          prm.sel_mu[o * E + e] = m_e;
        }
      }
      prm.sel_lam[o * 2 + 0] = lx; prm.sel_lam[o * 2 + 1] = ly;
      prm.sel_pts[o * 2 + 0] = gx; prm.sel_pts[o * 2 + 1] = gy;
      prm.sel_dist[o] = d;
      if (t == 0 && rank == 0 && prm.min_dist) prm.min_dist[b] = d;  // dune.py:97-98
    }
  }
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------------
// Default kernel: two 128-point tiles ("slots") per CTA, interleaved by the same 128 threads.  Thread r owns point r of
// both tiles.  While the threads run the epilogue of one slot (tcgen05.ld -> LayerNorm/tanh or ReLU -> fp16 split ->
// tcgen05.st), the tensor core works on the other slot, so the mbarrier wait that follows is already satisfied: no
// polling, no MMA latency on the critical path, no dedicated issuer warp.
//   TMEM per CTA: 128 columns = 2 slots x {D [0,32) | A_hi [32,48) | A_lo [48,64)};  4 CTAs per SM.
//   Element-wise math is packed two features per instruction (FFMA2/FADD2/FMUL2) and the fp16 split uses FHFMA.
// kSync = 0: a block barrier between the operand stores of a slot and its MMAs;  kSync = 1: the four warps arrive on an
// mbarrier instead and only warp 0 (whose lane 0 issues the MMAs) waits for it, so warps 1-3 run ahead into the other slot.
template <int kSync, bool kFast>
__global__ void __launch_bounds__(128, 4) dune_tcp_kernel(const DuneParams prm, const unsigned char* __restrict__ image) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];  // the attribute aligns the dynamic segment (UMMA operands need 128 B)
  using I = TcImage;
  unsigned char* simg = smem_dyn;
  const float* fl = reinterpret_cast<const float*>(simg + I::kFloatOff);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(simg + I::kBytes);
  float* smu = reinterpret_cast<float*>(simg + I::kBytes + (size_t)prm.N * 8);
  unsigned long long* cands = reinterpret_cast<unsigned long long*>(simg + I::kBytes + (size_t)prm.N * 8 + (((size_t)prm.N * prm.geo.E * 4 + 7) / 8) * 8);
  __shared__ __align__(8) unsigned long long mbar[4];  // [0,1]: D of slot 0/1 ready (tcgen05.commit);  [2,3]: operands of slot 0/1 stored
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < I::kBytes / 16; i += 128) reinterpret_cast<uint4*>(simg)[i] = reinterpret_cast<const uint4*>(image)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(&mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(&mbar[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(tc::smem_u32(&mbar[2])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(tc::smem_u32(&mbar[3])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // operand image -> visible to the tensor core (async proxy)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(tc::smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes
  const uint32_t simg_u = tc::smem_u32(simg);
  const uint32_t bar0 = tc::smem_u32(&mbar[0]);
  const uint32_t desc_w = tc::desc_lo(simg_u, 512), desc_ones = tc::desc_lo(simg_u + I::kOnesOff, 2048);
  uint32_t phases = 0;  // bit sl: parity of the next completion of mbar[sl];  bit 2 + sl: same for mbar[2 + sl]

  // activations (already split) of `slot` -> TMEM, then one thread starts the layer's MMAs (bias product first)
  auto publish = [&](const uint32_t (&hi)[16], const uint32_t (&lo)[16], int slot, int layer) {
    const uint32_t tS = trow + 64 * slot;
    tc::st16(tS + 32, hi);
    tc::st16(tS + 48, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // The issuing warp rotates: warp w of every CTA lives on scheduler w of its SM, so a fixed issuer would put the ~100
    // issue instructions (and, with kSync = 1, the waiting) of all resident CTAs on one of the four schedulers.
    const int issuer = (layer + slot + (int)blockIdx.x) & 3;
    if (kSync == 0) {
      __syncthreads();
    } else {
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 16 + 8 * slot) : "memory");
      if (warp == issuer) tc::mbar_wait(bar0 + 16 + 8 * slot, (phases >> (2 + slot)) & 1u);
      phases ^= 4u << slot;  // every warp tracks the parity, only the issuer waits
    }
    if (warp == issuer) {  // warp-uniform branch; elect.sync picks the issuing lane (no divergent-branch waterfall around UTCHMMA)
      uint32_t elected;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(elected));
      if (elected) tc::issue_layer_bias(tbase + 64 * slot, desc_w, desc_ones, layer, bar0 + 8 * slot);
    }
  };
  auto acquire = [&](int slot) {
    tc::mbar_wait(bar0 + 8 * slot, (phases >> slot) & 1u);
    phases ^= 1u << slot;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };

  const int T1 = prm.T + 1, N = prm.N, M = prm.M, E = prm.geo.E;
  const int items = prm.B * T1;
  const int n_work = prm.only_flagged ? *prm.flag_count : items;  // screening mode: only the items the screen could not narrow down
  for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
    const int item = prm.only_flagged ? prm.flag_list[wi] : wi;
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
    const tc::ItemFrame fr = tc::item_frame(prm, b, t);
    const bool reg_keys = n <= 512;  // the thread's own keys stay in registers (uniform per CTA)
    uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu, k3 = 0xFFFFFFFFu;

    // ---- phase 1: the network, 256 points (two slots) per pass ---------------------------------------------------------
#pragma unroll 1
    for (int base = 0; base < n; base += 256) {
      const int nslots = base + 128 < n ? 2 : 1;  // a pass whose second tile would be empty runs one slot
#pragma unroll 1
      for (int sl = 0; sl < nslots; ++sl) {  // stage 0: layer 0 (2 -> 32) on the FMA pipe + LayerNorm/tanh
        int i = base + sl * 128 + tid;
        i = i < n ? i : n - 1;
        float x0, y0;
        fr.local(i, x0, y0);
        const tc::f2 x2 = tc::pk(x0, x0), y2 = tc::pk(y0, y0);
        tc::f2 hp[16];
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const ulonglong2 wx = *reinterpret_cast<const ulonglong2*>(fl + I::W0X + 2 * c);
          const ulonglong2 wy = *reinterpret_cast<const ulonglong2*>(fl + I::W0Y + 2 * c);
          const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(fl + I::B0 + 2 * c);
          hp[c] = tc::fma2(wy.x, y2, tc::fma2(wx.x, x2, bb.x));
          hp[c + 1] = tc::fma2(wy.y, y2, tc::fma2(wx.y, x2, bb.y));
        }
        uint32_t hi[16], lo[16];
        tc::ln_tanh_split<kFast>(hp, fl + I::G1, fl + I::BE1, hi, lo);
        publish(hi, lo, sl, 0);
      }
#pragma unroll 1
      for (int st = 1; st < 5; ++st) {  // stage st: epilogue of dense layer st-1, operands of dense layer st (4 = head)
#pragma unroll 1
        for (int sl = 0; sl < nslots; ++sl) {
          tc::f2 hp[16];
          uint32_t hi[16], lo[16];
          acquire(sl);
          tc::ld32p(trow + 64 * sl, hp);
          if (st & 1) tc::relu_split(hp, hi, lo);
          else tc::ln_tanh_split<kFast>(hp, fl + I::G1 + 32 * st, fl + I::BE1 + 32 * st, hi, lo);  // st = 2: G6/BE6, st = 4: G11/BE11
          publish(hi, lo, sl, st);
        }
      }
#pragma unroll 1
      for (int sl = 0; sl < nslots; ++sl) {  // head epilogue
        float mu[8];
        acquire(sl);
        tc::ld8(trow + 64 * sl, mu);
        int i = base + sl * 128 + tid;
        const bool valid = i < n;
        i = valid ? i : n - 1;
        float x0, y0;
        fr.local(i, x0, y0);
        const uint32_t key = tc::finish_point(prm, mu, x0, y0, i, valid, E, smu, reg_keys ? nullptr : keys);
        const int j = (base >> 7) + sl;
        if (j == 0) k0 = key;
        else if (j == 1) k1 = key;
        else if (j == 2) k2 = key;
        else if (j == 3) k3 = key;
      }
    }
    __syncthreads();
    // ---- phases 2 / 3: top-M and the output rows -----------------------------------------------------------------------
    if (reg_keys) tc::select_and_write_reg(prm, fr, b, t, n, cnt, tid, k0, k1, k2, k3, smu, cands);
    else tc::select_and_write(prm, fr, b, t, n, cnt, warp, lane, keys, smu, cands);
    __syncthreads();  // smu / cands (and keys) are reused by the next item
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

// ------------------------------------------------------------------------------------------------------
// First tcgen05 version, kept for A/B measurements (NB_DUNE_TC=1): one 128-point tile per CTA at a time, 64 TMEM columns,
// 5 CTAs per SM, scalar FP32 element-wise math; every layer exposes the MMA latency behind an mbarrier poll.
__global__ void __launch_bounds__(128, 5) dune_tc_kernel(const DuneParams prm, const unsigned char* __restrict__ image) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  using I = TcImage;
  unsigned char* simg = smem_dyn;
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
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(tc::smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t simg_u = tc::smem_u32(simg);
  const uint32_t bar = tc::smem_u32(&mbar);
  uint32_t phase = 0;

  // one dense layer on the tensor core: h (activations) -> D (pre-activations incl. bias), left in TMEM
  auto dense = [&](const float (&h)[32], int layer) {
    uint32_t hi[16], lo[16];
    tc::split32(h, hi, lo);
    tc::st16(trow + 32, hi);
    tc::st16(trow + 48, lo);
    tc::st_bias32(trow, fl + I::BH + 32 * layer);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) tc::issue_layer(tbase, simg_u + layer * I::kLayerStride, bar);
    tc::mbar_wait(bar, phase);
    phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };

  const int T1 = prm.T + 1, N = prm.N, M = prm.M, E = prm.geo.E;
  const int items = prm.B * T1;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = item / T1, t = item - b * T1;
    if (prm.active && prm.active[b] == 0) continue;
    int n = prm.num_points ? prm.num_points[b] : N;
    n = n < 0 ? 0 : (n > N ? N : n);
    const int cnt = n < M ? n : M;
    if (t == 0 && tid == 0) {
      prm.sel_count[b] = cnt;
      if (n == 0 && prm.min_dist) prm.min_dist[b] = __int_as_float(0x7f800000);
    }
    if (n == 0) continue;
    const tc::ItemFrame fr = tc::item_frame(prm, b, t);

#pragma unroll 1
    for (int tile = 0; tile * 128 < n; ++tile) {
      int i = tile * 128 + tid;
      const bool valid = i < n;
      i = valid ? i : n - 1;
      float x0, y0;
      fr.local(i, x0, y0);
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
      tc::ld32(trow, h);
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
      dense(h, 1);
      tc::ld32(trow, h);
      tc::ln_tanh32(h, fl + I::G6, fl + I::BE6);
      dense(h, 2);
      tc::ld32(trow, h);
#pragma unroll
      for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j], 0.f);
      dense(h, 3);
      tc::ld32(trow, h);
      tc::ln_tanh32(h, fl + I::G11, fl + I::BE11);
      dense(h, 4);
      float mu[8];
      tc::ld8(trow, mu);
      tc::finish_point(prm, mu, x0, y0, i, valid, E, smu, keys);
    }
    __syncthreads();
    tc::select_and_write(prm, fr, b, t, n, cnt, warp, lane, keys, smu, cands);
    __syncthreads();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tbase) : "memory");
}

}  // namespace nb
