// DUNE kernel, tcgen05 version with TWO THREADS PER POINT (8 warps per 128-point tile).
//
// Why: the two-slot kernel of dune_tc_kernel.cuh (one thread per point, 123 registers, 16 warps per SM) is bound by
// dependent-issue latency -- issue slots 60 % busy, no pipe above 36 %, stalls per issue: fixed-latency wait 1.64, TMEM
// scoreboard 0.60, barrier 0.55 (profiles/r01_dune_tc_v4_final_ncu.txt).  Registers and TMEM pin it at 4 warps per scheduler.
// Here every point is shared by two threads that own 16 of the 32 features each: half the per-thread state (<= 80
// registers), twice the warps per tile, same TMEM footprint (128 columns per CTA: 2 slots x {D 32 | A_hi 16 | A_lo 16}).
//
//   thread (row r = tid & 127, half hf = tid >> 7): warps w and w + 4 address the same 32 TMEM lanes (lane quarter w & 3)
//   and split the columns: D columns [16 hf, 16 hf + 16), A_hi columns 32 + [8 hf, 8 hf + 8), A_lo columns 48 + [8 hf, ..).
//   LayerNorm needs the sum of squares over all 32 features: both threads read the other half as well (one extra
//   tcgen05.ld.x16 + 8 FFMA2) and add the two half sums -- the same two numbers, so both get the same bits (a + b = b + a)
//   without any exchange through shared memory.  Layer 0 (2 -> 32, FMA pipe) is evaluated in full by both threads.
//   The head epilogue (mu, distance, sort key) of slot s is done by the threads with hf == s.
// Everything else -- operand image, descriptors, bias product, centred LayerNorm, folded tanh map, packed FP32 math,
// shared-reciprocal tanh, elect.sync issue, rank-merge top-M -- is the contract of dune_tcp_kernel.
#pragma once
#include "dune_tc_kernel.cuh"

namespace nb {
namespace tc {

__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&a)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]),
               "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
               : "memory");
}

__device__ __forceinline__ void ld16p(uint32_t taddr, f2 (&hp)[8]) {
  uint32_t d[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]), "=r"(d[10]),
        "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int c = 0; c < 8; ++c) hp[c] = pku(d[2 * c], d[2 * c + 1]);
}

// sum of squares of 16 features held as 8 packed pairs; fixed association (two chains, then lanes) so that the two threads of a
// point obtain bit-identical half sums
__device__ __forceinline__ float sumsq16(const f2 (&hp)[8]) {
  f2 qa = 0ull, qb = 0ull;
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    qa = fma2(hp[c], hp[c], qa);
    qb = fma2(hp[c + 1], hp[c + 1], qb);
  }
  float q0, q1;
  upk(add2(qa, qb), q0, q1);
  return q0 + q1;
}

// own 16 features: a = h r g + be (g, be pre-scaled by 2 log2 e), rr = 1 / (exp2(a) + 1), fp16 split (see ln_tanh_split)
template <bool kFast>
__device__ __forceinline__ void tanh_split16(const f2 (&hp)[8], float r, const float* __restrict__ g, const float* __restrict__ be, uint32_t (&hi)[8],
                                             uint32_t (&lo)[8]) {
  const f2 r2 = pk(r, r), one2 = pk(1.0f, 1.0f);
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    const ulonglong2 gg = *reinterpret_cast<const ulonglong2*>(g + 2 * c);
    const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(be + 2 * c);
    float a0, a1, a2, a3, e0, e1, e2, e3;
    upk(fma2(mul2(hp[c], r2), gg.x, bb.x), a0, a1);
    upk(fma2(mul2(hp[c + 1], r2), gg.y, bb.y), a2, a3);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(a2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(a3));
    const f2 da = add2(pk(e0, e1), one2), db = add2(pk(e2, e3), one2);
    float r0, r1, r2s, r3;
    if (kFast) {
      float p0, p1, t;
      upk(mul2(da, db), p0, p1);
      const float pp = p0 * p1;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(pp));
      const f2 u = pk(t * p1, t * p0);
      upk(mul2(u, db), r0, r1);
      upk(mul2(u, da), r2s, r3);
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

__device__ __forceinline__ void relu_split16(const f2 (&hp)[8], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
  const unsigned short m1 = 0xBC00;  // -1.0h
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float a0, a1, l0, l1;
    upk(hp[c], a0, a1);
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi[c]) : "f"(a1), "f"(a0));
    unsigned short h0, h1;
    asm("mov.b32 {%0,%1}, %2;" : "=h"(h0), "=h"(h1) : "r"(hi[c]));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l0) : "h"(h0), "h"(m1), "f"(a0));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(l1) : "h"(h1), "h"(m1), "f"(a1));
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo[c]) : "f"(l1), "f"(l0));
  }
}

// Selection for n <= 512 in the 8-warp kernel: thread (row, hf) keeps the keys of its points row + 128 (2 j + hf), j = 0, 1.
// Per-warp REDUX rounds give each of the 8 warps its cnt smallest in ascending order; every candidate then computes its rank
// among the 8 x cnt candidates and the lanes of rank < cnt write the output rows.
__device__ __forceinline__ void select_and_write_reg8(const DuneParams& prm, const ItemFrame& fr, int b, int t, int n, int cnt, int tid, uint32_t k0,
                                                      uint32_t k1, const float* smu, unsigned long long* cands) {
  const int M = prm.M, E = prm.geo.E, T1 = prm.T + 1, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, hf = tid >> 7;
  unsigned long long* cand = cands + warp * M;
  for (int m = 0; m < cnt; ++m) {
    const uint32_t bd = min(k0, k1);
    const uint32_t md = __reduce_min_sync(0xffffffffu, bd);
    const int j = k0 == md ? 0 : 1;
    const uint32_t bi = bd == md ? (uint32_t)(row + 128 * (2 * j + hf)) : 0xFFFFFFFFu;
    const uint32_t mi = __reduce_min_sync(0xffffffffu, bi);
    if (md != 0xFFFFFFFFu && bi == mi) {  // the owner retires the key
      if (j == 0) k0 = 0xFFFFFFFFu;
      else k1 = 0xFFFFFFFFu;
    }
    if (lane == 0) cand[m] = md == 0xFFFFFFFFu ? ~0ull : (((unsigned long long)md << 32) | mi);
  }
  __syncthreads();
  if (lane < cnt) {
    const unsigned long long mine = cand[lane];
    int rank = lane;  // candidates of the own warp are sorted and distinct
    for (int w = 0; w < 8; ++w) {
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

// generic selection (any n): keys in shared memory, 8 warps; same algorithm as select_and_write with a stride of 256
__device__ __forceinline__ void select_and_write8(const DuneParams& prm, const ItemFrame& fr, int b, int t, int n, int cnt, int warp, int lane,
                                                  unsigned long long* keys, const float* smu, unsigned long long* cands) {
  const int M = prm.M, E = prm.geo.E, T1 = prm.T + 1;
  unsigned long long mine = ~0ull;
  {
    unsigned long long* cand = cands + warp * M;
    for (int m = 0; m < cnt; ++m) {
      unsigned bd = 0xFFFFFFFFu, bi = 0xFFFFFFFFu;
      for (int i = warp * 32 + lane; i < n; i += 256) {
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
  if (warp == 0) {  // merge the 8 candidate lists
    const int total = 8 * cnt;
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
      float gx, gy;
      fr.world(idx, gx, gy);
      const size_t o = ((size_t)b * T1 + t) * M + lane;
      float lx = 0.f, ly = 0.f;
#pragma unroll
      for (int e = 0; e < kMaxEdges; ++e) {
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
      if (t == 0 && lane == 0 && prm.min_dist) prm.min_dist[b] = d;
    }
  }
}

}  // namespace tc

// shared memory: operand image | keys (N x 8 B, generic path) | smu (N x E floats) | candidates (8 warps x M)
__host__ __device__ inline size_t dune_tc8_smem_bytes(int N, int E, int M) {
  return TcImage::kBytes + (size_t)N * 8 + (((size_t)N * E * 4 + 7) / 8) * 8 + (size_t)8 * M * 8 + 64;
}

template <bool kFast, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks) dune_tc8_kernel(const DuneParams prm, const unsigned char* __restrict__ image) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  using I = TcImage;
  unsigned char* simg = smem_dyn;
  const float* fl = reinterpret_cast<const float*>(simg + I::kFloatOff);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(simg + I::kBytes);
  float* smu = reinterpret_cast<float*>(simg + I::kBytes + (size_t)prm.N * 8);
  unsigned long long* cands = reinterpret_cast<unsigned long long*>(simg + I::kBytes + (size_t)prm.N * 8 + (((size_t)prm.N * prm.geo.E * 4 + 7) / 8) * 8);
  __shared__ __align__(8) unsigned long long mbar[2];  // D of slot 0 / 1 ready (tcgen05.commit)
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = tid & 127, hf = tid >> 7;  // hf is warp-uniform
  for (int i = tid; i < I::kBytes / 16; i += 256) reinterpret_cast<uint4*>(simg)[i] = reinterpret_cast<const uint4*>(image)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(&mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::smem_u32(&mbar[1])));
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
  const uint32_t trow = tbase + ((uint32_t)((warp & 3) * 32) << 16);  // warps w and w + 4 share a TMEM lane quarter
  const uint32_t simg_u = tc::smem_u32(simg);
  const uint32_t bar0 = tc::smem_u32(&mbar[0]);
  const uint32_t desc_w = tc::desc_lo(simg_u, 512), desc_ones = tc::desc_lo(simg_u + I::kOnesOff, 2048);
  uint32_t phases = 0;

  auto publish = [&](const uint32_t (&hi)[8], const uint32_t (&lo)[8], int slot, int layer) {
    const uint32_t tS = trow + 64 * slot;
    tc::st8(tS + 32 + 8 * hf, hi);
    tc::st8(tS + 48 + 8 * hf, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    const int issuer = (layer + slot + (int)blockIdx.x) & 7;  // rotate over the 8 warps (two per scheduler)
    __syncthreads();
    if (warp == issuer) {
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
    const tc::ItemFrame fr = tc::item_frame(prm, b, t);
    const bool reg_keys = n <= 512;
    uint32_t k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;

#pragma unroll 1
    for (int base = 0; base < n; base += 256) {
      const int nslots = base + 128 < n ? 2 : 1;
#pragma unroll 1
      for (int sl = 0; sl < nslots; ++sl) {  // stage 0: layer 0 (2 -> 32, both threads in full) + LayerNorm/tanh of the own half
        int i = base + sl * 128 + row;
        i = i < n ? i : n - 1;
        float x0, y0;
        fr.local(i, x0, y0);
        const tc::f2 x2 = tc::pk(x0, x0), y2 = tc::pk(y0, y0);
        tc::f2 own[8];
        float qo, qw;
        {
          tc::f2 oth[8];
          const int oo = 16 * (1 - hf), ow = 16 * hf;
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            const ulonglong2 wx = *reinterpret_cast<const ulonglong2*>(fl + I::W0X + oo + 2 * c);
            const ulonglong2 wy = *reinterpret_cast<const ulonglong2*>(fl + I::W0Y + oo + 2 * c);
            const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(fl + I::B0 + oo + 2 * c);
            oth[c] = tc::fma2(wy.x, y2, tc::fma2(wx.x, x2, bb.x));
            oth[c + 1] = tc::fma2(wy.y, y2, tc::fma2(wx.y, x2, bb.y));
          }
          qo = tc::sumsq16(oth);
#pragma unroll
          for (int c = 0; c < 8; c += 2) {
            const ulonglong2 wx = *reinterpret_cast<const ulonglong2*>(fl + I::W0X + ow + 2 * c);
            const ulonglong2 wy = *reinterpret_cast<const ulonglong2*>(fl + I::W0Y + ow + 2 * c);
            const ulonglong2 bb = *reinterpret_cast<const ulonglong2*>(fl + I::B0 + ow + 2 * c);
            own[c] = tc::fma2(wy.x, y2, tc::fma2(wx.x, x2, bb.x));
            own[c + 1] = tc::fma2(wy.y, y2, tc::fma2(wx.y, x2, bb.y));
          }
          qw = tc::sumsq16(own);
        }
        float r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(qw + qo, 1.0f / 32, 1e-5f)));
        uint32_t hi[8], lo[8];
        tc::tanh_split16<kFast>(own, r, fl + I::G1 + 16 * hf, fl + I::BE1 + 16 * hf, hi, lo);
        publish(hi, lo, sl, 0);
      }
#pragma unroll 1
      for (int st = 1; st < 5; ++st) {
#pragma unroll 1
        for (int sl = 0; sl < nslots; ++sl) {
          tc::f2 own[8];
          uint32_t hi[8], lo[8];
          acquire(sl);
          const uint32_t tD = trow + 64 * sl;
          if (st & 1) {
            tc::ld16p(tD + 16 * hf, own);
            tc::relu_split16(own, hi, lo);
          } else {
            float qo;
            {
              tc::f2 oth[8];
              tc::ld16p(tD + 16 * (1 - hf), oth);
              qo = tc::sumsq16(oth);
            }
            tc::ld16p(tD + 16 * hf, own);
            const float qw = tc::sumsq16(own);
            float r;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(qw + qo, 1.0f / 32, 1e-5f)));
            tc::tanh_split16<kFast>(own, r, fl + I::G1 + 32 * st + 16 * hf, fl + I::BE1 + 32 * st + 16 * hf, hi, lo);
          }
          publish(hi, lo, sl, st);
        }
      }
      // head epilogue: slot s by the threads with hf == s (every thread waits for its own slot only)
      if (hf < nslots) {
        float mu[8];
        acquire(hf);
        tc::ld8(trow + 64 * hf, mu);
        int i = base + hf * 128 + row;
        const bool valid = i < n;
        i = valid ? i : n - 1;
        float x0, y0;
        fr.local(i, x0, y0);
        const uint32_t key = tc::finish_point(prm, mu, x0, y0, i, valid, E, smu, reg_keys ? nullptr : keys);
        if (base == 0) k0 = key;
        else if (base == 256) k1 = key;
      }
      // the other half's phase bookkeeping: the slot it did not wait for completed one more phase
      if (nslots == 2) phases ^= 1u << (1 - hf);
      else if (hf == 1) phases ^= 1u;
      __syncthreads();  // D of both slots has been read before the next pass (or item) overwrites the operands / D
    }
    // ---- top-M and the output rows -----------------------------------------------------------------------------------
    if (reg_keys) tc::select_and_write_reg8(prm, fr, b, t, n, cnt, tid, k0, k1, smu, cands);
    else tc::select_and_write8(prm, fr, b, t, n, cnt, warp, lane, keys, smu, cands);
    __syncthreads();  // smu / cands (and keys) are reused by the next item
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}

}  // namespace nb
