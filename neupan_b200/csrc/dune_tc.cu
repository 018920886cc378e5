// Translation unit of the tcgen05 DUNE kernel: host-side operand image + launcher.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dune_tc_kernel.cuh"

namespace nb {

static inline void split_half_tc(float v, uint16_t& hi, uint16_t& lo) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  memcpy(&hi, &h, 2);
  memcpy(&lo, &l, 2);
}

// packed checkpoint (WeightLayout order, E outputs) -> TcImage bytes (canonical K-major / no-swizzle UMMA layout)
void build_tc_image(const float* w, int E, std::vector<unsigned char>& out) {
  using L = WeightLayout;
  using I = TcImage;
  out.assign(I::kBytes, 0);
  auto put = [&](size_t off, uint16_t v) { memcpy(out.data() + off, &v, 2); };
  // element (n, k) of a K-major operand with `lbo` bytes between the two 8-half K groups and `kstep` bytes per 16 of K
  auto at = [](int n, int k, int lbo, int kstep) { return (size_t)(k / 16) * kstep + ((k % 16) / 8) * lbo + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2; };
  float* fl = reinterpret_cast<float*>(out.data() + I::kFloatOff);
  const int hidden_w[4] = {L::W3, L::W5, L::W8, L::W10}, hidden_b[4] = {L::B3, L::B5, L::B8, L::B10};
  // layers fed by a tanh (MLP.3, MLP.8 and the head MLP.13) receive r = 1/(exp(2y)+1) instead of tanh(y) = 1 - 2r:
  //   W tanh + b = (b + rowsum(W)) + (-2 W) r
  const bool after_tanh[4] = {true, false, true, false};
  for (int l = 0; l < 4; ++l) {
    const size_t base = (size_t)l * I::kHiddenStride;
    for (int n = 0; n < 32; ++n) {
      double rowsum = 0.0;
      for (int k = 0; k < 32; ++k) {
        const float wv = w[hidden_w[l] + n * 32 + k];
        rowsum += (double)wv;
        uint16_t hi, lo;
        split_half_tc(after_tanh[l] ? -2.0f * wv : wv, hi, lo);
        put(base + at(n, k, 512, 1024), hi);
        put(base + 2048 + at(n, k, 512, 1024), lo);
      }
      fl[I::BH + 32 * l + n] = (float)((double)w[hidden_b[l] + n] + (after_tanh[l] ? rowsum : 0.0));
    }
  }
  for (int n = 0; n < E; ++n) {  // head (after a tanh), N padded to 16 with zero rows
    double rowsum = 0.0;
    for (int k = 0; k < 32; ++k) {
      const float wv = w[L::W13 + n * 32 + k];
      rowsum += (double)wv;
      uint16_t hi, lo;
      split_half_tc(-2.0f * wv, hi, lo);
      put(I::kHeadOff + at(n, k, 256, 512), hi);
      put(I::kHeadOff + 1024 + at(n, k, 256, 512), lo);
    }
    fl[I::BHEAD + n] = (float)((double)w[L::b13(E) + n] + rowsum);
  }
  memcpy(fl + I::W0, w + L::W0, 64 * 4);
  memcpy(fl + I::B0, w + L::B0, 32 * 4);
  const int g_src[3] = {L::G1, L::G6, L::G11}, b_src[3] = {L::BE1, L::BE6, L::BE11};
  const int g_dst[3] = {I::G1, I::G6, I::G11}, b_dst[3] = {I::BE1, I::BE6, I::BE11};
  for (int q = 0; q < 3; ++q)
    for (int i = 0; i < 32; ++i) {  // pre-multiplied by 2*log2(e): tanh(y) = 1 - 2/(exp2(2*log2(e)*y) + 1)
      fl[g_dst[q] + i] = (float)((double)w[g_src[q] + i] * 2.8853900817779268);
      fl[b_dst[q] + i] = (float)((double)w[b_src[q] + i] * 2.8853900817779268);
    }
}

int launch_dune_tc(const DuneParams& prm, const unsigned char* d_image, int sm_count, int max_smem_optin, cudaStream_t st, char* err, size_t errlen) {
  size_t smem = dune_tc_smem_bytes(prm.N, prm.geo.E, prm.M);
  if ((long long)smem > max_smem_optin) {
    snprintf(err, errlen, "N=%d needs %zu B of shared memory (limit %d)", prm.N, smem, max_smem_optin);
    return -3;
  }
  const int items = prm.B * (prm.T + 1);
  static int force = -1;  // NB_DUNE_TC: developer switch, 1 = single-slot kernel, 2 = ping-pong kernel
  if (force < 0) {
    const char* v = getenv("NB_DUNE_TC");
    force = v ? atoi(v) : 0;
  }
  // The single-slot kernel with 5 CTAs per SM is the default: measured 2.49 ms per launch at C4 against 3.14 ms for the
  // ping-pong kernel (two tiles in flight per CTA + issuer warp; its polling issuer and 4-CTA limit cost more than the
  // intra-CTA overlap gains).  NB_DUNE_TC=2 selects the ping-pong kernel for experiments.
  const bool pingpong = force == 2;
  // TMEM columns are held by a CTA for its whole (persistent) lifetime: 512 / columns-per-CTA CTAs may share an SM, one
  // more would sit in tcgen05.alloc until another CTA exits, and the block scheduler knows nothing about TMEM
  // (observed: a 5th CTA with 128 columns landing on an SM turned 2.7 ms into 4.3 ms per launch).  The shared-memory
  // request is therefore padded so that never more CTAs fit than registers and TMEM admit.
  const int threads = pingpong ? 160 : 128;
  const int want = pingpong ? 4 : 5;  // ping-pong: 128 columns -> 4 CTAs;  single slot: 64 columns, 96 registers -> 5 CTAs
  // smallest request that keeps a (want+1)-th CTA out (anything larger only shrinks the L1 cache: 44 KB instead of 38 KB
  // per CTA cost 2.49 -> 3.50 ms per launch)
  const size_t pad = (size_t)(233472 / (want + 1)) - 2048 + 512;
  if (smem < pad) smem = pad;
  cudaError_t e = pingpong ? cudaFuncSetAttribute(dune_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(dune_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    int per_sm = (int)(233472 / (smem + 2048));
    if (per_sm > want) per_sm = want;
    if (per_sm < 1) per_sm = 1;
    int grid = sm_count * per_sm;
    if (grid > items) grid = items;
    if (pingpong) dune_tc2_kernel<<<grid, threads, smem, st>>>(prm, d_image);
    else dune_tc_kernel<<<grid, threads, smem, st>>>(prm, d_image);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "dune_tc kernel launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace nb
