// Translation unit of the tcgen05 DUNE kernel: host-side operand image + launcher.
#include <cstdio>
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
  const int hidden_w[4] = {L::W3, L::W5, L::W8, L::W10}, hidden_b[4] = {L::B3, L::B5, L::B8, L::B10};
  for (int l = 0; l < 4; ++l) {
    const size_t base = (size_t)l * I::kHiddenStride;
    for (int n = 0; n < 32; ++n) {
      for (int k = 0; k < 32; ++k) {
        uint16_t hi, lo;
        split_half_tc(w[hidden_w[l] + n * 32 + k], hi, lo);
        put(base + at(n, k, 512, 1024), hi);
        put(base + 2048 + at(n, k, 512, 1024), lo);
      }
      uint16_t hi, lo;
      split_half_tc(w[hidden_b[l] + n], hi, lo);
      put(base + 4096 + at(n, 0, 512, 1024), hi);
      put(base + 4096 + at(n, 1, 512, 1024), lo);
    }
  }
  for (int n = 0; n < E; ++n) {  // head, N padded to 16 with zero rows
    for (int k = 0; k < 32; ++k) {
      uint16_t hi, lo;
      split_half_tc(w[L::W13 + n * 32 + k], hi, lo);
      put(I::kHeadOff + at(n, k, 256, 512), hi);
      put(I::kHeadOff + 1024 + at(n, k, 256, 512), lo);
    }
    uint16_t hi, lo;
    split_half_tc(w[L::b13(E) + n], hi, lo);
    put(I::kHeadOff + 2048 + at(n, 0, 256, 512), hi);
    put(I::kHeadOff + 2048 + at(n, 1, 256, 512), lo);
  }
  float* fl = reinterpret_cast<float*>(out.data() + I::kFloatOff);
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
  const size_t smem = dune_tc_smem_bytes(prm.N, prm.geo.E, prm.M);
  if ((long long)smem > max_smem_optin) {
    snprintf(err, errlen, "N=%d needs %zu B of shared memory (limit %d)", prm.N, smem, max_smem_optin);
    return -3;
  }
  const int items = prm.B * (prm.T + 1);
  cudaError_t e = cudaFuncSetAttribute(dune_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dune_tc_kernel, 128, smem);
  if (e == cudaSuccess) {
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;  // 4 x 128 TMEM columns per SM
    int grid = sm_count * per_sm;
    if (grid > items) grid = items;
    dune_tc_kernel<<<grid, 128, smem, st>>>(prm, d_image);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "dune_tc_kernel launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace nb
