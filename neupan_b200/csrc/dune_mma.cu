// Translation unit of the tensor-core DUNE kernel: host-side weight image + launcher.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dune_mma_kernel.cuh"

namespace nb {

static inline void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
static inline uint32_t pack2(__half a, __half b) {
  uint16_t x, y;
  memcpy(&x, &a, 2); memcpy(&y, &b, 2);
  return (uint32_t)x | ((uint32_t)y << 16);
}

// packed checkpoint (WeightLayout order, E outputs) -> MmaImage bytes
void build_mma_image(const float* w, int E, std::vector<unsigned char>& out) {
  using L = WeightLayout;
  using I = MmaImage;
  out.assign(I::kBytes, 0);
  uint32_t* frag = reinterpret_cast<uint32_t*>(out.data());
  float* fl = reinterpret_cast<float*>(out.data() + (size_t)I::kFragU4 * 16);
  const int hidden_off[4] = {L::W3, L::W5, L::W8, L::W10};
  auto emit = [&](uint32_t* dst, const float* W, int rows, int NT) {  // W: rows x 32 (out x in), rows <= 8*NT
    for (int s = 0; s < 2; ++s)
      for (int j = 0; j < NT; ++j)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3, n = 8 * j + g, k0 = 16 * s + 2 * t;
          float v[4] = {0, 0, 0, 0};
          if (n < rows) { v[0] = W[n * 32 + k0]; v[1] = W[n * 32 + k0 + 1]; v[2] = W[n * 32 + k0 + 8]; v[3] = W[n * 32 + k0 + 9]; }
          __half hi[4], lo[4];
          for (int i = 0; i < 4; ++i) split_half(v[i], hi[i], lo[i]);
          uint32_t* u = dst + ((size_t)(s * NT + j) * 32 + lane) * 4;
          u[0] = pack2(hi[0], hi[1]); u[1] = pack2(hi[2], hi[3]); u[2] = pack2(lo[0], lo[1]); u[3] = pack2(lo[2], lo[3]);
        }
  };
  for (int l = 0; l < 4; ++l) emit(frag + (size_t)l * I::kHiddenFragU4 * 4, w + hidden_off[l], 32, 4);
  emit(frag + (size_t)4 * I::kHiddenFragU4 * 4, w + L::W13, E, 1);
  auto cp = [&](int dst, int src, int n) { memcpy(fl + dst, w + src, (size_t)n * 4); };
  // LayerNorm gain / offset pre-multiplied by 2*log2(e): tanh(y) = 1 - 2/(exp2(2*log2(e)*y) + 1)
  auto cps = [&](int dst, int src, int n) {
    for (int i = 0; i < n; ++i) fl[dst + i] = (float)((double)w[src + i] * 2.8853900817779268);
  };
  cp(I::W0, L::W0, 64); cp(I::B0, L::B0, 32); cps(I::G1, L::G1, 32); cps(I::BE1, L::BE1, 32);
  cp(I::B3, L::B3, 32); cp(I::B5, L::B5, 32); cps(I::G6, L::G6, 32); cps(I::BE6, L::BE6, 32);
  cp(I::B8, L::B8, 32); cp(I::B10, L::B10, 32); cps(I::G11, L::G11, 32); cps(I::BE11, L::BE11, 32);
  cp(I::B13, L::b13(E), E);
}

int launch_dune_mma(const DuneParams& prm, const unsigned char* d_image, int sm_count, int max_smem_optin, int cta_per_sm_limit,
                    cudaStream_t st, char* err, size_t errlen) {
  const int N = prm.N;
  const int items = prm.B * (prm.T + 1);
  int warps = 8;
  while (warps > 1 && (long long)dune_mma_smem_bytes(N, warps) > max_smem_optin) warps >>= 1;
  if (items < sm_count * 8) warps = items < sm_count * 2 ? 1 : 2;  // few items: spread the warps over the SMs
  const size_t smem = dune_mma_smem_bytes(N, warps);
  if ((long long)smem > max_smem_optin) {
    snprintf(err, errlen, "N=%d needs %zu B of shared memory (limit %d)", N, smem, max_smem_optin);
    return -3;
  }
  const int threads = warps * 32;
  static int variant = -1;  // NB_DUNE_MT: developer switch between the 32-point (2 tiles) and 16-point (1 tile) warp pass
  if (variant < 0) {
    const char* v = getenv("NB_DUNE_MT");
    variant = v ? atoi(v) : 2;
  }
  auto kern = variant == 1 ? dune_mma_kernel<1, 3> : dune_mma_kernel<2, 2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
  if (e == cudaSuccess) {
    if (per_sm < 1) per_sm = 1;
    if (cta_per_sm_limit > 0 && per_sm > cta_per_sm_limit) per_sm = cta_per_sm_limit;
    int grid = sm_count * per_sm;
    const int need = (items + warps - 1) / warps;
    if (grid > need) grid = need;
    kern<<<grid, threads, smem, st>>>(prm, d_image);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "dune_mma_kernel launch failed: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace nb
