// Host-side launcher of the DUNE kernel, one explicit instantiation per edge count E
// (dune_inst.cu is compiled once per E with -DNB_E=<E>, in parallel, to keep build time down).
#pragma once
#include <cstdio>

#include "dune_kernel.cuh"

namespace nb {

// returns 0 on success; on failure writes a message into err and returns a negative NB_ERR_* code
template <int E>
int launch_dune_e(const DuneParams& prm, int sm_count, int max_smem_optin, cudaStream_t st, char* err, size_t errlen);

#ifdef NB_E
template <int E>
int launch_dune_e(const DuneParams& prm, int sm_count, int max_smem_optin, cudaStream_t st, char* err, size_t errlen) {
  const int N = prm.N;
  const size_t smem = dune_smem_bytes<E>(N);
  if ((long long)smem > max_smem_optin) {
    snprintf(err, errlen, "N=%d needs %zu B of shared memory (limit %d)", N, smem, max_smem_optin);
    return -3;
  }
  const int items = prm.B * (prm.T + 1);
  auto go = [&](auto kern, int threads) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e == cudaSuccess) {
      if (per_sm < 1) per_sm = 1;
      int grid = sm_count * per_sm;
      if (grid > items) grid = items;
      kern<<<grid, threads, smem, st>>>(prm);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
      snprintf(err, errlen, "dune_kernel<E=%d> launch failed: %s", E, cudaGetErrorString(e));
      return -2;
    }
    return 0;
  };
  if (N <= 128) return go(dune_kernel<E, 1, 128>, 128);
  if (N <= 256) return go(dune_kernel<E, 1, 256>, 256);
  return go(dune_kernel<E, 2, 256>, 256);
}
#endif

}  // namespace nb
