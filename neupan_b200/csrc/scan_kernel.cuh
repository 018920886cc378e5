// Lidar scan -> obstacle points, batched over environments (SURVEY 8f "next" row 2).
//
// Restates neupan.scan_to_point (neupan/neupan.py:173-222) and neupan.scan_to_point_velocity (neupan.py:224-281), followed by
// the decimation PAN applies when a scan yields more than dune_max_num points (pan.py:171-174 ->
// util.downsample_decimation, neupan/util/__init__.py:285-305), so that the result can be handed to nb_pan_forward as
// (points, velocities, num_points) without leaving the device.
//
// The reference does this arithmetic in float64 (numpy / math.cos) and the planner casts the points to float32
// (neupan.py:123-126); the kernel follows the same order of operations in FP64 (explicit __dmul_rn / __dadd_rn where numpy
// rounds twice) and casts at the store.  It is a byte-moving kernel: 4 B read per beam, 8 (+8) B written per kept point.
//
// One CTA per environment:
//   pass 1  keep flag per beam -> rank by ballot / popc + per-warp offsets; the beam index of every point that survives
//           the filter AND the ::down_sample stride goes to a shared-memory list (order preserved)
//   pass 2  thread j produces output column j: list lookup (through the linspace decimation map when the list is longer
//           than max_points), polar -> sensor frame -> robot frame -> world frame, coalesced float32 stores
#pragma once
#include <cstdint>

namespace nb {

struct ScanParams {
  int B, R, max_points;
  const float* ranges;     // (B, R)
  const float* velocity;   // (B, 2, R) or null
  const double* states;    // (B, 3): x, y, theta of the robot
  double angle_min, angle_max, range_min, range_max;
  double off_x, off_y, off_th;
  double angle_lo, angle_hi;
  int down_sample;
  int velocity_mode;       // 0: scan_to_point semantics, 1: scan_to_point_velocity semantics
  float* points;           // (B, 2, max_points)
  float* vel_out;          // (B, 2, max_points) or null
  int32_t* counts;         // (B)
};

// element i of numpy.linspace(start, stop, num): arange(num) * step + start with the last element set to stop
__device__ __forceinline__ double linspace_step(double start, double stop, int num) { return num > 1 ? (stop - start) / (double)(num - 1) : 0.0; }
__device__ __forceinline__ double linspace_at(double start, double stop, double step, int num, int i) {
  if (num == 1) return start;
  if (i == num - 1) return stop;
  return __dadd_rn(__dmul_rn((double)i, step), start);
}

__device__ __forceinline__ bool scan_keep(const ScanParams& prm, double r, double a, double upper) {
  const bool lo = prm.velocity_mode ? (r >= prm.range_min) : (r > prm.range_min);  // neupan.py:258 vs :207
  return r < upper && lo && a > prm.angle_lo && a < prm.angle_hi;
}

// pass 2 of both kernels: thread j produces output column j from the compacted beam list
__device__ __forceinline__ void scan_emit(const ScanParams& prm, int b, const int32_t* keep_list, int n, double astep) {
  const int m_out = n > prm.max_points ? prm.max_points : n;  // pan.py:171: decimate only when n > dune_max_num
  if (threadIdx.x == 0) prm.counts[b] = m_out;
  const float* rng = prm.ranges + (size_t)b * prm.R;
  const double sx = prm.states[3 * b], sy = prm.states[3 * b + 1], sth = prm.states[3 * b + 2];
  double cr, sr, co, so;
  sincos(sth, &sr, &cr);
  sincos(prm.off_th, &so, &co);
  const double dstep = linspace_step(0.0, (double)(n - 1), prm.max_points);
  float* px = prm.points + (size_t)b * 2 * prm.max_points;
  float* py = px + prm.max_points;
  for (int j = threadIdx.x; j < m_out; j += blockDim.x) {
    int k = j;
    if (n > prm.max_points) k = (int)linspace_at(0.0, (double)(n - 1), dstep, prm.max_points, j);  // np.linspace(0, n-1, m).astype(int)
    const int i = keep_list[k];
    const double r = (double)rng[i];
    const double a = linspace_at(prm.angle_min, prm.angle_max, astep, prm.R, i);
    double sa, ca;
    sincos(a, &sa, &ca);
    const double lx = r * ca, ly = r * sa;
    double tx, ty;
    if (prm.velocity_mode) {  // s_R^T (p - s_trans)   (neupan.py:271-273)
      const double dx = lx - prm.off_x, dy = ly - prm.off_y;
      tx = __dadd_rn(__dmul_rn(co, dx), __dmul_rn(so, dy));
      ty = __dadd_rn(__dmul_rn(-so, dx), __dmul_rn(co, dy));
    } else {  // s_R p + s_trans   (neupan.py:216-217)
      tx = __dadd_rn(__dadd_rn(__dmul_rn(co, lx), __dmul_rn(-so, ly)), prm.off_x);
      ty = __dadd_rn(__dadd_rn(__dmul_rn(so, lx), __dmul_rn(co, ly)), prm.off_y);
    }
    const double wx = __dadd_rn(__dadd_rn(__dmul_rn(cr, tx), __dmul_rn(-sr, ty)), sx);  // R temp + trans   (neupan.py:219-220 / :275-276)
    const double wy = __dadd_rn(__dadd_rn(__dmul_rn(sr, tx), __dmul_rn(cr, ty)), sy);
    px[j] = (float)wx;
    py[j] = (float)wy;
    if (prm.vel_out) {
      float* vx = prm.vel_out + (size_t)b * 2 * prm.max_points;
      const float* vin = prm.velocity ? prm.velocity + (size_t)b * 2 * prm.R : nullptr;
      vx[j] = vin ? vin[i] : 0.f;                       // scan.get("velocity", zeros)   (neupan.py:250)
      vx[prm.max_points + j] = vin ? vin[prm.R + i] : 0.f;
    }
  }
}

// Fast kernel (R <= 8 * blockDim): warp w owns the contiguous beams [w*CH*32, (w+1)*CH*32); all of a thread's ranges are
// loaded up front (independent loads in flight), the keep masks of the warp's CH chunks stay in registers, and ONE block
// barrier publishes the per-warp totals -- instead of three barriers and a dependent load per 256-beam chunk.
constexpr int kScanMaxChunks = 8;
__global__ void __launch_bounds__(1024) scan_to_points_fast_kernel(const ScanParams prm) {
  extern __shared__ int32_t keep_list[];  // R entries
  __shared__ int warp_total[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float* rng = prm.ranges + (size_t)b * prm.R;
  const double upper = prm.range_max - 0.02;  // neupan.py:207 / :258
  const double astep = linspace_step(prm.angle_min, prm.angle_max, prm.R);
  const int CH = (prm.R + blockDim.x - 1) / blockDim.x;  // chunks of 32 beams per warp (<= kScanMaxChunks)
  const int seg0 = warp * CH * 32;
  float rv[kScanMaxChunks];
#pragma unroll
  for (int c = 0; c < kScanMaxChunks; ++c) {
    const int i = seg0 + c * 32 + lane;
    rv[c] = (c < CH && i < prm.R) ? rng[i] : -1.0f;
  }
  unsigned masks[kScanMaxChunks];
  int cnt = 0;
#pragma unroll
  for (int c = 0; c < kScanMaxChunks; ++c) {
    const int i = seg0 + c * 32 + lane;
    bool keep = false;
    if (c < CH && i < prm.R) keep = scan_keep(prm, (double)rv[c], linspace_at(prm.angle_min, prm.angle_max, astep, prm.R, i), upper);
    masks[c] = __ballot_sync(0xffffffffu, keep);
    cnt += __popc(masks[c]);
  }
  if (lane == 0) warp_total[warp] = cnt;
  __syncthreads();
  int running = 0, kept = 0;
  for (int w = 0; w < nwarps; ++w) {
    const int t = warp_total[w];
    if (w < warp) running += t;
    kept += t;
  }
#pragma unroll
  for (int c = 0; c < kScanMaxChunks; ++c) {
    if ((masks[c] >> lane) & 1u) {
      const int rank = running + __popc(masks[c] & ((1u << lane) - 1u));
      if (rank % prm.down_sample == 0) keep_list[rank / prm.down_sample] = seg0 + c * 32 + lane;  // [:, ::down_sample]
    }
    running += __popc(masks[c]);
  }
  __syncthreads();
  scan_emit(prm, b, keep_list, (kept + prm.down_sample - 1) / prm.down_sample, astep);
}

// Generic kernel (any R that fits the shared-memory list): 256-beam chunks with a running base
__global__ void __launch_bounds__(256) scan_to_points_kernel(const ScanParams prm) {
  extern __shared__ int32_t keep_list[];  // R entries
  __shared__ int warp_total[8];
  __shared__ int base_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float* rng = prm.ranges + (size_t)b * prm.R;
  const double upper = prm.range_max - 0.02;  // neupan.py:207 / :258
  const double astep = linspace_step(prm.angle_min, prm.angle_max, prm.R);
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int c0 = 0; c0 < prm.R; c0 += blockDim.x) {
    const int i = c0 + tid;
    const bool keep = i < prm.R && scan_keep(prm, (double)rng[i], linspace_at(prm.angle_min, prm.angle_max, astep, prm.R, i), upper);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_total[warp] = __popc(m);
    __syncthreads();
    int rank = base_s + __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) rank += warp_total[w];
    if (keep && rank % prm.down_sample == 0) keep_list[rank / prm.down_sample] = i;  // [:, ::down_sample]
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < nwarps; ++w) t += warp_total[w];
      base_s += t;
    }
    __syncthreads();
  }
  scan_emit(prm, b, keep_list, (base_s + prm.down_sample - 1) / prm.down_sample, astep);
}

}  // namespace nb
