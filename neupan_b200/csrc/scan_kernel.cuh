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
//   pass 1  beams in chunks of blockDim: keep flag -> rank by ballot / popc + per-warp offsets; the beam index of every
//           point that survives the filter AND the ::down_sample stride goes to a shared-memory list (order preserved)
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
__device__ __forceinline__ double linspace_at(double start, double stop, int num, int i) {
  if (num == 1) return start;
  if (i == num - 1) return stop;
  const double step = (stop - start) / (double)(num - 1);
  return __dadd_rn(__dmul_rn((double)i, step), start);
}

__global__ void __launch_bounds__(256) scan_to_points_kernel(const ScanParams prm) {
  extern __shared__ int32_t keep_list[];  // R entries
  __shared__ int warp_total[8];
  __shared__ int base_s;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float* rng = prm.ranges + (size_t)b * prm.R;
  const double upper = prm.range_max - 0.02;  // neupan.py:207 / :258
  if (tid == 0) base_s = 0;
  __syncthreads();

  // ---- pass 1: ordered compaction of the surviving beam indices -----------------------------------
  for (int c0 = 0; c0 < prm.R; c0 += blockDim.x) {
    const int i = c0 + tid;
    bool keep = false;
    if (i < prm.R) {
      const double r = (double)rng[i];
      const double a = linspace_at(prm.angle_min, prm.angle_max, prm.R, i);
      const bool lo = prm.velocity_mode ? (r >= prm.range_min) : (r > prm.range_min);  // neupan.py:258 vs :207
      keep = r < upper && lo && a > prm.angle_lo && a < prm.angle_hi;
    }
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
  const int kept = base_s;
  const int n = (kept + prm.down_sample - 1) / prm.down_sample;  // columns after the stride
  const int m_out = n > prm.max_points ? prm.max_points : n;     // pan.py:171: decimate only when n > dune_max_num
  if (tid == 0) prm.counts[b] = m_out;

  // ---- pass 2: one output column per thread --------------------------------------------------------
  const double sx = prm.states[3 * b], sy = prm.states[3 * b + 1], sth = prm.states[3 * b + 2];
  const double cr = cos(sth), sr = sin(sth), co = cos(prm.off_th), so = sin(prm.off_th);
  float* px = prm.points + (size_t)b * 2 * prm.max_points;
  float* py = px + prm.max_points;
  for (int j = tid; j < m_out; j += blockDim.x) {
    int k = j;
    if (n > prm.max_points) k = (int)linspace_at(0.0, (double)(n - 1), prm.max_points, j);  // np.linspace(0, n-1, m).astype(int)
    const int i = keep_list[k];
    const double r = (double)rng[i];
    const double a = linspace_at(prm.angle_min, prm.angle_max, prm.R, i);
    const double lx = r * cos(a), ly = r * sin(a);
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

}  // namespace nb
