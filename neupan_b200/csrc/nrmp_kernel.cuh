// NRMP kernel: one warp solves one environment's convex program per launch.
//
// Replaces NRMP.forward (neupan/blocks/nrmp.py:114-150) including
//   robot.generate_state_parameter_value / linear_{ackermann,diff,omni}_model  robot.py:239-316
//   NRMP.generate_coefficient_parameter_value (fa, fb, padding)                nrmp.py:220-261
//   the cvxpylayers -> diffcp -> ECOS solve of the program built at           nrmp.py:263-383,
//                                                                              robot.py:73-236
//   PAN.stop_criteria                                                          pan.py:215-243
//
// Method (derivation mirrored by the CPU checker oracle/ipm.py): the states are eliminated through
// the linearised dynamics, s_{t+1} = s0_{t+1} + F_t u; the squared hinge of robot.py:183-198 is
// lifted with slacks w_tm >= 0, w_tm >= D_t + fb_tm - fa_tm.s_{t+1,xy}; a Mehrotra
// predictor-corrector primal-dual interior point method runs in FP64 from a strictly feasible
// start.  Inside each Newton step the T*M hinge slacks and the T distances D_t are eliminated
// analytically (both blocks are diagonal), so the only factorisation is a dense 2T x 2T Cholesky
// held in the warp's shared memory.  Rows of every matrix are owned by lanes; all
// synchronisation is __syncwarp, a warp can finish early without affecting its neighbours.
#pragma once
#include "common.cuh"

namespace nb {

struct NrmpParams {
  const float* nom_s;   // (B,3,T+1) nominal states in
  const float* nom_u;   // (B,2,T)
  const float* ref_s;   // (B,3,T+1)
  const float* ref_us;  // (B,T)
  const float* fa;      // (B,T,M,2) explicit coefficients or nullptr
  const float* fb;      // (B,T,M)
  const float* sel_mu;  // (B,T+1,M,E) DUNE selections (used when fa == nullptr) or nullptr
  const float* sel_lam; // (B,T+1,M,2)
  const float* sel_pts; // (B,T+1,M,2)
  const int32_t* sel_count;  // (B)
  float* out_s;         // (B,3,T+1)  may alias nom_s (each warp reads its env before writing)
  float* out_u;         // (B,2,T)
  float* out_d;         // (B,T)
  int32_t* status;      // (B) or nullptr
  int32_t* iters;       // (B) or nullptr: incremented per executed iteration
  int32_t* active;      // (B) or nullptr: envs with 0 are skipped; cleared when the stop test fires
  // PAN.current_nom_values (pan.py:100-105), per env; nullptr disables the stop criterion
  float* prev_s; float* prev_u; float* prev_mu; float* prev_lam; int32_t* prev_count; int32_t* prev_valid;
  int B, T, M, E, kin;
  int max_ipm_iter;
  float iter_threshold;
  double dt, L;
  float q[3], p_u, eta, d_max, d_min;
  double ro, bk;
  double speed[2], acce[2];  // speed_bound, acce_bound = max_acce*dt (robot.py:68-69)
  float h[kMaxEdges];
};

__host__ __device__ inline size_t nrmp_warp_doubles(int T, int M) {
  const int nU = 2 * T, nR = nU - 2, TD = M > 0 ? T : 0, TM = T * M;
  const int m = 2 * nU + 2 * nR + 2 * TD + 2 * TM;
  size_t n = 0;
  n += 3 * (size_t)T * nU;        // F
  n += 3 * (size_t)T;             // s0
  n += (size_t)nU * (nU + 1) / 2; // Hc
  n += (size_t)nU * (nU + 1);     // H (padded rows)
  n += 2 * (size_t)T * nU;        // Gx, Gy
  n += 6 * (size_t)nU;            // c, x, dU, rhs, rdU, invd
  n += 18 * (size_t)T;            // per-step scalars
  n += 4 * (size_t)m;             // s, z, ds, dz
  n += 6 * (size_t)TM;            // fax, fay, kk, wrh, iHww, bw
  n += 11 * (size_t)T;            // linearisation a02,a12,B(6),C(3)
  return n + 8;
}

#define NB_LL(i, n) for (int i = lane; i < (n); i += 32)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256) nrmp_kernel(const NrmpParams prm, int warps_per_cta, int warp_doubles) {
  extern __shared__ __align__(16) double smem_d[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * warps_per_cta + warp;
  if (b >= prm.B) return;
  if (prm.active && prm.active[b] == 0) return;

  const int T = prm.T, M = prm.M, T1 = T + 1;
  const int nU = 2 * T, nR = nU - 2, TD = M > 0 ? T : 0, TM = T * M;
  const int oBU = 0, oBL = nU, oRU = 2 * nU, oRL = oRU + nR, oDU = oRL + nR, oDL = oDU + TD, oHW = oDL + TD, oHR = oHW + TM;
  const int m = oHR + TM;
  const int HS = nU + 1;  // row stride of H

  // ---- carve this warp's workspace -------------------------------------------------------
  double* wsp = smem_d + (size_t)warp * warp_doubles;
  double* F = wsp;            wsp += 3 * T * nU;   // F[r][t][j]
  double* s0 = wsp;           wsp += 3 * T;        // s0[r][t]
  double* Hc = wsp;           wsp += nU * (nU + 1) / 2;
  double* H = wsp;            wsp += nU * HS;
  double* Gx = wsp;           wsp += T * nU;
  double* Gy = wsp;           wsp += T * nU;
  double* cv = wsp;           wsp += nU;
  double* x = wsp;            wsp += nU;
  double* dU = wsp;           wsp += nU;
  double* rhs = wsp;          wsp += nU;
  double* rdU = wsp;          wsp += nU;
  double* invd = wsp;         wsp += nU;
  double* Dv = wsp;           wsp += T;
  double* dD = wsp;           wsp += T;
  double* rdD = wsp;          wsp += T;
  double* bD = wsp;           wsp += T;
  double* N00 = wsp;          wsp += T;
  double* N01 = wsp;          wsp += T;
  double* N11 = wsp;          wsp += T;
  double* n0 = wsp;           wsp += T;
  double* n1 = wsp;           wsp += T;
  double* iHDD = wsp;         wsp += T;
  double* e0 = wsp;           wsp += T;
  double* e1 = wsp;           wsp += T;
  double* qx = wsp;           wsp += T;
  double* qy = wsp;           wsp += T;
  double* gt0 = wsp;          wsp += T;
  double* gt1 = wsp;          wsp += T;
  double* gam_b = wsp;        wsp += T;
  wsp += T;  // spare
  double* cs = wsp;           wsp += m;   // slacks
  double* cz = wsp;           wsp += m;   // multipliers
  double* cds = wsp;          wsp += m;
  double* cdz = wsp;          wsp += m;   // holds rc on entry of a Newton solve, dz on exit
  double* fax = wsp;          wsp += TM;
  double* fay = wsp;          wsp += TM;
  double* kk = wsp;           wsp += TM;
  double* wrh = wsp;          wsp += TM;
  double* iHww = wsp;         wsp += TM;
  double* bw = wsp;           wsp += TM;
  double* a02 = wsp;          wsp += T;
  double* a12 = wsp;          wsp += T;
  double* Bm = wsp;           wsp += 6 * T;
  double* Cm = wsp;           wsp += 3 * T;

  const float* ns = prm.nom_s + (size_t)b * 3 * T1;
  const float* nu = prm.nom_u + (size_t)b * 2 * T;
  const float* rs = prm.ref_s + (size_t)b * 3 * T1;
  const float* rus = prm.ref_us + (size_t)b * T;
  const bool omni = prm.kin == 2;
  const int rows_s = omni ? 2 : 3;
  const bool en_speed[2] = {isfinite(prm.speed[0]), isfinite(prm.speed[1])};
  const bool en_acce[2] = {isfinite(prm.acce[0]), isfinite(prm.acce[1])};
  auto enabled = [&](int k) -> bool {
    if (k < oRU) return en_speed[(k % nU) & 1];
    if (k < oDU) return en_acce[((k - oRU) % nR) & 1];
    return true;
  };
  int m_active = 2 * TD + 2 * TM;
  for (int c = 0; c < 2; ++c) m_active += (en_speed[c] ? 2 * T : 0) + (en_acce[c] ? 2 * (T - 1) : 0);

  // ---- 1. kinematics linearisation (robot.py:272-316, float32 tensor semantics) -------------
  const float fdt = (float)prm.dt;
  NB_LL(t, T) {
    const float v = nu[t];
    const float phi = (prm.kin == 2) ? nu[T + t] : ns[2 * T1 + t];
    const double snd = sin((double)phi), csd = cos((double)phi);
    const float snf = (float)snd, csf = (float)csd;
    a02[t] = omni ? 0.0 : (double)__fmul_rn(__fmul_rn(-v, fdt), snf);
    a12[t] = omni ? 0.0 : (double)__fmul_rn(__fmul_rn(v, fdt), csf);
    Cm[3 * t + 0] = (double)__fmul_rn(__fmul_rn(__fmul_rn(phi, v), snf), fdt);
    Cm[3 * t + 1] = (double)__fmul_rn(__fmul_rn(__fmul_rn(-phi, v), csf), fdt);
    Cm[3 * t + 2] = 0.0;
    double* Bt = Bm + 6 * t;  // row-major 3x2
    Bt[0] = (double)(float)(csd * prm.dt); Bt[1] = 0.0;
    Bt[2] = (double)(float)(snd * prm.dt); Bt[3] = 0.0;
    Bt[4] = 0.0; Bt[5] = 0.0;
    if (prm.kin == 0) {
      Bt[5] = (double)fdt;
    } else if (prm.kin == 1) {
      const float psi = nu[T + t];
      const double cp = cos((double)psi);
      const float den = (float)(prm.L * (cp * cp));
      Bt[4] = (double)(float)(tan((double)psi) * prm.dt / prm.L);
      Bt[5] = (double)__fdiv_rn(__fmul_rn(v, fdt), den);
      Cm[3 * t + 2] = (double)__fdiv_rn(__fmul_rn(__fmul_rn(-psi, v), fdt), den);
    } else {
      Bt[1] = (double)__fmul_rn(__fmul_rn(-v, snf), fdt);
      Bt[3] = (double)__fmul_rn(__fmul_rn(v, csf), fdt);
    }
    gam_b[t] = (double)__fmul_rn(prm.p_u, rus[t]);  // p_u * ref_us in float32 (nrmp.py:158)
  }
  __syncwarp();

  // ---- 2. free response s0 and sensitivities F (condensing) ---------------------------------
  if (lane == 0) {
    double px = ns[0], py = ns[T1], pth = ns[2 * T1];
    for (int t = 0; t < T; ++t) {
      const double nx = px + a02[t] * pth + Cm[3 * t], ny = py + a12[t] * pth + Cm[3 * t + 1], nth = pth + Cm[3 * t + 2];
      s0[t] = nx; s0[T + t] = ny; s0[2 * T + t] = nth;
      px = nx; py = ny; pth = nth;
    }
  }
  NB_LL(j, nU) {
    const int tj = j >> 1, cj = j & 1;
    double fx = 0, fy = 0, fth = 0;
    for (int t = 0; t < T; ++t) {
      if (t == tj) {
        fx = Bm[6 * t + cj]; fy = Bm[6 * t + 2 + cj]; fth = Bm[6 * t + 4 + cj];
      } else if (t > tj) {
        fx += a02[t] * fth; fy += a12[t] * fth;
      }
      F[(0 * T + t) * nU + j] = fx; F[(1 * T + t) * nU + j] = fy; F[(2 * T + t) * nU + j] = fth;
    }
  }
  __syncwarp();

  // ---- 3. quadratic cost 0.5 u^T Hc u + c^T u ------------------------------------------------
  double qd[3], qq[3];
  for (int r = 0; r < 3; ++r) {
    qq[r] = r < rows_s ? (double)prm.q[r] : 0.0;
    qd[r] = 2.0 * qq[r] * qq[r] + prm.bk;
  }
  const double pu = (double)prm.p_u;
  NB_LL(j, nU) {
    double acc = 0;
    for (int t = j >> 1; t < T; ++t)
      for (int r = 0; r < 3; ++r) {
        const double ga = (double)__fmul_rn(prm.q[r], rs[r * T1 + t + 1]);  // q_s * ref_s in float32
        const double g = 2.0 * qq[r] * ga + prm.bk * (double)ns[r * T1 + t + 1];
        acc += F[(r * T + t) * nU + j] * (qd[r] * s0[r * T + t] - g);
      }
    if ((j & 1) == 0) acc += -2.0 * pu * gam_b[j >> 1];
    cv[j] = acc;
  }
  for (int p = lane; p < nU * (nU + 1) / 2; p += 32) {
    int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while (i * (i + 1) / 2 > p) --i;
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    const int j = p - i * (i + 1) / 2;
    double acc = 0;
    for (int t = i >> 1; t < T; ++t)
      for (int r = 0; r < 3; ++r) acc += qd[r] * F[(r * T + t) * nU + i] * F[(r * T + t) * nU + j];
    if (i == j && (i & 1) == 0) acc += 2.0 * pu * pu;
    Hc[p] = acc;
  }

  // ---- 4. obstacle coefficients (nrmp.py:220-261) ---------------------------------------------
  const int cnt = (prm.fa || M == 0) ? M : (prm.sel_count ? prm.sel_count[b] : 0);
  NB_LL(k, TM) {
    const int t = k / M, mm = k - t * M;
    float fx = 0.f, fy = 0.f, fbv = 0.f;
    if (prm.fa) {
      fx = prm.fa[((size_t)b * TM + k) * 2]; fy = prm.fa[((size_t)b * TM + k) * 2 + 1]; fbv = prm.fb[(size_t)b * TM + k];
    } else if (cnt > 0) {
      const int src = mm < cnt ? mm : 0;  // rows pn..M copy row 0 (nrmp.py:258-259)
      const size_t o = ((size_t)b * T1 + (t + 1)) * M + src;  // list entry t+1 (nrmp.py:244)
      fx = prm.sel_lam[o * 2]; fy = prm.sel_lam[o * 2 + 1];
      const float tmp = __fadd_rn(__fmul_rn(fx, prm.sel_pts[o * 2]), __fmul_rn(fy, prm.sel_pts[o * 2 + 1]));
      float muh = 0.f;
      for (int e = 0; e < prm.E; ++e) muh = fmaf(prm.sel_mu[o * prm.E + e], prm.h[e], muh);
      fbv = __fadd_rn(tmp, muh);
    }
    fax[k] = fx; fay[k] = fy;
    kk[k] = (double)fbv;  // completed with -fa.s0 below
  }
  __syncwarp();
  NB_LL(k, TM) {
    const int t = k / M;
    kk[k] -= fax[k] * s0[t] + fay[k] * s0[T + t];
  }

  // ---- 5. strictly feasible start -------------------------------------------------------------
  const double rho = prm.ro, eta = (double)prm.eta;
  const double dlo = fmax((double)prm.d_min, 0.0), dhi = (double)prm.d_max;
  int stat = 0;
  if (TD > 0 && !(dhi > dlo)) stat |= 4;
  for (int c = 0; c < 2; ++c) {
    if (en_speed[c] && !(prm.speed[c] > 0)) stat |= 4;
    if (en_acce[c] && !(prm.acce[c] > 0)) stat |= 4;
  }
  NB_LL(i, nU) {
    x[i] = 0.0;
    const int c = i & 1;
    cs[oBU + i] = en_speed[c] ? prm.speed[c] : 1.0; cs[oBL + i] = cs[oBU + i];
    cz[oBU + i] = en_speed[c] ? 1.0 / cs[oBU + i] : 0.0; cz[oBL + i] = cz[oBU + i];
    if (i < nR) {
      cs[oRU + i] = en_acce[c] ? prm.acce[c] : 1.0; cs[oRL + i] = cs[oRU + i];
      cz[oRU + i] = en_acce[c] ? 1.0 / cs[oRU + i] : 0.0; cz[oRL + i] = cz[oRU + i];
    }
  }
  NB_LL(t, TD) {
    Dv[t] = 0.5 * (dlo + dhi);
    cs[oDU + t] = dhi - Dv[t]; cs[oDL + t] = Dv[t] - dlo;
    cz[oDU + t] = 1.0 / cs[oDU + t]; cz[oDL + t] = 1.0 / cs[oDL + t];
  }
  NB_LL(k, TM) {
    const double r = 0.5 * (dlo + dhi) + kk[k];
    const double w = fmax(r, 0.0) + 1.0;
    cs[oHW + k] = w; cs[oHR + k] = w - r;
    cz[oHW + k] = 1.0 / w; cz[oHR + k] = 1.0 / (w - r);
  }
  __syncwarp();

  // One Newton solve with the factorised H: reads rc from cdz, writes dU, dD, cds, cdz.
  auto newton = [&]() {
    NB_LL(k, TM) {
      const double vw = cdz[oHW + k] / cs[oHW + k], vr = cdz[oHR + k] / cs[oHR + k];
      const double rdw = rho * cs[oHW + k] - cz[oHW + k] - cz[oHR + k];
      const double b_w = -rdw + vw + vr;
      bw[k] = b_w;
    }
    __syncwarp();
    NB_LL(t, TD) {
      double Y0 = 0, Y1 = 0, Ys = 0;
      for (int mm = 0; mm < M; ++mm) {
        const int k = t * M + mm;
        const double y = wrh[k] * bw[k] - cdz[oHR + k] / cs[oHR + k];
        Y0 += y * fax[k]; Y1 += y * fay[k]; Ys += y;
      }
      const double vdu = cdz[oDU + t] / cs[oDU + t], vdl = cdz[oDL + t] / cs[oDL + t];
      const double b = -rdD[t] - (vdu - vdl) + Ys;
      bD[t] = b;
      e0[t] = -Y0 + n0[t] * b * iHDD[t];
      e1[t] = -Y1 + n1[t] * b * iHDD[t];
    }
    __syncwarp();
    NB_LL(i, nU) {
      double acc = -rdU[i] - (cdz[oBU + i] / cs[oBU + i] - cdz[oBL + i] / cs[oBL + i]);
      if (i >= 2) acc -= cdz[oRU + i - 2] / cs[oRU + i - 2] - cdz[oRL + i - 2] / cs[oRL + i - 2];
      if (i < nR) acc += cdz[oRU + i] / cs[oRU + i] - cdz[oRL + i] / cs[oRL + i];
      if (TD > 0)
        for (int t = i >> 1; t < T; ++t) acc += F[(0 * T + t) * nU + i] * e0[t] + F[(1 * T + t) * nU + i] * e1[t];
      rhs[i] = acc;
    }
    __syncwarp();
    // forward / backward substitution with the Cholesky factor in H (lower), invd = 1/diag
    for (int k = 0; k < nU; ++k) {
      const double yk = rhs[k] * invd[k];
      __syncwarp();
      NB_LL(i, nU) {
        if (i > k) rhs[i] -= H[i * HS + k] * yk;
        else if (i == k) rhs[i] = yk;
      }
      __syncwarp();
    }
    for (int k = nU - 1; k >= 0; --k) {
      const double yk = rhs[k] * invd[k];
      __syncwarp();
      NB_LL(i, nU) {
        if (i < k) rhs[i] -= H[k * HS + i] * yk;
        else if (i == k) rhs[i] = yk;
      }
      __syncwarp();
    }
    NB_LL(i, nU) dU[i] = rhs[i];
    __syncwarp();
    NB_LL(t, TD) {
      double ax = 0, ay = 0;
      for (int i = 0; i < 2 * (t + 1); ++i) {
        ax += F[(0 * T + t) * nU + i] * dU[i];
        ay += F[(1 * T + t) * nU + i] * dU[i];
      }
      qx[t] = ax; qy[t] = ay;
      dD[t] = (bD[t] + n0[t] * ax + n1[t] * ay) * iHDD[t];
    }
    __syncwarp();
    NB_LL(k, TM) {
      const int t = k / M;
      const double Jdx = dD[t] - (fax[k] * qx[t] + fay[k] * qy[t]);
      const double dw = bw[k] * iHww[k] + wrh[k] * Jdx;
      const double dsw = dw, dsr = dw - Jdx;
      cdz[oHW + k] = (cdz[oHW + k] - cz[oHW + k] * dsw) / cs[oHW + k];
      cdz[oHR + k] = (cdz[oHR + k] - cz[oHR + k] * dsr) / cs[oHR + k];
      cds[oHW + k] = dsw; cds[oHR + k] = dsr;
    }
    NB_LL(i, nU) {
      const double d = dU[i];
      cdz[oBU + i] = (cdz[oBU + i] + cz[oBU + i] * d) / cs[oBU + i]; cds[oBU + i] = -d;
      cdz[oBL + i] = (cdz[oBL + i] - cz[oBL + i] * d) / cs[oBL + i]; cds[oBL + i] = d;
      if (i < nR) {
        const double dd = dU[i + 2] - d;
        cdz[oRU + i] = (cdz[oRU + i] + cz[oRU + i] * dd) / cs[oRU + i]; cds[oRU + i] = -dd;
        cdz[oRL + i] = (cdz[oRL + i] - cz[oRL + i] * dd) / cs[oRL + i]; cds[oRL + i] = dd;
      }
    }
    NB_LL(t, TD) {
      const double d = dD[t];
      cdz[oDU + t] = (cdz[oDU + t] + cz[oDU + t] * d) / cs[oDU + t]; cds[oDU + t] = -d;
      cdz[oDL + t] = (cdz[oDL + t] - cz[oDL + t] * d) / cs[oDL + t]; cds[oDL + t] = d;
    }
    __syncwarp();
  };

  auto max_step = [&]() -> double {
    double a = 1.0;
    NB_LL(k, m) {
      if (!enabled(k)) continue;
      const double ds = cds[k], dz = cdz[k];
      if (ds < 0) a = fmin(a, -cs[k] / ds);
      if (dz < 0) a = fmin(a, -cz[k] / dz);
    }
    return warp_min(a);
  };

  // ---- 6. interior point iterations -------------------------------------------------------------
  int it = 0;
  bool converged = false;
  if (stat == 0) {
    for (it = 0; it < prm.max_ipm_iter; ++it) {
      // dual residual
      NB_LL(t, TD) {
        double gx = 0, gy = 0, sz = 0;
        for (int mm = 0; mm < M; ++mm) {
          const int k = t * M + mm;
          const double zr = cz[oHR + k];
          gx += zr * fax[k]; gy += zr * fay[k]; sz += zr;
        }
        gt0[t] = gx; gt1[t] = gy;
        rdD[t] = -eta + cz[oDU + t] - cz[oDL + t] + sz;
      }
      __syncwarp();
      double res = 0.0;
      NB_LL(i, nU) {
        double acc = cv[i];
        for (int j = 0; j < nU; ++j) {
          const int hi = i > j ? i : j, lo = i > j ? j : i;
          acc += Hc[hi * (hi + 1) / 2 + lo] * x[j];
        }
        acc += cz[oBU + i] - cz[oBL + i];
        if (i >= 2) acc += cz[oRU + i - 2] - cz[oRL + i - 2];
        if (i < nR) acc -= cz[oRU + i] - cz[oRL + i];
        if (TD > 0)
          for (int t = i >> 1; t < T; ++t) acc -= F[(0 * T + t) * nU + i] * gt0[t] + F[(1 * T + t) * nU + i] * gt1[t];
        rdU[i] = acc;
        res = fmax(res, fabs(acc));
      }
      NB_LL(t, TD) res = fmax(res, fabs(rdD[t]));
      double gsum = 0.0;
      NB_LL(k, m) if (enabled(k)) gsum += cs[k] * cz[k];
      NB_LL(k, TM) res = fmax(res, fabs(rho * cs[oHW + k] - cz[oHW + k] - cz[oHR + k]));
      res = warp_max(res);
      const double gap = warp_sum(gsum) / (double)(m_active > 0 ? m_active : 1);
      if (!(gap == gap) || !(res == res)) { stat |= 2; break; }
      if (gap < 1e-10 && res < 1e-8) { converged = true; break; }

      // barrier weights, per-step reductions, D elimination
      NB_LL(t, TD) {
        double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0, nn = 0;
        for (int mm = 0; mm < M; ++mm) {
          const int k = t * M + mm;
          const double Ww = cz[oHW + k] / cs[oHW + k], Wr = cz[oHR + k] / cs[oHR + k];
          const double ih = 1.0 / (rho + Ww + Wr);
          const double om = Wr * (rho + Ww) * ih;
          wrh[k] = Wr * ih; iHww[k] = ih;
          const double fx = fax[k], fy = fay[k];
          a00 += om * fx * fx; a01 += om * fx * fy; a11 += om * fy * fy; b0 += om * fx; b1 += om * fy; nn += om;
        }
        const double hdd = nn + cz[oDU + t] / cs[oDU + t] + cz[oDL + t] / cs[oDL + t];
        const double ih = 1.0 / hdd;
        iHDD[t] = ih; n0[t] = b0; n1[t] = b1;
        N00[t] = a00 - b0 * b0 * ih; N01[t] = a01 - b0 * b1 * ih; N11[t] = a11 - b1 * b1 * ih;
      }
      __syncwarp();
      if (TD > 0) {
        NB_LL(j, nU)
          for (int t = j >> 1; t < T; ++t) {
            const double fx = F[(0 * T + t) * nU + j], fy = F[(1 * T + t) * nU + j];
            Gx[t * nU + j] = N00[t] * fx + N01[t] * fy;
            Gy[t * nU + j] = N01[t] * fx + N11[t] * fy;
          }
        __syncwarp();
      }
      // assemble the reduced Hessian (lower triangle)
      for (int p = lane; p < nU * (nU + 1) / 2; p += 32) {
        int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > p) --i;
        while ((i + 1) * (i + 2) / 2 <= p) ++i;
        const int j = p - i * (i + 1) / 2;
        double acc = Hc[p];
        if (TD > 0)
          for (int t = i >> 1; t < T; ++t) acc += F[(0 * T + t) * nU + i] * Gx[t * nU + j] + F[(1 * T + t) * nU + i] * Gy[t * nU + j];
        if (i == j) {
          acc += cz[oBU + i] / cs[oBU + i] + cz[oBL + i] / cs[oBL + i];
          if (i >= 2) acc += cz[oRU + i - 2] / cs[oRU + i - 2] + cz[oRL + i - 2] / cs[oRL + i - 2];
          if (i < nR) acc += cz[oRU + i] / cs[oRU + i] + cz[oRL + i] / cs[oRL + i];
        } else if (i == j + 2) {
          acc -= cz[oRU + j] / cs[oRU + j] + cz[oRL + j] / cs[oRL + j];
        }
        H[i * HS + j] = acc;
      }
      __syncwarp();
      // Cholesky, left-looking by column; lanes own rows
      bool bad = false;
      for (int k = 0; k < nU; ++k) {
        NB_LL(i, nU) {
          if (i >= k) {
            double acc = H[i * HS + k];
            for (int p = 0; p < k; ++p) acc -= H[i * HS + p] * H[k * HS + p];
            H[i * HS + k] = acc;
          }
        }
        __syncwarp();
        const double d = H[k * HS + k];
        __syncwarp();
        if (!(d > 0.0)) { bad = true; break; }
        const double ild = rsqrt(d);
        NB_LL(i, nU) {
          if (i > k) H[i * HS + k] *= ild;
          else if (i == k) { H[i * HS + k] = d * ild; invd[k] = ild; }
        }
        __syncwarp();
      }
      if (bad) { stat |= 2; break; }

      // predictor
      NB_LL(k, m) cdz[k] = enabled(k) ? -cs[k] * cz[k] : 0.0;
      __syncwarp();
      newton();
      const double a_aff = max_step();
      double ga = 0.0;
      NB_LL(k, m) if (enabled(k)) ga += (cs[k] + a_aff * cds[k]) * (cz[k] + a_aff * cdz[k]);
      const double gap_aff = warp_sum(ga) / (double)(m_active > 0 ? m_active : 1);
      const double sr = gap_aff / gap;
      const double sigma = sr * sr * sr;
      // corrector
      NB_LL(k, m) cdz[k] = enabled(k) ? (-cs[k] * cz[k] + sigma * gap - cds[k] * cdz[k]) : 0.0;
      __syncwarp();
      newton();
      double a = max_step();
      a = fmin(1.0, 0.995 * a);
      NB_LL(i, nU) x[i] += a * dU[i];
      NB_LL(t, TD) Dv[t] += a * dD[t];
      NB_LL(k, m) if (enabled(k)) { cs[k] += a * cds[k]; cz[k] += a * cdz[k]; }
      __syncwarp();
    }
    if (!converged && !(stat & 2)) stat |= 1;
  }

  // ---- 7. outputs: S = s0 + F u, U, D cast to float32 (nrmp.py:145-148) --------------------------
  float* os = prm.out_s + (size_t)b * 3 * T1;
  float* ou = prm.out_u + (size_t)b * 2 * T;
  float* od = prm.out_d + (size_t)b * T;
  const bool keep_nominal = (stat & 6) != 0;  // numeric failure / infeasible: hand back the nominal
  __syncwarp();
  const float init_s = lane < 3 ? ns[lane * T1] : 0.f;  // initial state column (robot.py:234)
  NB_LL(t, T) {
    double sx = s0[t], sy = s0[T + t], sth = s0[2 * T + t];
    for (int i = 0; i < 2 * (t + 1); ++i) {
      sx += F[(0 * T + t) * nU + i] * x[i]; sy += F[(1 * T + t) * nU + i] * x[i]; sth += F[(2 * T + t) * nU + i] * x[i];
    }
    if (keep_nominal) { sx = ns[t + 1]; sy = ns[T1 + t + 1]; sth = ns[2 * T1 + t + 1]; }
    // stash in Gx (no longer needed) so that reads of nom_s above are complete before any write
    Gx[3 * t] = sx; Gx[3 * t + 1] = sy; Gx[3 * t + 2] = sth;
  }
  NB_LL(i, nU) Gy[i] = keep_nominal ? (double)nu[(i & 1) * T + (i >> 1)] : x[i];
  __syncwarp();
  if (lane < 3) os[lane * T1] = init_s;
  NB_LL(t, T) {
    os[t + 1] = (float)Gx[3 * t]; os[T1 + t + 1] = (float)Gx[3 * t + 1]; os[2 * T1 + t + 1] = (float)Gx[3 * t + 2];
    od[t] = TD > 0 ? (float)Dv[t] : 0.f;
  }
  NB_LL(i, nU) ou[(i & 1) * T + (i >> 1)] = (float)Gy[i];
  if (lane == 0) {
    if (prm.status) prm.status[b] = stat;
    if (prm.iters) prm.iters[b] += 1;
  }
  __syncwarp();

  // ---- 8. stop criterion (pan.py:215-243) ----------------------------------------------------------
  if (prm.prev_valid) {
    const int E = prm.E;
    const int cur_cnt = (prm.fa || M == 0 || !prm.sel_count) ? 0 : prm.sel_count[b];
    const int valid = prm.prev_valid[b];
    const int pcnt = prm.prev_count[b];
    float* ps = prm.prev_s + (size_t)b * 3 * T1;
    float* pu_ = prm.prev_u + (size_t)b * 2 * T;
    float* pmu = prm.prev_mu ? prm.prev_mu + (size_t)b * T1 * M * E : nullptr;
    float* plam = prm.prev_lam ? prm.prev_lam + (size_t)b * T1 * M * 2 : nullptr;
    const float* cmu = prm.sel_mu ? prm.sel_mu + (size_t)b * T1 * M * E : nullptr;
    const float* clam = prm.sel_lam ? prm.sel_lam + (size_t)b * T1 * M * 2 : nullptr;
    float diff = 0.f;
    if (valid) {
      if (cur_cnt == 0 || pcnt == 0) {
        double a1 = 0, a2 = 0;
        NB_LL(i, 3 * T1) { const double d = (double)os[i] - (double)ps[i]; a1 += d * d; }
        NB_LL(i, 2 * T) { const double d = (double)ou[i] - (double)pu_[i]; a2 += d * d; }
        const float n1f = sqrtf((float)warp_sum(a1)), n2f = sqrtf((float)warp_sum(a2));
        diff = n1f * n1f + n2f * n2f;
      } else {
        const int en = cur_cnt < pcnt ? cur_cnt : pcnt;
        double a1 = 0, a2 = 0;
        NB_LL(i, T1 * en * E) {
          const int tt = i / (en * E), rem = i - tt * en * E;  // rem = col*E + e, col < en
          const size_t o = (size_t)tt * M * E + rem;
          const double d = (double)cmu[o] - (double)pmu[o];
          a1 += d * d;
        }
        NB_LL(i, T1 * en * 2) {
          const int tt = i / (en * 2), rem = i - tt * en * 2;
          const size_t o = (size_t)tt * M * 2 + rem;
          const double d = (double)clam[o] - (double)plam[o];
          a2 += d * d;
        }
        const float md = sqrtf((float)warp_sum(a1)) / (float)en, ld = sqrtf((float)warp_sum(a2)) / (float)en;
        diff = md * md + ld * ld;
      }
    }
    __syncwarp();
    NB_LL(i, 3 * T1) ps[i] = os[i];
    NB_LL(i, 2 * T) pu_[i] = ou[i];
    if (pmu && cmu) {
      NB_LL(i, T1 * M * E) pmu[i] = cmu[i];
      NB_LL(i, T1 * M * 2) plam[i] = clam[i];
    }
    if (lane == 0) {
      prm.prev_valid[b] = 1;
      prm.prev_count[b] = cur_cnt;
      if (valid && diff < prm.iter_threshold && prm.active) prm.active[b] = 0;
    }
  }
}

#undef NB_LL
}  // namespace nb
