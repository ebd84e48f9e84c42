// Warp-per-band PVQ band quantiser (reference pvq_theta, src/pvq_encoder.c:333, with the closed-form
// rate of od_pvq_rate, :247, and pvq_search_rdo_double, :93).
//
// ONE warp owns ONE band.  Design goals, in this order: (1) the latency of a single band is what the
// keyframe luma intra wavefront is bound by (od_hv_intra_pred couples a band to the same band of the
// top / left neighbour, src/intra.c:37), so everything that the reference does serially per band but is
// independent per CANDIDATE runs one candidate per lane; (2) the greedy pulse search costs one
// redux.sync per pulse instead of a five-level shuffle tree.
//
//  * Element j of the band lives in lane j % 32, slot j / 32 (E = 1 for n <= 32, E = 4 for n = 128),
//    in registers.
//  * Every sum the reference accumulates in double is a sum of exactly representable integers far below
//    2^53, hence order independent: reduced as integers with redux.sync (__reduce_add_sync).
//  * Candidates (gain i, theta j) of pvq_theta: lane c < 12 holds with-reference candidate
//    (gain index c / 4, theta index c % 4) in the reference's insertion order, lanes 12 / 13 the
//    no-reference gains.  Their quantised angle, K, distortion pre-test, rate constants are computed
//    once, in parallel.  The reference's stable sort by K and its `prev_k` reuse of search results are
//    reproduced by processing "events" = runs of surviving candidates with equal K in ascending K order:
//    one (incremental) search per event, then the costs of the event's candidates in parallel, then the
//    reference's sequential `<` / `<=` fold over them.
//  * Per-pulse arg-max.  Plain pulses maximise (xy + x_j)^2 / (yy + 2 y_j + 1): every lane builds an
//    f32 approximation of the ratio (relative error < 2^-20), redux.max picks the approximate maximum
//    and the elements within 64 ulps of it are the contenders.  The exact maximum is always a
//    contender; if there is exactly one it is the reference's choice.  Otherwise (ties, near ties, or
//    products that could exceed 2^53, where the reference's own comparisons round) the contenders are
//    compared with the reference's literal double-precision test in index order.  RDO pulses maximise
//    a double; its f32 rounding is monotone, so the contenders are the elements whose rounded value
//    equals the maximum.
//
// The function compiles for the host under DAALA_B200_EMU (tests/emu/simt_emu.h supplies the warp
// primitives on 32 fibres) so that CPU tests can pin it against the reference build.
#pragma once
#include <math.h>
#include <stdint.h>

#include "pvq_math.cuh"

namespace daala_b200 {
namespace pvq {

constexpr unsigned kFull = 0xffffffffu;

// coverage counters of the host emulation build (tests/emu); nothing on the device
// below this bound every product of the plain-pulse ratio test is exact in double (tests lower it to
// force the literal-scan path)
#ifndef DAALA_B200_PVQ_EXACT_BOUND
#define DAALA_B200_PVQ_EXACT_BOUND 4503599627370496.
#endif
#ifdef DAALA_B200_EMU_STATS
#define PVQ_WARP_STAT(i) (daala_b200_pvq_warp_stats[i]++)
#else
#define PVQ_WARP_STAT(i) ((void)0)
#endif

__device__ __forceinline__ int wsum(int v) { return __reduce_add_sync(kFull, v); }
__device__ __forceinline__ int wmax(int v) { return __reduce_max_sync(kFull, v); }
__device__ __forceinline__ int wmin(int v) { return __reduce_min_sync(kFull, v); }
__device__ __forceinline__ unsigned wmaxu(unsigned v) { return __reduce_max_sync(kFull, v); }
// exact 64-bit sum of per-lane values |s| < 2^55
__device__ __forceinline__ long long wsum64(long long s) {
  const int lo = (int)(s & 0xffffff);
  const int hi = (int)(s >> 24);
  return ((long long)wsum(hi) << 24) + wsum(lo);
}
__device__ __forceinline__ double wbcast(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(kFull, lo, src);
  hi = __shfl_sync(kFull, hi, src);
  return __hiloint2double(hi, lo);
}
// monotone map of a float onto signed integers
__device__ __forceinline__ int ordered_key(float f) {
  const int b = __float_as_int(f);
  return b ^ ((b >> 31) & 0x7fffffff);
}

// a[slot] of a register array without dynamic indexing
template <int E>
__device__ __forceinline__ int pick(const int (&a)[E], int slot) {
  int v = a[0];
#pragma unroll
  for (int e = 1; e < E; e++) if (slot == e) v = a[e];
  return v;
}

// od_apply_householder (src/pvq.c:560) on distributed int16 vectors; in place allowed.
template <int E>
__device__ __forceinline__ void householder_apply_warp(int lane, int (&out)[E], const int (&x)[E], const int (&r)[E], int n) {
  int32_t l2r = 0, proj = 0;
#pragma unroll
  for (int e = 0; e < E; e++) {
    if (e * 32 + lane < n) {
      l2r += mul16(r[e], r[e]);
      proj += mul16(r[e], x[e]);
    }
  }
  l2r = wsum(l2r);
  proj = wsum(proj);
  int l2r_shift = (ilog((uint32_t)l2r) - 1) - 14;
  int16_t l2r_norm = (int16_t)vshr_round(l2r, l2r_shift);
  int16_t rcp = rcp16(l2r_norm);
  int proj_shift = (ilog((uint32_t)abs(proj)) - 1) - 14;
  int16_t proj_norm = (int16_t)vshr_round(proj, proj_shift);
  int16_t proj_1 = (int16_t)mul16_q15(proj_norm, rcp);
  int outshift = 14 - proj_shift - 1 + l2r_shift;
  if (outshift > 30) outshift = 30;
#pragma unroll
  for (int e = 0; e < E; e++) {
    int32_t t = mul16(r[e], proj_1);
    t = outshift >= 0 ? shr_round(t, outshift) : shl(t, -outshift);
    out[e] = (int16_t)(x[e] - t);
  }
}

// One band by one warp.  x0 / r0 / out / yout / qm / qm_inv point at the band's first entry.  Scalar
// results are identical in every lane.  int16 quantities of the reference are kept sign-extended in ints.
template <int E>
__device__ __forceinline__ int quantise_band_warp(int lane, int32_t* out, const int32_t* x0, const int32_t* r0, int n,
                                                  int q0, int32_t* yout, int* itheta, int* max_theta, int* vk, int beta,
                                                  double* skip_term, int is_keyframe, int pli, const int16_t* qm,
                                                  const int16_t* qm_inv, double pvq_norm_lambda) {
  const double gain_weight = 1.4;
  const double cgain_1 = 1. / kCgainOne;
  const double cgain_2 = cgain_1 * cgain_1;
  const double theta_scale = (1 << kThetaShift) * 2. / M_PI;
  const double theta_scale_1 = 1. / theta_scale;
  const double trig_1 = 1. / 32768;
  int x16[E], r16[E], xr[E];
  int32_t sx = 0, sr = 0;
  int r_nonnull = 0;
  {
    int32_t xv[E], rv[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int j = e * 32 + lane;
      const bool valid = j < n;
      xv[e] = valid ? x0[j] : 0;
      rv[e] = valid ? r0[j] : 0;
      const int16_t tx = (int16_t)(xv[e] >> 8), tr = (int16_t)(rv[e] >> 8);
      sx += tx * (int32_t)tx;
      sr += tr * (int32_t)tr;
      r_nonnull |= rv[e] != 0;
    }
    sx = wsum(sx);
    sr = wsum(sr);
    r_nonnull = __any_sync(kFull, r_nonnull);
    int xshift = 9 + ilog((uint32_t)(n + sx)) / 2 - 15;
    int rshift = 9 + ilog((uint32_t)(n + sr)) / 2 - 14;
    if (xshift < 0) xshift = 0;
    if (rshift < 0) rshift = 0;
    // from here on xshift / rshift live in sx / sr's place
    sx = xshift;
    sr = rshift;
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int j = e * 32 + lane;
      const int qmv = j < n ? qm[j] : 0;
      x16[e] = (int16_t)shr_round(xv[e] * qmv, kQmShift + xshift);
      r16[e] = (int16_t)shr_round(rv[e] * qmv, kQmShift + rshift);
    }
  }
  const int xshift = sx, rshift = sr;
  long long scorr = 0;
  int32_t accx = 0, accr = 0;
#pragma unroll
  for (int e = 0; e < E; e++) {
    scorr += mul16(x16[e], r16[e]);
    accx += x16[e] * x16[e];
    accr += r16[e] * r16[e];
  }
  double corr = (double)wsum64(scorr);
  accx = wsum(accx);
  accr = wsum(accr);
  int32_t g, gr;
  const int cfl_enabled = is_keyframe && pli != 0;
  const int32_t cg = compute_gain_from_energy(accx, q0, &g, beta, xshift);
  int32_t cgr = compute_gain_from_energy(accr, q0, &gr, beta, rshift);
  if (cfl_enabled) cgr = kCgainOne;
  const int icgr = shr_round(cgr, kCgainShift);
  int32_t gain_offset = cgr - shl(icgr, kCgainShift);
  double best_dist = gain_weight * cg * cg * cgain_2;
  double best_cost = best_dist + pvq_norm_lambda * 0.;  // od_pvq_rate(0, 0, -1, 0, ...) == 0
  corr = corr / (1e-100 + g * (double)gr / shl(1, xshift + rshift));
  corr = corr < 1. ? corr : 1.;
  corr = corr > -1. ? corr : -1.;
  double skip_dist;
  if (is_keyframe) {
    skip_dist = gain_weight * cg * cg * cgain_2;
  } else {
    skip_dist = gain_weight * (cg - cgr) * (cg - cgr) + cgr * (double)cg * (2 - 2 * corr);
    skip_dist *= cgain_2;
  }
  if (!is_keyframe) {
    const int32_t scgr = gain_offset > 0 ? gain_offset : 0;
    if (icgr == 0) {
      best_dist = gain_weight * (cg - scgr) * (cg - scgr) + scgr * (double)cg * (2 - 2 * corr);
      best_dist *= cgain_2;
    }
    best_cost = best_dist + pvq_norm_lambda * 0.;  // od_pvq_rate(0, icgr, 0, 0, ...) == 0
  }
  const double dist0 = best_dist;

  // ---- with-reference setup: angle, Householder reflection (uniform) ------------------------------------
  const bool have_ref = r_nonnull && corr > 0;
  int32_t theta = 0;
  int m = 0, s = 1;
  if (have_ref) {
    theta = round32(theta_scale * acos(corr));
    // od_compute_householder, src/pvq.c:498: first largest |r| (strict ">" from maxr = 0)
    int key = -1;
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int j = e * 32 + lane;
      const int a = j < n ? abs(r16[e]) : 0;
      const int kk = j < n ? (a << 7) | (127 - j) : -1;
      key = kk > key ? kk : key;
    }
    key = wmax(key);
    m = (key >> 7) > 0 ? 127 - (key & 127) : 0;
    const int rm = __shfl_sync(kFull, pick<E>(r16, m >> 5), m & 31);
    s = rm > 0 ? 1 : -1;
#pragma unroll
    for (int e = 0; e < E; e++)
      if (e * 32 + lane == m) r16[e] = (int16_t)(r16[e] + shr_round(gr * s, rshift));
    householder_apply_warp<E>(lane, xr, x16, r16, n);
    // drop element m: xr[j] <- xr[j + 1] for j >= m
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int t1 = __shfl_sync(kFull, xr[e], (lane + 1) & 31);
      const int t2 = e + 1 < E ? __shfl_sync(kFull, xr[e + 1 < E ? e + 1 : e], 0) : 0;
      const int nxt = lane == 31 ? t2 : t1;
      if (e * 32 + lane >= m) xr[e] = nxt;
    }
  } else {
#pragma unroll
    for (int e = 0; e < E; e++) xr[e] = 0;
  }

  // ---- candidates: one per lane -------------------------------------------------------------------------
  int c_gain = 0, c_theta = -1, c_ts = 0, c_k = 0;
  int32_t c_qcg = 0, c_qtheta = 0;
  bool c_alive = false;
  double c_sinprod = 0, c_g2 = 0, c_rate_ts = 0, c_dist = 0;
  if (lane < 12) {
    if (have_ref) {
      const int gain_bound = (cg - gain_offset) >> kCgainShift;
      const int i = (gain_bound - 1 > 1 ? gain_bound - 1 : 1) + (lane >> 2);
      if (i <= gain_bound + 1) {
        const int32_t qcg = shl(i, kCgainShift) + gain_offset;
        const int ts = compute_max_theta(qcg, beta);
        int lo = (int)floor(.5 + theta * theta_scale_1 * 2 / M_PI * ts) - 2;
        int hi = (int)ceil(theta * theta_scale_1 * 2 / M_PI * ts);
        if (lo < 0) lo = 0;
        if (hi > ts - 1) hi = ts - 1;
        const int j = lo + (lane & 3);
        if (j <= hi) {
          c_gain = i;
          c_theta = j;
          c_qcg = qcg;
          c_ts = ts;
          c_qtheta = compute_theta(j, ts);
          c_k = compute_k(qcg, j, 0, n, beta);
          const double dist_theta = 2 - 2. * pvq_cos(theta - c_qtheta) * trig_1;
          double dist = gain_weight * (qcg - cg) * (qcg - cg) + qcg * (double)cg * dist_theta;
          dist *= cgain_2;
          c_alive = !(dist > dist0 + 1.0 * pvq_norm_lambda && c_k != 0);
          const double sin_theta = pvq_sin(theta) * trig_1;
          c_sinprod = sin_theta * pvq_sin(c_qtheta) * trig_1;
          c_g2 = qcg * (double)cg * c_sinprod * cgain_2;
          c_rate_ts = .9 * (M_LOG2E * log((double)ts));
        }
      }
    }
  } else if (lane < 14) {
    if ((is_keyframe && pli == 0) || corr < .5 || cg < (int32_t)shl(2, kCgainShift)) {
      const int gain_bound = cg >> kCgainShift;
      const int i = (gain_bound > 1 ? gain_bound : 1) + (lane - 12);
      if (i <= gain_bound + 1) {
        c_gain = i;
        c_qcg = shl(i, kCgainShift);
        c_k = compute_k(c_qcg, -1, 1, n, beta);
        double dist = gain_weight * (c_qcg - cg) * (c_qcg - cg);
        dist *= cgain_2;
        c_alive = !(dist > dist0 && c_k != 0);
        c_g2 = c_qcg * (double)cg * cgain_2;
      }
    }
  }
  unsigned alive_w = __ballot_sync(kFull, c_alive && lane < 12);
  unsigned alive_n = __ballot_sync(kFull, c_alive && lane >= 12);

  // ---- events ------------------------------------------------------------------------------------------
  int ya[E];     // pulses of the running search, magnitudes (the reference's y_tmp without signs)
  int ybest[E];  // signed pulses of the best candidate so far
#pragma unroll
  for (int e = 0; e < E; e++) ya[e] = ybest[e] = 0;
  int prev_k = 0, best_lane = -1;
  bool noref_started = false;
  while (alive_w | alive_n) {
    unsigned grp;
    int kcur, leader;
    const bool noref_ev = alive_w == 0;
    if (!noref_ev) {
      const int key = (alive_w >> lane) & 1 ? (c_k << 4) | lane : 0x7fffffff;
      const int mn = wmin(key);
      kcur = mn >> 4;
      leader = mn & 15;
      grp = __ballot_sync(kFull, ((alive_w >> lane) & 1) && c_k == kcur);
      alive_w &= ~grp;
    } else {
      leader = __ffs(alive_n) - 1;
      grp = 1u << leader;
      alive_n &= ~grp;
      kcur = __shfl_sync(kFull, c_k, leader);
      if (!noref_started) prev_k = 0;
      noref_started = true;
    }
    const int nn = noref_ev ? n : n - 1;
    double cos_dist = 0;
    if (!noref_ev && kcur == 0) {
#pragma unroll
      for (int e = 0; e < E; e++) ya[e] = 0;
    } else {
      // ---- pvq_search_rdo_double (src/pvq_encoder.c:93) on |x| ------------------------------------------
      const double g2 = wbcast(c_g2, leader);
      const int k = kcur;
      int xa[E];
      double xd[E];
      float xf[E];
      long long sxx = 0;
      int xmax = 0, sl1 = 0;
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int v = noref_ev ? x16[e] : xr[e];
        const int a = e * 32 + lane < nn ? abs(v) : 0;
        xa[e] = a;
        xd[e] = (double)a;  // == fabs((double)(float)xcoeff[j]): |int16| is exact in float
        xf[e] = (float)a;
        sxx += (long long)a * a;
        sl1 += a;
        xmax = a > xmax ? a : xmax;
      }
      const double xx = (double)wsum64(sxx);
      xmax = wmax(xmax);
      const double norm_1 = 1. / sqrt(1e-30 + xx);
      const double lambda = pvq_norm_lambda / (1e-30 + g2);
      double xy = 0, yy = 0;
      int i = 0;
      if (prev_k > 0 && prev_k <= k) {
        long long sxy = 0, syy = 0;
        int si = 0;
#pragma unroll
        for (int e = 0; e < E; e++) {
          if (e * 32 + lane >= nn) ya[e] = 0;
          sxy += (long long)xa[e] * ya[e];
          syy += (long long)ya[e] * ya[e];
          si += ya[e];
        }
        xy = (double)wsum64(sxy);
        yy = (double)wsum64(syy);
        i = wsum(si);
      } else if (k > 2) {
        const double l1_norm = (double)wsum(sl1);
        const double l1_inv = 1. / (l1_norm > 1e-100 ? l1_norm : 1e-100);
        long long sxy = 0, syy = 0;
        int si = 0;
#pragma unroll
        for (int e = 0; e < E; e++) {
          const double tmp = k * xd[e] * l1_inv;
          const int f = (int)floor(tmp);
          ya[e] = (e * 32 + lane < nn && f > 0) ? f : 0;
          sxy += (long long)xa[e] * ya[e];
          syy += (long long)ya[e] * ya[e];
          si += ya[e];
        }
        xy = (double)wsum64(sxy);
        yy = (double)wsum64(syy);
        i = wsum(si);
      } else {
#pragma unroll
        for (int e = 0; e < E; e++) ya[e] = 0;
      }
      const int rdo_pulses = 1 + k / 4;
      double delta_rate = 3. / nn;
      double accel_rate = 0.;
      if (k == 1) {
        if (nn == 15) {
          accel_rate = -8. / nn;
          delta_rate = 4.5 / nn - accel_rate;
        } else if (nn == 8) {
          accel_rate = 5.7 / nn;
          delta_rate = 9.3 / nn - accel_rate;
        }
      }
      for (; i < k; i++) {
        const bool plain = i < k - rdo_pulses;
        // contenders per slot
        bool cont[E];
        double tval[E];  // RDO: the element's objective
        int total;
        if (plain) {
          const double bound = (xy + xmax) * (xy + xmax) * (yy + 2. * i + 1.);
          if (bound < DAALA_B200_PVQ_EXACT_BOUND) {
            const float xyf = (float)xy, yyf1 = (float)(yy + 1.);
            unsigned key[E], kmax = 0;
#pragma unroll
            for (int e = 0; e < E; e++) {
              float a = xyf + xf[e];
              a *= a;
              const float b = yyf1 + (float)(2 * ya[e]);
              key[e] = e * 32 + lane < nn ? __float_as_uint(__fdividef(a, b)) : 0u;
              kmax = key[e] > kmax ? key[e] : kmax;
            }
            const unsigned mx = wmaxu(kmax);
            const unsigned thr = mx > 64u ? mx - 64u : 0u;
            int cnt = 0;
#pragma unroll
            for (int e = 0; e < E; e++) {
              cont[e] = e * 32 + lane < nn && key[e] >= thr;
              cnt += cont[e];
            }
            total = wsum(cnt);
          } else {
            // products may round: the reference's literal scan over every element
#pragma unroll
            for (int e = 0; e < E; e++) cont[e] = e * 32 + lane < nn;
            total = 2;
            PVQ_WARP_STAT(3);
          }
#pragma unroll
          for (int e = 0; e < E; e++) tval[e] = 0;
        } else {
          double tbl[4];
#pragma unroll
          for (int j = 0; j < 4; j++) tbl[j] = rsqrt_small_tbl((int)(yy + 2 * j + 1));
          int key[E], kmax = (int)0x80000000;
#pragma unroll
          for (int e = 0; e < E; e++) {
            const int j = e * 32 + lane;
            double t = xy + xd[e];
            const int yj = ya[e];
            const double ryy = yj < 4 ? (yj == 0 ? tbl[0] : yj == 1 ? tbl[1] : yj == 2 ? tbl[2] : tbl[3])
                                      : rsqrt_small_tbl((int)(yy + 2 * yj + 1));
            t = 2 * t * norm_1 * ryy - lambda * j * (delta_rate + j * accel_rate);
            tval[e] = t;
            key[e] = j < nn ? ordered_key((float)t) : (int)0x80000000;
            kmax = key[e] > kmax ? key[e] : kmax;
          }
          const int mx = wmax(kmax);
          int cnt = 0;
#pragma unroll
          for (int e = 0; e < E; e++) {
            cont[e] = e * 32 + lane < nn && key[e] == mx;
            cnt += cont[e];
          }
          total = wsum(cnt);
        }
        int pos;  // index of the chosen element
        PVQ_WARP_STAT(total == 1 ? 0 : plain ? 1 : 2);
        if (total == 1) {
          int mine = -1;
#pragma unroll
          for (int e = 0; e < E; e++) if (cont[e]) mine = e * 32 + lane;
          const unsigned who = __ballot_sync(kFull, mine >= 0);
          pos = __shfl_sync(kFull, mine, __ffs(who) - 1);
        } else {
          // the reference's sequential scan restricted to the contenders, in index order
          pos = -1;
          double ba = 0, bb = 1;
#pragma unroll
          for (int e = 0; e < E; e++) {
            unsigned mk = __ballot_sync(kFull, cont[e]);
            while (mk) {
              const int l = __ffs(mk) - 1;
              mk &= mk - 1;
              if (plain) {
                const int xj = __shfl_sync(kFull, xa[e], l);
                const int yj = __shfl_sync(kFull, ya[e], l);
                double a = xy + (double)xj;
                const double b = yy + 2 * yj + 1;
                a *= a;
                if (pos < 0 || a * bb > ba * b) { ba = a; bb = b; pos = e * 32 + l; }
              } else {
                const double t = wbcast(tval[e], l);
                if (pos < 0 || t > ba) { ba = t; pos = e * 32 + l; }
              }
            }
          }
        }
        const int src = pos & 31, slot = pos >> 5;
        const int px = __shfl_sync(kFull, pick<E>(xa, slot), src);
        const int py = __shfl_sync(kFull, pick<E>(ya, slot), src);
        xy = xy + (double)px;
        yy = yy + 2 * py + 1;
#pragma unroll
        for (int e = 0; e < E; e++) if (e * 32 + lane == pos) ya[e]++;
      }
      cos_dist = xy / (1e-100 + sqrt(xx * yy));
    }
    prev_k = kcur;
    // ---- od_pvq_rate, closed form (src/pvq_encoder.c:247), shared part of the event -------------------------
    double rate_base = 0;
    if (kcur != 0) {
      int sj = 0;
#pragma unroll
      for (int e = 0; e < E; e++) if (e * 32 + lane < nn) sj += (e * 32 + lane) * ya[e];
      const int sum = wsum(sj);
      const double f = sum / (double)(kcur * n);
      const double t = log(n * 2 * (1 * f + .025)) * kcur / n;
      rate_base = (1 + .4 * f) * n * (M_LOG2E * log(1 + (0 > t ? 0 : t))) + 3;
    }
    // ---- cost of every candidate of the event, then the reference's in-order fold ------------------------
    double cost = 0;
    if ((grp >> lane) & 1) {
      double rate = rate_base;
      double dist;
      if (noref_ev) {
        dist = gain_weight * (c_qcg - cg) * (c_qcg - cg) + c_qcg * (double)cg * (2 - 2 * cos_dist);
      } else {
        if (c_gain > 0 && c_theta >= 0) {
          rate += c_rate_ts;
          if (is_keyframe && pli == 0) rate += 6;
          if (c_gain == icgr) rate -= .5;
        }
        const double dist_theta = 2 - 2. * pvq_cos(theta - c_qtheta) * trig_1 + c_sinprod * (2 - 2 * cos_dist);
        dist = gain_weight * (c_qcg - cg) * (c_qcg - cg) + c_qcg * (double)cg * dist_theta;
      }
      dist *= cgain_2;
      c_dist = dist;
      cost = dist + pvq_norm_lambda * rate;
    }
    bool improved = false;
    for (unsigned gm = grp; gm; gm &= gm - 1) {
      const int l = __ffs(gm) - 1;
      const double cl = wbcast(cost, l);
      if (noref_ev ? cl <= best_cost : cl < best_cost) {
        best_cost = cl;
        best_lane = l;
        improved = true;
      }
    }
    if (improved) {
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int v = noref_ev ? x16[e] : xr[e];
        ybest[e] = e * 32 + lane < nn ? (v < 0 ? -ya[e] : ya[e]) : 0;
      }
    }
  }

  // ---- the winner --------------------------------------------------------------------------------------
  int qg = 0, best_k = 0;
  int noref = is_keyframe ? 1 : 0;
  *itheta = is_keyframe ? -1 : 0;
  *max_theta = 0;
  theta = 0;  // best_qtheta
  if (best_lane >= 0) {
    qg = __shfl_sync(kFull, c_gain, best_lane);
    best_k = __shfl_sync(kFull, c_k, best_lane);
    *itheta = __shfl_sync(kFull, c_theta, best_lane);
    *max_theta = __shfl_sync(kFull, c_ts, best_lane);
    theta = __shfl_sync(kFull, c_qtheta, best_lane);
    best_dist = wbcast(c_dist, best_lane);
    noref = best_lane >= 12;
  }
  int skip = 0;
  if (noref) {
    if (qg == 0) skip = 1;
  } else {
    if (!is_keyframe && qg == 0) skip = icgr ? 1 : 2;
    if (qg == icgr && *itheta == 0 && !cfl_enabled) skip = 2;
  }
  int32_t res[E];
  if (skip) {
#pragma unroll
    for (int e = 0; e < E; e++) res[e] = (skip == 2 && e * 32 + lane < n) ? r0[e * 32 + lane] : 0;
  } else {
    if (noref) gain_offset = 0;
    g = gain_expand(shl(qg, kCgainShift) + gain_offset, q0, beta);
    // od_pvq_synthesis_partial, src/pvq.c:1037
    const int nn = n - !noref;
    int syy = 0;
#pragma unroll
    for (int e = 0; e < E; e++) if (e * 32 + lane < nn) syy += ybest[e] * ybest[e];
    const int yy = wsum(syy);
    int gshift = ilog((uint32_t)g) - 14;
    if (gshift < 0) gshift = 0;
    int32_t scale;
    if (yy == 0) {
      scale = 0;
    } else {
      int rsh;
      const int16_t rs = rsqrt32(yy, &rsh);
      scale = vshr_round64(rs * (int64_t)g, rsh + gshift - 16);
    }
    const int qshift = kQmInvShift - gshift;
    if (noref) {
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int j = e * 32 + lane;
        const int32_t v = mul16_32_q16(ybest[e], scale);
        res[e] = j < n ? shr_round(v * qm_inv[j], qshift) : 0;
      }
    } else {
      scale = round32(scale * (1. / 32768) * pvq_sin(theta));
      int xs[E], f[E];
#pragma unroll
      for (int e = 0; e < E; e++) f[e] = e * 32 + lane < nn ? (int16_t)mul16_32_q16(ybest[e], scale) : 0;
      const int xm = (int16_t)floor(.5 + -s * (shr_round(g, gshift)) * (1. / 32768) * pvq_cos(theta));
#pragma unroll
      for (int e = 0; e < E; e++) {
        // value of element idx - 1
        const int t1 = __shfl_sync(kFull, f[e], (lane + 31) & 31);
        const int t2 = e > 0 ? __shfl_sync(kFull, f[e > 0 ? e - 1 : 0], 31) : 0;
        const int fprev = lane == 0 ? t2 : t1;
        const int j = e * 32 + lane;
        xs[e] = j < m ? f[e] : (j == m ? xm : fprev);
        if (j >= n) xs[e] = 0;
      }
      householder_apply_warp<E>(lane, xs, xs, r16, n);
#pragma unroll
      for (int e = 0; e < E; e++) {
        const int j = e * 32 + lane;
        res[e] = j < n ? shr_round(xs[e] * qm_inv[j], qshift) : 0;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int j = e * 32 + lane;
    if (j < n) {
      out[j] = res[e];
      yout[j] = ybest[e];
    }
  }
  *vk = best_k;
  *skip_term = skip_dist - best_dist;
  if (is_keyframe) return noref ? qg : (qg < icgr ? -2 * (qg - icgr) - 1 : (qg < 2 * icgr ? 2 * (qg - icgr) : qg - 1));
  {
    const int a = qg + 1, b = icgr + 1;
    return noref ? qg - 1 : (a < b ? -2 * (a - b) - 1 : (a < 2 * b ? 2 * (a - b) : a - 1));
  }
}

}  // namespace pvq
}  // namespace daala_b200
