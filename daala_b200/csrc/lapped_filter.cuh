// Daala's 4-point lapped pre/post filter as register code.
//
// Only the 4-point filter is live in the codec (OD_FILT_SIZE() == 0,
// reference src/filter.h:77); it straddles a block edge: two samples on each
// side.  Semantics follow od_pre_filter4 (src/filter.c:147) and
// od_post_filter4 (src/filter.c:195) with the R=f parameter set
// s0 = 85/64, s1 = 75/64, p0 = -15/64, u0 = 33/64 (src/filter.c:137-140).
#pragma once

namespace daala_b200 {

// Forward (encoder side).  v0..v3 are consecutive samples across the edge.
__device__ __forceinline__ void lap_pre4(int& v0, int& v1, int& v2, int& v3) {
  int d3 = v0 - v3;
  int d2 = v1 - v2;
  int m1 = v1 - (d2 >> 1);
  int m0 = v0 - (d3 >> 1);
  // Biorthogonal scaling; strictly positive results are bumped by one so that
  // the truncating division of the post filter inverts the step exactly.
  d2 = (d2 * 85) >> 6;
  d2 += (int)((unsigned)(-d2) >> 31);
  d3 = (d3 * 75) >> 6;
  d3 += (int)((unsigned)(-d3) >> 31);
  d3 += (d2 * -15 + 32) >> 6;
  d2 += (d3 * 33 + 32) >> 6;
  m0 += d3 >> 1;
  m1 += d2 >> 1;
  v0 = m0;
  v1 = m1;
  v2 = m1 - d2;
  v3 = m0 - d3;
}

// Inverse (decoder side / encoder reconstruction).
__device__ __forceinline__ void lap_post4(int& v0, int& v1, int& v2, int& v3) {
  int d3 = v0 - v3;
  int d2 = v1 - v2;
  int m1 = v1 - (d2 >> 1);
  int m0 = v0 - (d3 >> 1);
  d2 -= (d3 * 33 + 32) >> 6;
  d3 -= (d2 * -15 + 32) >> 6;
  d3 = (d3 * 64) / 75;  // C division: truncates toward zero
  d2 = (d2 * 64) / 85;
  m0 += d3 >> 1;
  m1 += d2 >> 1;
  v0 = m0;
  v1 = m1;
  v2 = m1 - d2;
  v3 = m0 - d3;
}

// Apply to four samples `stride` ints apart, in place.
template <bool kPost>
__device__ __forceinline__ void lap4_inplace(int* p, int stride) {
  int a = p[0], b = p[stride], c = p[2 * stride], d = p[3 * stride];
  if (kPost) lap_post4(a, b, c, d); else lap_pre4(a, b, c, d);
  p[0] = a;
  p[stride] = b;
  p[2 * stride] = c;
  p[3 * stride] = d;
}

}  // namespace daala_b200
