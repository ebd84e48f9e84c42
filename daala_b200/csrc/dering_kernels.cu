// Directional deringing filter of whole planes (reference src/dering.c: od_dering :252 with
// od_dir_find8 :61, od_filter_dering_direction_c :132, od_filter_dering_orthogonal_c :172,
// od_compute_thresh :237) -- SURVEY.md 8(f) rank 1, the row after the transform / PVQ / MC path.
//
// Parity: bit-exact against oracle/port_dering.c (pinned against od_dering) in tests/test_gpu_dering.py.
// Consumers: the keyframe engine's deringing stage (csrc/kf_engine.cu, levels given) and the level search
// (csrc/dering_search.cu, src/encode.c:2680-2811).
//
// Mapping: one 256-thread CTA per superblock.  The (B+6)^2 int16 window (3-sample apron, 30000 where the
// frame ends) is staged once in shared memory; one thread per 8x8 block finds the direction (8 x 64
// integer adds, exact) and the threshold; the two filter passes run pixel-parallel out of shared memory
// (pass 2 reads pass 1's output inside the superblock and the unfiltered apron outside it, as the
// reference's `in` buffer does).  HBM traffic is the minimum: every sample is read once (+apron) and
// written once, 2 B each.
#include <cuda_runtime.h>
#include <stdint.h>

#include "daala_b200.h"

namespace daala_b200 {
namespace dering {

constexpr int kBorder = 3;
constexpr int kPitch = 64 + 2 * kBorder;
constexpr int kOutside = 30000;

// step k = 1..3 along direction d as (rows, columns); OD_DIRECTION_OFFSETS_TABLE, src/dering.c:39-48
__constant__ signed char kStep[8][3][2] = {
    {{-1, 1}, {-2, 2}, {-3, 3}}, {{0, 1}, {-1, 2}, {-1, 3}}, {{0, 1}, {0, 2}, {0, 3}}, {{0, 1}, {1, 2}, {1, 3}},
    {{1, 1}, {2, 2}, {3, 3}},    {{1, 0}, {2, 1}, {3, 1}},   {{1, 0}, {2, 0}, {3, 0}}, {{1, 0}, {2, -1}, {3, -1}},
};
// OD_THRESH_TABLE_Q8, src/dering.c:225
__constant__ short kThreshQ8[18] = {128, 134, 150, 168, 188, 210, 234, 262, 292, 327, 365, 408, 455, 509, 569, 635, 710, 768};

__device__ __forceinline__ int line_of(int d, int i, int j) {
  switch (d) {
    case 0: return i + j;
    case 1: return i + j / 2;
    case 2: return i;
    case 3: return 3 + i - j / 2;
    case 4: return 7 + i - j;
    case 5: return 3 - i / 2 + j;
    case 6: return j;
    default: return i / 2 + j;
  }
}

// Direction of one 8x8 block at `img` (shared-memory window, pitch kPitch).  Same integers as od_dir_find8:
// cost[d] = sum over the lines of direction d of (line sum)^2 * 840 / (line length).
__device__ int find_direction(const int16_t* img, int coeff_shift, int32_t* var) {
  int32_t cost[8];
#pragma unroll 1
  for (int d = 0; d < 8; d++) {
    int sum[15], len[15];
#pragma unroll
    for (int l = 0; l < 15; l++) sum[l] = len[l] = 0;
    for (int i = 0; i < 8; i++) {
      for (int j = 0; j < 8; j++) {
        const int l = line_of(d, i, j);
        sum[l] += img[i * kPitch + j] >> coeff_shift;
        len[l]++;
      }
    }
    int32_t c = 0;
    for (int l = 0; l < 15; l++)
      if (len[l]) c += sum[l] * sum[l] * (840 / len[l]);
    cost[d] = c;
  }
  int best = 0;
  int32_t best_cost = 0;
  for (int d = 0; d < 8; d++) {
    if (cost[d] > best_cost) {
      best_cost = cost[d];
      best = d;
    }
  }
  *var = (best_cost - cost[(best + 4) & 7]) >> 10;
  return best;
}

// frames of a batch: blockIdx.z, element pitches between consecutive frames (all zero for a single plane)
struct BatchPitch {
  long long y, x, dir, thr;
};

__global__ void __launch_bounds__(256) k_dering_sb(const __grid_constant__ daala_b200_dering_params p0,
                                                   const __grid_constant__ BatchPitch bp) {
  daala_b200_dering_params p = p0;
  p.y += blockIdx.z * bp.y;
  p.x += blockIdx.z * bp.x;
  p.dir += blockIdx.z * bp.dir;
  if (p.sb_threshold) p.sb_threshold += blockIdx.z * bp.thr;
  __shared__ int16_t win[kPitch * kPitch];    // unfiltered input + apron
  __shared__ int16_t mid[kPitch * kPitch];    // pass-1 output inside the superblock, input in the apron
  __shared__ int s_thr[64];
  __shared__ int s_dir[64];
  const int sbx = blockIdx.x, sby = blockIdx.y;
  const int lb = 3 - p.xdec, n = 1 << lb, B = 64 >> p.xdec;
  const int16_t* x = p.x + (size_t)sby * B * p.xstride + (size_t)sbx * B;
  int16_t* y = p.y + (size_t)sby * B * p.ystride + (size_t)sbx * B;
  const int i_lo = sby ? -kBorder : 0, i_hi = B + (sby != p.nvsb - 1 ? kBorder : 0);
  const int j_lo = sbx ? -kBorder : 0, j_hi = B + (sbx != p.nhsb - 1 ? kBorder : 0);
  for (int idx = threadIdx.x; idx < kPitch * kPitch; idx += 256) {
    const int i = idx / kPitch - kBorder, j = idx % kPitch - kBorder;
    const bool inside = i >= i_lo && i < i_hi && j >= j_lo && j < j_hi;
    const int16_t v = inside ? x[(ptrdiff_t)i * p.xstride + j] : (int16_t)kOutside;
    win[idx] = v;
    mid[idx] = v;
  }
  __syncthreads();
  const int16_t* in = win + kBorder * kPitch + kBorder;
  int16_t* in2 = mid + kBorder * kPitch + kBorder;
  if (threadIdx.x < 64) {
    const int by = threadIdx.x >> 3, bx = threadIdx.x & 7;
    int32_t* dslot = p.dir + (size_t)(sby * 8 + by) * p.dir_stride + sbx * 8 + bx;
    const int base = p.sb_threshold ? p.sb_threshold[sby * p.nhsb + sbx] : p.threshold;
    int d, thr;
    if (p.pli == 0) {
      int32_t var;
      if (p.dir_format == 2) {
        // direction and variance of an earlier pass over the same input (the level search filters one plane
        // five times, and the final application a sixth)
        const int packed = *dslot;
        d = packed & 7;
        var = packed >> 3;
      } else {
        d = find_direction(in + by * 8 * kPitch + bx * 8, p.coeff_shift, &var);
        *dslot = p.dir_format == 1 ? (d | (var << 3)) : d;
      }
      int v = var >> 6;
      if (v > 32767) v = 32767;
      const int lg = v ? 32 - __clz(v) : 0;
      thr = (base * kThreshQ8[lg] + 128) >> 8;
    } else {
      d = p.dir_format ? (*dslot & 7) : *dslot;
      thr = base;
    }
    // skipped neighbourhood -> untouched (DAALA_ODINTRIN form, src/dering.c:298-318)
    int u0 = 0, v0 = 0, u1 = 2 >> p.xdec, v1 = 2 >> p.xdec;
    if (p.overlap) {
      u0 -= sbx != 0;
      v0 -= sby != 0;
      u1 += sbx != p.nhsb - 1;
      v1 += sby != p.nvsb - 1;
    }
    // per-plane skip flags, one per 4x4 block of THIS plane: 16 >> xdec flags per superblock side
    // (call site src/encode.c:2789-2791)
    const uint8_t* sk = p.bskip + (size_t)(sby * (16 >> p.xdec)) * p.skip_stride + sbx * (16 >> p.xdec);
    bool all = true;
    for (int i = v0; i < v1; i++)
      for (int j = u0; j < u1; j++) all = all && sk[(ptrdiff_t)(((by << 1) >> p.xdec) + i) * p.skip_stride + ((bx << 1) >> p.xdec) + j];
    s_thr[threadIdx.x] = all ? 0 : thr;
    s_dir[threadIdx.x] = d;
  }
  __syncthreads();
  // pass 1: along the direction, taps 3 2 1 on either side
  for (int idx = threadIdx.x; idx < B * B; idx += 256) {
    const int i = idx / B, j = idx % B;
    const int blk = ((i >> lb) << 3) | (j >> lb);
    const int t = s_thr[blk], d = s_dir[blk];
    const int16_t c = in[i * kPitch + j];
    int16_t acc = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int o = kStep[d][k][0] * kPitch + kStep[d][k][1];
      const int16_t a = (int16_t)(in[i * kPitch + j + o] - c);
      const int16_t b = (int16_t)(in[i * kPitch + j - o] - c);
      if (abs((int)a) < t) acc = (int16_t)(acc + (3 - k) * a);
      if (abs((int)b) < t) acc = (int16_t)(acc + (3 - k) * b);
    }
    in2[i * kPitch + j] = (int16_t)(c + ((acc + 8) >> 4));
  }
  __syncthreads();
  // pass 2: across it, four unit taps with the tighter, change-dependent threshold
  for (int idx = threadIdx.x; idx < B * B; idx += 256) {
    const int i = idx / B, j = idx % B;
    const int blk = ((i >> lb) << 3) | (j >> lb);
    const int t = s_thr[blk], d = s_dir[blk];
    const int o = (d > 0 && d < 4) ? kPitch : 1;
    const int16_t c = in2[i * kPitch + j];
    const int moved = abs((int)c - (int)in[i * kPitch + j]);
    const int16_t lim = (int16_t)(t / 3 + moved < t ? t / 3 + moved : t);
    int16_t acc = 0;
    int16_t q;
    q = (int16_t)(in2[i * kPitch + j + o] - c);
    if (abs((int)q) < lim) acc = (int16_t)(acc + q);
    q = (int16_t)(in2[i * kPitch + j - o] - c);
    if (abs((int)q) < lim) acc = (int16_t)(acc + q);
    q = (int16_t)(in2[i * kPitch + j + 2 * o] - c);
    if (abs((int)q) < lim) acc = (int16_t)(acc + q);
    q = (int16_t)(in2[i * kPitch + j - 2 * o] - c);
    if (abs((int)q) < lim) acc = (int16_t)(acc + q);
    y[(size_t)i * p.ystride + j] = (int16_t)(c + ((3 * acc + 8) >> 4));
  }
  (void)n;
}

}  // namespace dering
}  // namespace daala_b200

extern "C" int daala_b200_dering_plane(const daala_b200_dering_params* prm, void* stream) {
  if (!prm || prm->nhsb < 1 || prm->nvsb < 1 || prm->xdec < 0 || prm->xdec > 1) return (int)cudaErrorInvalidValue;
  // not in place: a superblock's apron would read its neighbours' filtered output
  if (!prm->x || !prm->y || (const void*)prm->x == (const void*)prm->y) return (int)cudaErrorInvalidValue;
  dim3 grid(prm->nhsb, prm->nvsb);
  daala_b200::dering::BatchPitch bp = {0, 0, 0, 0};
  daala_b200::dering::k_dering_sb<<<grid, 256, 0, (cudaStream_t)stream>>>(*prm, bp);
  return (int)cudaGetLastError();
}

// The same for `nframes` planes of one geometry in one launch (internal: the keyframe engine's deringing stage).
extern "C" int daala_b200_dering_plane_batch(const daala_b200_dering_params* prm, int nframes, long long y_pitch,
                                             long long x_pitch, long long dir_pitch, long long thr_pitch, void* stream) {
  if (!prm || nframes < 1 || prm->nhsb < 1 || prm->nvsb < 1 || prm->xdec < 0 || prm->xdec > 1) return (int)cudaErrorInvalidValue;
  if (!prm->x || !prm->y || (const void*)prm->x == (const void*)prm->y) return (int)cudaErrorInvalidValue;
  dim3 grid(prm->nhsb, prm->nvsb, nframes);
  daala_b200::dering::BatchPitch bp = {y_pitch, x_pitch, dir_pitch, thr_pitch};
  daala_b200::dering::k_dering_sb<<<grid, 256, 0, (cudaStream_t)stream>>>(*prm, bp);
  return (int)cudaGetLastError();
}
