// TF (time-frequency resolution switching) helpers of the reference, src/tf.c:38-277, as drop-in
// host-pointer symbols (section A of include/daala_b200.h).  Only od_tf_up_hv_lp is live in the codec
// (od_resample_luma_coeffs, src/intra.c:72; fused into the engine's chroma-from-luma kernel); the others
// are exported because tf.h declares them and tools link them.  One CTA per call, the block in shared
// memory; the 1-D lifting of od_tf_filter is a serial chain per row / column (one thread each).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "daala_b200.h"

namespace {

constexpr int kMax = 64;   // OD_BSIZE_MAX

enum Op { kUpHLp, kUpVLp, kUpHvLp, kUpHv, kDownHv, kFilter2d, kFilterInv2d, kConvertDown };

// OD_HAAR_KERNEL, src/tf.h:35
__device__ __forceinline__ void haar_kernel(int32_t& ll, int32_t& lh, int32_t& hl, int32_t& hh) {
  ll += hl;
  hh -= lh;
  const int32_t t = (ll - hh) >> 1;
  lh = t - lh;
  hl = t - hl;
  ll -= lh;
  hh += hl;
}

// OD_DCT_RSHIFT = OD_UNBIASED_RSHIFT32, src/filter.h:38
__device__ __forceinline__ int32_t dct_rshift(int32_t a, int b) {
  return (int32_t)(((uint32_t)a >> (32 - b)) + (uint32_t)a) >> b;
}

// od_tf_filter / od_tf_filter_inv (src/tf.c:158,172) along one line of n samples with stride `st`
__device__ void filter_line(int32_t* p, int st, int n, bool inv) {
  const int m = (n >> 1) - 1;
  if (!inv) {
    int32_t* v = p + st;
    for (int i = 0; i < m; i++) {
      int32_t* u = v;
      v += st << 1;
      *u += *v >> 1;
      *v -= *u >> 1;
    }
  } else {
    int32_t* u = p + st * (n - 1);
    for (int i = 0; i < m; i++) {
      int32_t* v = u;
      u -= st << 1;
      *v += *u >> 1;
      *u -= *v >> 1;
    }
  }
}

__device__ void filter_2d(int32_t* b, int ld, int x0, int y0, int n, bool inv) {
  // forward: rows then columns; inverse: columns then rows (src/tf.c:186-224)
  for (int pass = 0; pass < 2; pass++) {
    const bool rows = (pass == 0) != inv;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (rows) filter_line(b + (y0 + i) * ld + x0, 1, n, inv);
      else filter_line(b + y0 * ld + x0 + i, ld, n, inv);
    }
    __syncthreads();
  }
}

// dst (n x n at (x0, y0) of d) = od_tf_down_hv of the n x n block at (x0, y0) of s
__device__ void down_hv(int32_t* d, const int32_t* s, int ld, int x0, int y0, int n) {
  const int h = n >> 1;
  for (int i = threadIdx.x; i < h * h; i += blockDim.x) {
    const int y = i / h, x = i % h;
    const int vs = y & 1, hs = x & 1;
    int32_t ll = s[(y0 + 2 * y + vs) * ld + x0 + 2 * x + hs];
    int32_t lh = s[(y0 + 2 * y + vs) * ld + x0 + 2 * x + 1 - hs];
    int32_t hl = s[(y0 + 2 * y + 1 - vs) * ld + x0 + 2 * x + hs];
    int32_t hh = s[(y0 + 2 * y + 1 - vs) * ld + x0 + 2 * x + 1 - hs];
    haar_kernel(ll, lh, hl, hh);
    d[(y0 + y) * ld + x0 + x] = ll;
    d[(y0 + y) * ld + x0 + x + h] = lh;
    d[(y0 + y + h) * ld + x0 + x] = hl;
    d[(y0 + y + h) * ld + x0 + x + h] = hh;
  }
  __syncthreads();
}

// io: packed [kMax][kMax] block in, result out (same layout).  a, b: op-specific (dx / dy, sizes, filter).
__global__ void __launch_bounds__(256) k_tf(int32_t* io, int op, int n, int a, int b) {
  __shared__ int32_t s0[kMax * kMax], s1[kMax * kMax];
  for (int i = threadIdx.x; i < kMax * kMax; i += blockDim.x) { s0[i] = io[i]; s1[i] = 0; }
  __syncthreads();
  int32_t* res = s1;
  if (op == kUpHLp) {
    for (int i = threadIdx.x; i < n * (n >> 1); i += blockDim.x) {
      const int y = i / (n >> 1), x = i % (n >> 1);
      int32_t ll = s0[y * kMax + x], lh = s0[y * kMax + x + a];
      lh = ll - lh;
      ll -= dct_rshift(lh, 1);
      const int hs = x & 1;
      s1[y * kMax + 2 * x + hs] = ll;
      s1[y * kMax + 2 * x + 1 - hs] = lh;
    }
  } else if (op == kUpVLp) {
    for (int i = threadIdx.x; i < (n >> 1) * n; i += blockDim.x) {
      const int y = i / n, x = i % n;
      int32_t ll = s0[y * kMax + x], hl = s0[(y + a) * kMax + x];
      hl = ll - hl;
      ll -= dct_rshift(hl, 1);
      const int vs = y & 1;
      s1[(2 * y + vs) * kMax + x] = ll;
      s1[(2 * y + 1 - vs) * kMax + x] = hl;
    }
  } else if (op == kUpHvLp || op == kUpHv) {
    // _lp: quarter at offsets (dx, dy) = (a, b), n/2 x n/2 outputs pairs; up_hv: 2x2 group of n x n, offsets n
    const int cnt = op == kUpHvLp ? n >> 1 : n;
    const int dx = op == kUpHvLp ? a : n, dy = op == kUpHvLp ? b : n;
    for (int i = threadIdx.x; i < cnt * cnt; i += blockDim.x) {
      const int y = i / cnt, x = i % cnt;
      int32_t ll = s0[y * kMax + x], lh = s0[y * kMax + x + dx];
      int32_t hl = s0[(y + dy) * kMax + x], hh = s0[(y + dy) * kMax + x + dx];
      haar_kernel(ll, hl, lh, hh);   // lh and hl swapped, as in the reference
      const int vs = y & 1, hs = x & 1;
      s1[(2 * y + vs) * kMax + 2 * x + hs] = ll;
      s1[(2 * y + vs) * kMax + 2 * x + 1 - hs] = lh;
      s1[(2 * y + 1 - vs) * kMax + 2 * x + hs] = hl;
      s1[(2 * y + 1 - vs) * kMax + 2 * x + 1 - hs] = hh;
    }
  } else if (op == kDownHv) {
    down_hv(s1, s0, kMax, 0, 0, n);
  } else if (op == kFilter2d || op == kFilterInv2d) {
    filter_2d(s0, kMax, 0, 0, n, op == kFilterInv2d);
    res = s0;
  } else {
    // od_convert_block_down (src/tf.c:227): from size n down to size a, b = filter; every level: inverse TF
    // filter of each block, then od_tf_down_hv of it into its four quadrants
    int32_t *cur = s0, *nxt = s1;
    for (int m = n; m > a; m >>= 1) {
      for (int by = 0; by < n; by += m)
        for (int bx = 0; bx < n; bx += m) {
          if (b) filter_2d(cur, kMax, bx, by, m, true);
          down_hv(nxt, cur, kMax, bx, by, m);
        }
      int32_t* t = cur; cur = nxt; nxt = t;
    }
    res = cur;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kMax * kMax; i += blockDim.x) io[i] = res[i];
}

struct Ctx {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  int32_t* pinned = nullptr;
  int32_t* dev = nullptr;
};

[[noreturn]] void fatal(const char* what, cudaError_t e) {
  fprintf(stderr, "libdaala_b200: fatal: %s: %s (no CPU fallback exists)\n", what, cudaGetErrorString(e));
  abort();
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) fatal(#x, e_); } while (0)

// rows x cols of src (stride) -> kernel -> orows x ocols into dst
void run(int op, int n, int a, int b, int32_t* dst, int dstride, int orows, int ocols, const int32_t* src, int sstride,
         int rows, int cols) {
  static Ctx c;
  std::lock_guard<std::mutex> g(c.mu);
  if (rows > kMax || cols > kMax || orows > kMax || ocols > kMax) fatal("TF block larger than 64x64", cudaErrorInvalidValue);
  if (!c.stream) {
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    CK(cudaMallocHost((void**)&c.pinned, sizeof(int32_t) * kMax * kMax));
    CK(cudaMalloc((void**)&c.dev, sizeof(int32_t) * kMax * kMax));
  }
  memset(c.pinned, 0, sizeof(int32_t) * kMax * kMax);
  for (int y = 0; y < rows; y++) memcpy(c.pinned + y * kMax, src + (size_t)y * sstride, sizeof(int32_t) * cols);
  CK(cudaMemcpyAsync(c.dev, c.pinned, sizeof(int32_t) * kMax * kMax, cudaMemcpyHostToDevice, c.stream));
  k_tf<<<1, 256, 0, c.stream>>>(c.dev, op, n, a, b);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(c.pinned, c.dev, sizeof(int32_t) * kMax * kMax, cudaMemcpyDeviceToHost, c.stream));
  CK(cudaStreamSynchronize(c.stream));
  for (int y = 0; y < orows; y++) memcpy(dst + (size_t)y * dstride, c.pinned + y * kMax, sizeof(int32_t) * ocols);
}

}  // namespace

extern "C" {

void od_tf_up_h_lp(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int dx, int n) {
  run(kUpHLp, n, dx, 0, dst, dstride, n, n, src, sstride, n, (n >> 1) + dx);
}
void od_tf_up_v_lp(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int dy, int n) {
  run(kUpVLp, n, dy, 0, dst, dstride, n, n, src, sstride, (n >> 1) + dy, n);
}
void od_tf_up_hv_lp(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int dx, int dy, int n) {
  run(kUpHvLp, n, dx, dy, dst, dstride, n, n, src, sstride, (n >> 1) + dy, (n >> 1) + dx);
}
void od_tf_up_hv(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int n) {
  run(kUpHv, n, 0, 0, dst, dstride, 2 * n, 2 * n, src, sstride, 2 * n, 2 * n);
}
void od_tf_down_hv(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int n) {
  run(kDownHv, n, 0, 0, dst, dstride, n, n, src, sstride, n, n);
}
void od_tf_filter_2d(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int n) {
  run(kFilter2d, n, 0, 0, dst, dstride, n, n, src, sstride, n, n);
}
void od_tf_filter_inv_2d(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int n) {
  run(kFilterInv2d, n, 0, 0, dst, dstride, n, n, src, sstride, n, n);
}
void od_convert_block_down(od_coeff* dst, int dstride, const od_coeff* src, int sstride, int curr_size, int dest_size,
                           int filter) {
  const int n = 4 << curr_size;   // OD_LOG_BSIZE0 = 2
  run(kConvertDown, n, 4 << dest_size, filter, dst, dstride, n, n, src, sstride, n, n);
}

}  // extern "C"
