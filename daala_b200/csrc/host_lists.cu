// Host-side construction of the work lists of one keyframe batch (no device code in this file).
//
// Everything the PVQ stage consumes besides pixels follows from the block-size map: block
// descriptors, the same-size neighbour of every luma block (od_hv_intra_pred's `top` / `left`,
// reference src/intra.c:46-47), dependency depths, the band-granular wave lists of the luma intra
// wavefront (pvq_kernels.cu: k_intra_band_ref) and the per-size-class band lists.  A real encoder
// decides block sizes per frame, so this runs once per frame on the host's critical path:
// linear passes and counting sorts, one thread per frame.  daala_b200/pvq.py holds the reference
// construction in numpy (block_list, sort_by_depth, band_wave_lists, ...); this produces the same
// arrays in the same order (tests/test_host_logic.py compares them element by element).
//
// Geometry: 4:2:0, three planes; the map has one byte per 8x8 luma unit (log2(block size) - 2),
// blocks never cross superblocks (reference src/block_size.h).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <stdio.h>
#include <thread>
#include <vector>

#include "daala_b200.h"

namespace {

inline int num_bands(int bs) { return bs == 0 ? 1 : bs == 1 ? 4 : bs == 2 ? 7 : 9; }
inline int band_class(int band) { return band < 3 ? 0 : band < 6 ? 1 : 2; }   // n <= 16, 32, 128
inline bool band_free(int band) { return band == 3 || band == 6; }

struct FrameLuma {
  std::vector<daala_b200_pvq_block> blocks;     // raster order of the block origin
  std::vector<int32_t> top, left;               // frame-local indices or -1
  std::vector<int32_t> d0, dh, dv;              // depth over both / the top / the left chain
};

void build_frame_luma(const uint8_t* map, int bstride, int nhsb, int nvsb, int frame, FrameLuma* out) {
  const int w4 = nhsb * 16, h4 = nvsb * 16;     // 4-pixel grid
  std::vector<int32_t> index((size_t)w4 * h4, -1);
  auto& b = out->blocks;
  b.clear();
  b.reserve((size_t)nhsb * nvsb * 48);
  for (int y4 = 0; y4 < h4; y4++) {
    const uint8_t* row = map + (size_t)(y4 >> 1) * bstride;
    const int uy = y4 >> 1;
    for (int ux = 0; ux < nhsb * 8; ux++) {
      const int bs = row[ux];
      int nb;                                    // blocks starting in this unit on this 4-pixel row
      if (bs == 0) {
        nb = 2;
      } else {
        const int mask = (1 << (bs - 1)) - 1;    // 8x8 units per block side - 1
        nb = (!(y4 & 1) && !(uy & mask) && !(ux & mask)) ? 1 : 0;
      }
      for (int k = 0; k < nb; k++) {
        const int x4 = 2 * ux + k;
        daala_b200_pvq_block blk;
        blk.coef_off = 0;
        blk.x0 = (uint16_t)(x4 * 4);
        blk.y0 = (uint16_t)(y4 * 4);
        blk.bs = (uint8_t)bs;
        blk.pli = 0;
        blk.xdec = 0;
        blk.frame = (uint8_t)frame;
        index[(size_t)y4 * w4 + x4] = (int32_t)b.size();
        b.push_back(blk);
      }
    }
  }
  const size_t n = b.size();
  out->top.assign(n, -1);
  out->left.assign(n, -1);
  out->d0.assign(n, 1);
  out->dh.assign(n, 1);
  out->dv.assign(n, 1);
  for (size_t i = 0; i < n; i++) {
    const int bs = b[i].bs, n4 = 1 << bs;
    const int x4 = b[i].x0 >> 2, y4 = b[i].y0 >> 2;
    int t = -1, l = -1;
    if (y4 > 0 && map[(size_t)((y4 - 1) >> 1) * bstride + (x4 >> 1)] == bs) t = index[(size_t)(y4 - n4) * w4 + x4];
    if (x4 > 0 && map[(size_t)(y4 >> 1) * bstride + ((x4 - 1) >> 1)] == bs) l = index[(size_t)y4 * w4 + (x4 - n4)];
    out->top[i] = t;
    out->left[i] = l;
    // neighbours come earlier in raster order: one pass gives the longest chain
    const int dt = t >= 0 ? out->d0[t] : 0, dl = l >= 0 ? out->d0[l] : 0;
    out->d0[i] = 1 + std::max(dt, dl);
    out->dh[i] = 1 + (t >= 0 ? out->dh[t] : 0);
    out->dv[i] = 1 + (l >= 0 ? out->dv[l] : 0);
  }
}

// run fn(0..n-1) on up to `nthreads` threads
template <class F>
void parallel_for(int n, int nthreads, F fn) {
  const int nt = std::max(1, std::min(nthreads, n));
  if (nt == 1) {
    for (int i = 0; i < n; i++) fn(i);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; t++)
    pool.emplace_back([&, t] {
      for (int i = t; i < n; i += nt) fn(i);
    });
  for (auto& th : pool) th.join();
}

template <class T>
T* to_c(const std::vector<T>& v) {
  T* p = (T*)malloc(sizeof(T) * (v.size() ? v.size() : 1));
  if (p && !v.empty()) memcpy(p, v.data(), sizeof(T) * v.size());
  return p;
}

long long assign_offsets(std::vector<daala_b200_pvq_block>& b) {
  long long off = 0;
  for (auto& x : b) {
    x.coef_off = (int32_t)off;
    const long long n2 = 16ll << (2 * x.bs);
    off += n2 < 512 ? n2 : 512;
  }
  return off;
}

}  // namespace

extern "C" {

daala_b200_keyframe_lists* daala_b200_host_keyframe_lists(const uint8_t* bsize, int nframes,
                                                          long long bsize_frame_pitch, int bstride, int nhsb,
                                                          int nvsb, int nthreads) {
  if (!bsize || nframes < 1 || nframes > 255 || nhsb < 1 || nvsb < 1) return nullptr;
  // coef_off is 32-bit: a batch never holds more coded coefficients than luma samples
  if ((long long)nframes * nhsb * nvsb * 4096 >= (1ll << 31)) return nullptr;
  auto* L = (daala_b200_keyframe_lists*)calloc(1, sizeof(daala_b200_keyframe_lists));
  if (!L) return nullptr;
  const bool prof = getenv("DAALA_B200_PROFILE_LISTS") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!prof) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "host_lists: %-28s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  // --- luma, per frame in parallel ---
  std::vector<FrameLuma> fl(nframes);
  parallel_for(nframes, nthreads, [&](int f) {
    build_frame_luma(bsize + (size_t)f * bsize_frame_pitch, bstride, nhsb, nvsb, f, &fl[f]);
  });
  lap("luma frames (parallel)");
  size_t nl = 0;
  std::vector<size_t> base(nframes);
  for (int f = 0; f < nframes; f++) { base[f] = nl; nl += fl[f].blocks.size(); }
  // stable counting sort of the frame-major raster order by dependency depth
  int maxd = 0;
  for (auto& f : fl) for (int d : f.d0) maxd = std::max(maxd, d);
  // offset of (depth d, frame f) in the sorted order: depth-major, then frame, then raster
  std::vector<size_t> off((size_t)(maxd + 1) * nframes + 1, 0);
  for (int f = 0; f < nframes; f++)
    for (int d : fl[f].d0) off[(size_t)d * nframes + f + 1]++;
  for (size_t k = 1; k < off.size(); k++) off[k] += off[k - 1];
  std::vector<daala_b200_pvq_block> luma(nl);
  std::vector<int32_t> top(nl), left(nl), d0(nl), dh(nl), dv(nl);
  parallel_for(nframes, nthreads, [&](int f) {
    const FrameLuma& F = fl[f];
    const size_t n = F.blocks.size();
    std::vector<int32_t> pos(n);
    std::vector<size_t> cur(maxd + 1);
    for (int d = 0; d <= maxd; d++) cur[d] = off[(size_t)d * nframes + f];
    for (size_t i = 0; i < n; i++) pos[i] = (int32_t)cur[F.d0[i]]++;
    for (size_t i = 0; i < n; i++) {
      const int32_t p = pos[i];
      luma[p] = F.blocks[i];
      top[p] = F.top[i] >= 0 ? pos[F.top[i]] : -1;
      left[p] = F.left[i] >= 0 ? pos[F.left[i]] : -1;
      d0[p] = F.d0[i];
      dh[p] = F.dh[i];
      dv[p] = F.dv[i];
    }
  });
  L->luma_total = assign_offsets(luma);
  lap("depth sort + remap");
  // --- band-granular waves, one task per size class ---
  parallel_for(3, nthreads, [&](int c) {
    std::vector<uint32_t> bulk, chain;
    std::vector<uint16_t> wave;
    std::vector<int32_t> first, count;
    int topd = 0;
    // entries of this class: wave of (block, band) and counts per (wave, band)
    const int band0 = c * 3;
    auto depth_of = [&](size_t i, int band) { return band % 3 == 0 ? d0[i] : band % 3 == 1 ? dh[i] : dv[i]; };
    for (int band = band0; band < band0 + 3; band++) {
      if (band_free(band)) {
        for (size_t i = 0; i < nl; i++)
          if (num_bands(luma[i].bs) > band) bulk.push_back((uint32_t)(i << 4) | (uint32_t)band);
        continue;
      }
      for (size_t i = 0; i < nl; i++)
        if (num_bands(luma[i].bs) > band) topd = std::max(topd, (int)depth_of(i, band));
    }
    // band 0 is special only in name: bands 0 / 3 / 6 sit at band % 3 == 0, and 3 / 6 are the free ones
    std::vector<size_t> cnt((size_t)(topd + 1) * 3 + 1, 0);
    for (int band = band0; band < band0 + 3; band++) {
      if (band_free(band)) continue;
      for (size_t i = 0; i < nl; i++)
        if (num_bands(luma[i].bs) > band) cnt[(size_t)(depth_of(i, band) - 1) * 3 + (band - band0) + 1]++;
    }
    for (size_t k = 1; k < cnt.size(); k++) cnt[k] += cnt[k - 1];
    chain.resize(cnt.back());
    wave.resize(cnt.back());
    {
      std::vector<size_t> cur(cnt.begin(), cnt.end() - 1);
      for (int band = band0; band < band0 + 3; band++) {
        if (band_free(band)) continue;
        for (size_t i = 0; i < nl; i++) {
          if (num_bands(luma[i].bs) <= band) continue;
          const int d = depth_of(i, band);
          const size_t p = cur[(size_t)(d - 1) * 3 + (band - band0)]++;
          chain[p] = (uint32_t)(i << 4) | (uint32_t)band;
          wave[p] = (uint16_t)(d - 1);
        }
      }
    }
    for (int d = 1; d <= topd; d++) {
      first.push_back((int32_t)cnt[(size_t)(d - 1) * 3]);
      count.push_back((int32_t)(cnt[(size_t)d * 3] - cnt[(size_t)(d - 1) * 3]));
    }
    L->chain[c] = to_c(chain);
    L->n_chain[c] = (int)chain.size();
    L->chain_wave[c] = to_c(wave);
    L->n_waves[c] = topd;
    L->wave_first[c] = to_c(first);
    L->wave_count[c] = to_c(count);
    L->bulk[c] = to_c(bulk);
    L->n_bulk[c] = (int)bulk.size();
  });
  lap("band waves");
  L->luma = to_c(luma);
  L->n_luma = (int)nl;
  L->dep_top = to_c(top);
  L->dep_left = to_c(left);
  L->depth = to_c(d0);
  // --- chroma: order (bs, frame, plane, y0, x0); bit 7 of xdec marks "co-located luma is 4x4" ---
  std::vector<daala_b200_pvq_block> part[4];
  parallel_for(4, nthreads, [&](int bs) {
    const int lvl = bs + 1, span = 1 << (lvl - 1);
    auto& out = part[bs];
    out.reserve((size_t)nframes * nhsb * nvsb * (128 >> (2 * bs)) / 2);
    for (int f = 0; f < nframes; f++) {
      const uint8_t* map = bsize + (size_t)f * bsize_frame_pitch;
      for (int pli = 1; pli < 3; pli++) {
        for (int uy = 0; uy < nvsb * 8; uy += span) {
          for (int ux = 0; ux < nhsb * 8; ux += span) {
            const int v = map[(size_t)uy * bstride + ux];
            const int eff = v > 1 ? v : 1;             // bs = max(obs, xdec), reference src/encode.c:1467
            if (eff != lvl) continue;
            daala_b200_pvq_block blk;
            blk.coef_off = 0;
            blk.x0 = (uint16_t)(ux * 4);
            blk.y0 = (uint16_t)(uy * 4);
            blk.bs = (uint8_t)bs;
            blk.pli = (uint8_t)pli;
            blk.xdec = (uint8_t)(1 | ((bs == 0 && v == 0) ? 0x80 : 0));
            blk.frame = (uint8_t)f;
            out.push_back(blk);
          }
        }
      }
    }
  });
  std::vector<daala_b200_pvq_block> chroma;
  size_t first_of_size[6] = {0, 0, 0, 0, 0, 0};   // first block with bs >= k
  chroma.reserve(part[0].size() + part[1].size() + part[2].size() + part[3].size());
  for (int bs = 0; bs < 4; bs++) {
    first_of_size[bs] = chroma.size();
    chroma.insert(chroma.end(), part[bs].begin(), part[bs].end());
  }
  first_of_size[4] = first_of_size[5] = chroma.size();
  L->chroma_total = assign_offsets(chroma);
  lap("chroma blocks");
  {
    // blocks are sorted by size, so the blocks that own band b are a suffix of the array
    std::vector<uint32_t> lists[3];
    for (int band = 0; band < 9; band++) {
      const int min_bs = band == 0 ? 0 : band < 4 ? 1 : band < 7 ? 2 : 3;
      auto& v = lists[band_class(band)];
      for (size_t i = first_of_size[min_bs]; i < chroma.size(); i++) v.push_back((uint32_t)(i << 4) | (uint32_t)band);
    }
    for (int c = 0; c < 3; c++) {
      L->chroma_list[c] = to_c(lists[c]);
      L->n_chroma_list[c] = (int)lists[c].size();
    }
  }
  L->chroma = to_c(chroma);
  L->n_chroma = (int)chroma.size();
  lap("chroma lists");
  return L;
}

void daala_b200_host_keyframe_lists_free(daala_b200_keyframe_lists* L) {
  if (!L) return;
  free(L->luma);
  free(L->chroma);
  free(L->dep_top);
  free(L->dep_left);
  free(L->depth);
  for (int c = 0; c < 3; c++) {
    free(L->chain[c]);
    free(L->chain_wave[c]);
    free(L->wave_first[c]);
    free(L->wave_count[c]);
    free(L->bulk[c]);
    free(L->chroma_list[c]);
  }
  free(L);
}

}  // extern "C"
