// The warp-per-band PVQ quantiser (daala_b200/csrc/pvq_warp.cuh) compiled for the host on the SIMT
// emulation of simt_emu.h: lets the CPU test suite pin the device algorithm against the reference build.
// Test infrastructure; never loaded by the product path.
#include "simt_emu.h"

// [0] unique contender, [1] plain-pulse exact fallback, [2] RDO-pulse exact fallback, [3] literal scan (x32 lanes)
#define DAALA_B200_EMU_STATS
extern "C" { long long daala_b200_pvq_warp_stats[4]; }

#include "pvq_warp.cuh"

using namespace daala_b200::pvq;

extern "C" int emu_quantise_band(int32_t* out, const int32_t* x0, const int32_t* r0, int n, int q0, int32_t* y, int* itheta,
                                 int* max_theta, int* vk, int beta, double* skip_term, int is_keyframe, int pli,
                                 const int16_t* qm, const int16_t* qm_inv, double lambda) {
  int gain[32], it[32], mt[32], k[32];
  double sd[32];
  static int16_t snap[kSnapEntries];
  static double* rsq = nullptr;
  if (!rsq) {
    rsq = new double[kTableDoubles];
    for (int i = 0; i < kTableDoubles; i++) pvq_fill_rsqrt_table(rsq, i);
  }
  simt_emu::run_warp([&](int lane) {
    gain[lane] = quantise_band_warp(lane, snap, rsq, out, x0, r0, n, q0, y, &it[lane], &mt[lane], &k[lane], beta, &sd[lane],
                                    is_keyframe, pli, qm, qm_inv, lambda);
  });
  for (int l = 1; l < 32; l++) {
    if (gain[l] != gain[0] || it[l] != it[0] || mt[l] != mt[0] || k[l] != k[0] || sd[l] != sd[0]) {
      fprintf(stderr, "emu_quantise_band: lanes disagree on scalar results\n");
      abort();
    }
  }
  *itheta = it[0];
  *max_theta = mt[0];
  *vk = k[0];
  *skip_term = sd[0];
  return gain[0];
}
