// Host emulation of the warp primitives the warp-per-band PVQ code uses (test infrastructure).
//
// 32 fibres (ucontext) run the same function, one per lane.  A warp collective deposits the lane's
// operand and yields to the scheduler; when every live fibre has arrived the scheduler computes the
// per-lane results and resumes them.  Only full-mask collectives are supported, which is all
// daala_b200/csrc/pvq_warp.cuh uses, and every lane must reach the same sequence of collectives
// (checked).
#pragma once
#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__ const

namespace simt_emu {

enum Op { kNone, kShfl, kAdd, kMax, kMin, kMaxU, kBallot, kAny };

struct Warp {
  ucontext_t main_ctx, ctx[32];
  char* stacks[32];
  bool done[32];
  int op[32];
  int64_t val[32];
  int arg[32];
  int64_t res[32];
  std::function<void(int)> body;
};

inline Warp*& cur_warp() { static thread_local Warp* w = nullptr; return w; }
inline int& cur_lane() { static thread_local int l = 0; return l; }

inline void trampoline() {
  Warp* w = cur_warp();
  const int l = cur_lane();
  w->body(l);
  w->done[l] = true;
  swapcontext(&w->ctx[l], &w->main_ctx);
}

inline int64_t collective(int op, int64_t v, int arg) {
  Warp* w = cur_warp();
  const int l = cur_lane();
  w->op[l] = op;
  w->val[l] = v;
  w->arg[l] = arg;
  swapcontext(&w->ctx[l], &w->main_ctx);
  return w->res[l];
}

inline void run_warp(std::function<void(int)> body) {
  Warp* w = new Warp;
  cur_warp() = w;
  w->body = body;
  const size_t kStack = 1 << 18;
  for (int l = 0; l < 32; l++) {
    w->done[l] = false;
    w->op[l] = kNone;
    w->stacks[l] = (char*)malloc(kStack);
    getcontext(&w->ctx[l]);
    w->ctx[l].uc_stack.ss_sp = w->stacks[l];
    w->ctx[l].uc_stack.ss_size = kStack;
    w->ctx[l].uc_link = &w->main_ctx;
    makecontext(&w->ctx[l], (void (*)())trampoline, 0);
  }
  for (;;) {
    int alive = 0;
    for (int l = 0; l < 32; l++) {
      if (w->done[l]) continue;
      cur_lane() = l;
      swapcontext(&w->main_ctx, &w->ctx[l]);
      if (!w->done[l]) alive++;
    }
    if (alive == 0) break;
    if (alive != 32) {
      fprintf(stderr, "simt_emu: %d lanes wait at a full-mask collective while others returned\n", alive);
      abort();
    }
    const int op = w->op[0];
    for (int l = 1; l < 32; l++) {
      if (w->op[l] != op) {
        fprintf(stderr, "simt_emu: lanes diverged at a collective (lane 0 op %d, lane %d op %d)\n", op, l, w->op[l]);
        abort();
      }
    }
    switch (op) {
      case kShfl:
        for (int l = 0; l < 32; l++) w->res[l] = w->val[w->arg[l] & 31];
        break;
      case kAdd: {
        uint32_t s = 0;
        for (int l = 0; l < 32; l++) s += (uint32_t)w->val[l];
        for (int l = 0; l < 32; l++) w->res[l] = (int32_t)s;
        break;
      }
      case kMax: {
        int32_t s = (int32_t)w->val[0];
        for (int l = 1; l < 32; l++) s = (int32_t)w->val[l] > s ? (int32_t)w->val[l] : s;
        for (int l = 0; l < 32; l++) w->res[l] = s;
        break;
      }
      case kMin: {
        int32_t s = (int32_t)w->val[0];
        for (int l = 1; l < 32; l++) s = (int32_t)w->val[l] < s ? (int32_t)w->val[l] : s;
        for (int l = 0; l < 32; l++) w->res[l] = s;
        break;
      }
      case kMaxU: {
        uint32_t s = (uint32_t)w->val[0];
        for (int l = 1; l < 32; l++) s = (uint32_t)w->val[l] > s ? (uint32_t)w->val[l] : s;
        for (int l = 0; l < 32; l++) w->res[l] = s;
        break;
      }
      case kBallot:
      case kAny: {
        uint32_t m = 0;
        for (int l = 0; l < 32; l++) if (w->val[l]) m |= 1u << l;
        for (int l = 0; l < 32; l++) w->res[l] = op == kAny ? (m != 0) : m;
        break;
      }
      default:
        abort();
    }
  }
  for (int l = 0; l < 32; l++) free(w->stacks[l]);
  delete w;
  cur_warp() = nullptr;
}

}  // namespace simt_emu

#define SIMT_FULL(m) assert((m) == 0xffffffffu)
inline int __shfl_sync(unsigned m, int v, int src) { SIMT_FULL(m); return (int)simt_emu::collective(simt_emu::kShfl, v, src); }
inline int __reduce_add_sync(unsigned m, int v) { SIMT_FULL(m); return (int)simt_emu::collective(simt_emu::kAdd, v, 0); }
inline int __reduce_max_sync(unsigned m, int v) { SIMT_FULL(m); return (int)simt_emu::collective(simt_emu::kMax, v, 0); }
inline int __reduce_min_sync(unsigned m, int v) { SIMT_FULL(m); return (int)simt_emu::collective(simt_emu::kMin, v, 0); }
inline unsigned __reduce_max_sync(unsigned m, unsigned v) { SIMT_FULL(m); return (unsigned)simt_emu::collective(simt_emu::kMaxU, v, 0); }
inline unsigned __ballot_sync(unsigned m, int p) { SIMT_FULL(m); return (unsigned)simt_emu::collective(simt_emu::kBallot, p != 0, 0); }
inline int __any_sync(unsigned m, int p) { SIMT_FULL(m); return (int)simt_emu::collective(simt_emu::kAny, p != 0, 0); }
inline int __double2loint(double d) { int64_t b; memcpy(&b, &d, 8); return (int)(b & 0xffffffff); }
inline int __double2hiint(double d) { int64_t b; memcpy(&b, &d, 8); return (int)(b >> 32); }
inline double __hiloint2double(int hi, int lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double d; memcpy(&d, &b, 8); return d; }
inline int __float_as_int(float f) { int b; memcpy(&b, &f, 4); return b; }
inline unsigned __float_as_uint(float f) { unsigned b; memcpy(&b, &f, 4); return b; }
inline float __fdividef(float a, float b) { return a / b; }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
