// SIMT emulation for host builds of the warp-cooperative kernels (TEST INFRASTRUCTURE, see oracle/__init__.py).
//
// A "group" of n lanes (a warp: 32, a thread block: up to 1024) runs as n cooperative fibers (ucontext) on ONE
// OS thread.  WSYNC()/BSYNC() switch to the next fiber round-robin; because every lane of a convergent kernel calls
// them the same number of times, a lane resumes after a barrier only when all other lanes have reached it -- barrier
// semantics without any OS synchronisation.  Reductions and broadcasts go through a per-group scratch array.
// Bulk copies (StageFeed) complete immediately.  Races between lanes cannot be detected here (the interleaving is
// deterministic); compute-sanitizer racecheck on the GPU does that job.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {
struct Group {
  int n = 0, cur = 0, done = 0;
  std::vector<ucontext_t> ctx;
  ucontext_t main;
  std::vector<char> stacks;
  std::function<void(int)> body;
  double slot[1024];
};
inline thread_local Group* g_grp = nullptr;
inline int lane_id() { return g_grp->cur; }
inline void sync() {
  Group* g = g_grp;
  if (g->n == 1) return;
  const int me = g->cur, nx = (me + 1) % g->n;
  g->cur = nx;
  swapcontext(&g->ctx[me], &g->ctx[nx]);
}
inline void trampoline(int i) {
  Group* g = g_grp;
  g->body(i);
  g->done++;
  if (g->done < g->n) {
    const int nx = (i + 1) % g->n;
    g->cur = nx;
    setcontext(&g->ctx[nx]);
  }
  setcontext(&g->main);
}
// run body(lane) for lanes 0..n-1 as one convergent group
inline void run(int n, const std::function<void(int)>& body, size_t stack_bytes = 1 << 20) {
  Group g;
  g.n = n;
  g.body = body;
  g.ctx.resize(n);
  g.stacks.resize((size_t)n * stack_bytes);
  Group* prev = g_grp;
  g_grp = &g;
  for (int i = 0; i < n; ++i) {
    getcontext(&g.ctx[i]);
    g.ctx[i].uc_stack.ss_sp = g.stacks.data() + (size_t)i * stack_bytes;
    g.ctx[i].uc_stack.ss_size = stack_bytes;
    g.ctx[i].uc_link = &g.main;
    makecontext(&g.ctx[i], (void (*)())trampoline, 1, i);
  }
  g.cur = 0;
  swapcontext(&g.main, &g.ctx[0]);
  g_grp = prev;
}
}  // namespace simt

#define CH_DEV inline
#define WSYNC() simt::sync()
#define BSYNC() simt::sync()

namespace rlmpc {
inline double wsum(double v) {
  simt::Group* g = simt::g_grp;
  g->slot[g->cur] = v;
  simt::sync();
  double s = 0.0;
  for (int i = 0; i < 32; ++i) s += g->slot[i];
  simt::sync();
  return s;
}
inline double wmax(double v) {
  simt::Group* g = simt::g_grp;
  g->slot[g->cur] = v;
  simt::sync();
  double s = g->slot[0];
  for (int i = 1; i < 32; ++i) s = std::fmax(s, g->slot[i]);
  simt::sync();
  return s;
}
inline double wmin(double v) {
  simt::Group* g = simt::g_grp;
  g->slot[g->cur] = v;
  simt::sync();
  double s = g->slot[0];
  for (int i = 1; i < 32; ++i) s = std::fmin(s, g->slot[i]);
  simt::sync();
  return s;
}
inline bool wany(bool p) { return wmax(p ? 1.0 : 0.0) > 0.5; }
inline int atomic_next(int* counter) { return (*counter)++; }
// emulation of the m8n8k4 FP64 tensor-core product: fragments are exchanged through the group's scratch
inline void dmma_8x8x4(double a, double b, double& c0, double& c1) {
  simt::Group* g = simt::g_grp;
  const int l = g->cur;
  g->slot[l] = a;
  g->slot[32 + l] = b;
  simt::sync();
  const int r = l / 4, n0 = 2 * (l % 4);
  for (int k = 0; k < 4; ++k) {
    c0 += g->slot[4 * r + k] * g->slot[32 + 4 * n0 + k];
    c1 += g->slot[4 * r + k] * g->slot[32 + 4 * (n0 + 1) + k];
  }
  simt::sync();
}

struct StageFeed {
  double* buf[2];
  void init(double* b0, double* b1, uint64_t*, int) { buf[0] = b0; buf[1] = b1; }
  void publish(int) { simt::sync(); }
  void issue(int slot, int lane, const double* s0, int o0, int n0, const double* s1 = nullptr, int o1 = 0, int n1 = 0,
             const double* s2 = nullptr, int o2 = 0, int n2 = 0) {
    if (lane != 0) return;
    std::memcpy(buf[slot] + o0, s0, sizeof(double) * n0);
    if (n1 > 0) std::memcpy(buf[slot] + o1, s1, sizeof(double) * n1);
    if (n2 > 0) std::memcpy(buf[slot] + o2, s2, sizeof(double) * n2);
  }
  // the copy of lane 0 happened at issue time; lanes that run before lane 0 in the round-robin would read stale
  // data, so a wait is a barrier here
  const double* wait(int slot) { simt::sync(); return buf[slot]; }
};
}  // namespace rlmpc
