// Host stand-ins for the CUDA constructs used by generated pass units (TEST INFRASTRUCTURE: compiled only
// by tests/_jit_emu.py with g++ -DQLS_HOST_EMU; nothing in the product path includes this file).
// A consumer group is emulated by 256 real host threads and a pthread barrier, so the generated
// straight-line code (with its barriers) runs unmodified.
#ifndef QLS_RS2_HOST_SHIM_H
#define QLS_RS2_HOST_SHIM_H

#include <pthread.h>

#include <cmath>
#include <thread>
#include <vector>

typedef unsigned long long u64;
typedef unsigned int u32;
struct double2 {
  double x, y;
};
#define QLS_DEV static inline
#define QLS_CONST static const
#define QLS_UNIFORM(x) (x)
#define QLS_FRESH_TID(k) (tid_in | (u32)(gbase >> (42 + (k))))
template <class T>
static inline T __ldg(const T* p) {
  return *p;
}
using std::fma;

static pthread_barrier_t qls_emu_bar;
#define QLS_BAR() pthread_barrier_wait(&qls_emu_bar)
#define QLS_BODY_BAR() QLS_BAR()

template <class Copy, class Body>
static void qls_emu_run(u64 n_tiles, Copy copy, Body body) {
  const unsigned n_threads = 256;
  pthread_barrier_init(&qls_emu_bar, nullptr, n_threads);
  std::vector<std::thread> th;
  for (unsigned tid = 0; tid < n_threads; ++tid) {
    th.emplace_back([&, tid] {
      for (u64 tile = 0; tile < n_tiles; ++tile) {
        if (tid == 0) copy(tile, false);
        pthread_barrier_wait(&qls_emu_bar);
        body(tile, tid);
        pthread_barrier_wait(&qls_emu_bar);
        if (tid == 0) copy(tile, true);
        pthread_barrier_wait(&qls_emu_bar);
      }
    });
  }
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&qls_emu_bar);
}

#endif
