// Library-level plumbing of the C ABI: version, thread-local error text, device queries.
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"

namespace mccb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// Launch-configuration errors only (cudaPeekAtLastError does not synchronise, so the
// call stays asynchronous and graph-capturable).
int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(MCCB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return MCCB200_OK;
}

// ---- optional phase timing (bench.py roofline): events recorded on the launching stream
struct PhaseMark { const char* name; cudaEvent_t ev; };
static bool g_prof_on = false;
static PhaseMark g_marks[4096];
static int g_n_marks = 0;

void phase_mark(const char* name, cudaStream_t st) {
  if (!g_prof_on || g_n_marks >= 4096) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  g_marks[g_n_marks++] = {name, ev};
}

static unsigned long long g_launches = 0;
void count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static thread_local int g_side_ctas = 0;
int side_ctas() { return g_side_ctas; }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace mccb

extern "C" void mccb200_set_side_residency(int ctas_per_sm) {
  mccb::g_side_ctas = ctas_per_sm < 0 ? 0 : (ctas_per_sm > 8 ? 8 : ctas_per_sm);
}
extern "C" int mccb200_version(void) { return MCCB200_VERSION; }
extern "C" const char* mccb200_last_error(void) { return mccb::g_err; }
extern "C" int mccb200_sm_count(void) { return mccb::sm_count(); }

extern "C" void mccb200_profile_enable(int on) {
  for (int i = 0; i < mccb::g_n_marks; ++i) cudaEventDestroy(mccb::g_marks[i].ev);
  mccb::g_n_marks = 0;
  mccb::g_prof_on = on != 0;
}

// Duration of phase i = time from its mark to the next mark ("end" marks close a call).
// Synchronises on the last recorded event.  Returns the number of phases written.
extern "C" int mccb200_profile_read(const char** names, float* ms, int max_entries) {
  using namespace mccb;
  if (g_n_marks < 2) return 0;
  cudaEventSynchronize(g_marks[g_n_marks - 1].ev);
  int out = 0;
  for (int i = 0; i + 1 < g_n_marks && out < max_entries; ++i) {
    if (strcmp(g_marks[i].name, "end") == 0) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, g_marks[i].ev, g_marks[i + 1].ev) != cudaSuccess) continue;
    names[out] = g_marks[i].name;
    ms[out] = t;
    ++out;
  }
  return out;
}
extern "C" unsigned long long mccb200_launch_count(void) { return __atomic_load_n(&mccb::g_launches, __ATOMIC_RELAXED); }
