// Library-level pieces of the C ABI: error string, version, device probe.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace simx {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- profiler -------------------------------------------------------------------------------
struct ProfRec {
  char name[48];
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
static thread_local const char* g_prof_label = nullptr;

void prof_label(const char* label) { g_prof_label = label; }

ProfScope::ProfScope(const char* default_name, double flops, double bytes, cudaStream_t s) : slot(-1), st(s) {
  if (!g_prof_on) return;
  ProfRec r;
  snprintf(r.name, sizeof(r.name), "%s", g_prof_label ? g_prof_label : default_name);
  r.flops = flops;
  r.bytes = bytes;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].e1, st);
}

}  // namespace simx

extern "C" int simx_profile_begin(void) {
  std::lock_guard<std::mutex> lk(simx::g_prof_mu);
  for (auto& r : simx::g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  simx::g_prof.clear();
  simx::g_prof_on = true;
  return SIMX_OK;
}

// Stops profiling, synchronises the device and writes a JSON array
// [{"name":..,"launches":n,"ms":t,"flops":f,"bytes":b}, ...] (aggregated by label, first-seen order) into buf.
extern "C" int simx_profile_end(char* buf, size_t cap) {
  using namespace simx;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) SIMX_FAIL(SIMX_ERR_CUDA, "profile_end: device synchronize failed");
  struct Agg { std::string name; long n; double ms, flops, bytes; };
  std::vector<Agg> aggs;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    Agg* a = nullptr;
    for (auto& x : aggs) if (x.name == r.name) { a = &x; break; }
    if (!a) { aggs.push_back(Agg{r.name, 0, 0, 0, 0}); a = &aggs.back(); }
    a->n += 1; a->ms += ms; a->flops += r.flops; a->bytes += r.bytes;
  }
  g_prof.clear();
  std::string out = "[";
  for (size_t i = 0; i < aggs.size(); ++i) {
    char line[256];
    snprintf(line, sizeof(line), "%s{\"name\":\"%s\",\"launches\":%ld,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}",
             i ? "," : "", aggs[i].name.c_str(), aggs[i].n, aggs[i].ms, aggs[i].flops, aggs[i].bytes);
    out += line;
  }
  out += "]";
  SIMX_CHECK(buf != nullptr && out.size() + 1 <= cap, SIMX_ERR_INVALID, "profile_end: buffer too small (%zu needed)",
             out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return SIMX_OK;
}

extern "C" const char* simx_last_error(void) { return simx::g_err; }
extern "C" int simx_version(void) { return 100; }
extern "C" int simx_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
