// Library-level C-ABI entry points (include/cosa_b200.h): version, error text, launch counter and the
// optional per-kernel event timing used by bench.py for the roofline numbers.
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace cosa {

unsigned long long g_launches = 0;
bool g_prof_on = false;

struct ProfRec {
  const char *name;
  cudaEvent_t start, stop;
};
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t take_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void prof_mark(const char *name, cudaStream_t stream, bool is_start) {
  if (is_start) {
    ProfRec r{name, take_event(), take_event()};
    cudaEventRecord(r.start, stream);
    g_prof_recs.push_back(r);
  } else if (!g_prof_recs.empty()) {
    cudaEventRecord(g_prof_recs.back().stop, stream);
  }
}

}  // namespace cosa

using namespace cosa;

extern "C" int cosa_abi_version(void) { return COSA_B200_ABI_VERSION; }

extern "C" unsigned long long cosa_launch_count(void) { return g_launches; }

extern "C" const char *cosa_strerror(int code) {
  switch (code) {
    case COSA_OK: return "ok";
    case COSA_E_ARG: return "invalid argument (shape, null pointer or unsupported parameter)";
    case COSA_E_WORKSPACE: return "workspace too small";
    case COSA_E_KEYRANGE: return "lattice coordinate outside the packed-key range";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown cosa_b200 error";
  }
}

extern "C" void cosa_profile_begin(void) {
  for (auto &r : g_prof_recs) { g_event_pool.push_back(r.start); g_event_pool.push_back(r.stop); }
  g_prof_recs.clear();
  g_prof_on = true;
}

// Stops profiling, synchronises the device and writes one line per kernel, "name count total_ms\n",
// into buf (truncated to buf_len).  Returns the number of distinct kernels.
extern "C" int cosa_profile_end(char *buf, size_t buf_len) {
  g_prof_on = false;
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long long, double>> acc;
  for (auto &r : g_prof_recs) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      std::string n(r.name);
      const size_t lt = n.find('<');
      if (lt != std::string::npos) n = n.substr(0, lt);
      auto &a = acc[n];
      a.first += 1;
      a.second += ms;
    }
    g_event_pool.push_back(r.start);
    g_event_pool.push_back(r.stop);
  }
  g_prof_recs.clear();
  size_t off = 0;
  if (buf && buf_len) buf[0] = 0;
  for (auto &kv : acc) {
    char line[256];
    const int len = snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    if (buf && off + (size_t)len + 1 < buf_len) {
      memcpy(buf + off, line, (size_t)len + 1);
      off += (size_t)len;
    }
  }
  return (int)acc.size();
}
