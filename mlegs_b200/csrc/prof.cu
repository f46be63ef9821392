// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
#include <cstdio>
#include <cstring>
#include <map>

#include "kernels.h"

namespace mlegs {

struct ProfRec {
  const char *name;
  cudaEvent_t a, b;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void prof_begin(const char *name, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfRec r;
  r.name = name;
  r.a = get_event();
  r.b = get_event();
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
}

void prof_end(cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_recs.back().b, st);
}

}  // namespace mlegs

using namespace mlegs;

extern "C" {

int mlegs_b200_prof_enable(int on) {
  g_prof_on = on != 0;
  return MLEGS_OK;
}

// Writes a JSON object {"kernel": {"launches": n, "ms": total}, ...} and clears the records.
int mlegs_b200_prof_report(char *buf, size_t nbuf) {
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)ctx().stream));
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto &r : g_recs) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    auto &p = agg[r.name];
    p.first += 1;
    p.second += ms;
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  std::string out = "{";
  bool first = true;
  for (auto &kv : agg) {
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    out += tmp;
    first = false;
  }
  out += "}";
  if (out.size() + 1 > nbuf) return fail(MLEGS_E_ARG, "mlegs_b200_prof_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return MLEGS_OK;
}

}  // extern "C"
