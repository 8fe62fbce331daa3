// Launch counter and optional per-call-site kernel timing (see sard_prof_* in sard_b200.h).
#include "common.cuh"
#include <mutex>
#include <string>
#include <vector>

long long g_sard_launches = 0;
int g_sard_pdl = 1;
extern "C" int sard_set_pdl(int enabled) { g_sard_pdl = enabled ? 1 : 0; return 0; }
extern "C" int sard_get_pdl(void) { return g_sard_pdl; }
thread_local int g_sard_site = -1;

namespace {
struct Rec { int site; double flops, bytes; cudaEvent_t a, b; cudaStream_t st; };
unsigned g_mask = 0;
int g_capacity = 0;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
std::mutex g_prof_mu;          // the critic and the policy may be issued from two host threads
thread_local double g_pending_bytes = 0.0, g_pending_flops = 0.0;     // work figures of the next launch (per-kernel table)
}  // namespace

ProfScope::ProfScope(cudaStream_t stream, double flops, double bytes) : slot(-1), st(stream) {
  if (g_sard_prof_all) { g_pending_bytes = bytes; g_pending_flops = flops; }      // the per-kernel table gets the same work figures
  const int site = g_sard_site;
  if (g_mask == 0 || site < 0 || site >= SARD_PROF_SITES || !((g_mask >> site) & 1u)) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if ((int)g_recs.size() >= g_capacity) return;
  Rec r;
  r.site = site; r.flops = flops; r.bytes = bytes; r.st = stream;
  for (cudaEvent_t* e : {&r.a, &r.b}) {
    if (!g_pool.empty()) { *e = g_pool.back(); g_pool.pop_back(); }
    else if (cudaEventCreate(e) != cudaSuccess) return;
  }
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  slot = (int)g_recs.size() - 1;
}

ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  cudaEventRecord(g_recs[slot].b, st);
}

extern "C" long long sard_launch_count(void) { return g_sard_launches; }

// ---------------------------------------------------------------------------------------- every kernel, by name
int g_sard_prof_all = 0;
namespace {
struct KRec { int name; double bytes, flops; cudaEvent_t a, b; cudaStream_t st; };
std::vector<KRec> g_krecs;
std::vector<std::string> g_knames;
int g_kcapacity = 0;
thread_local int g_kopen = -1;           // index of this thread's record awaiting its end event / label
cudaEvent_t pool_event() {
  cudaEvent_t e = nullptr;
  if (!g_pool.empty()) { e = g_pool.back(); g_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
}  // namespace

void sard_prof_bytes(double bytes) { if (g_sard_prof_all) g_pending_bytes = bytes; }
void sard_prof_pre(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_kopen = -1;
  if ((int)g_krecs.size() >= g_kcapacity) return;
  KRec r;
  r.name = -1; r.bytes = g_pending_bytes; r.flops = g_pending_flops; r.st = st; r.a = pool_event(); r.b = pool_event();
  g_pending_bytes = 0.0; g_pending_flops = 0.0;
  cudaEventRecord(r.a, st);
  g_krecs.push_back(r);
  g_kopen = (int)g_krecs.size() - 1;
}
void sard_prof_post(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (g_kopen >= 0) cudaEventRecord(g_krecs[g_kopen].b, st);
}
void sard_prof_label(const char* name) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (g_kopen < 0) return;
  const int k = g_kopen;
  g_kopen = -1;
  for (size_t i = 0; i < g_knames.size(); ++i)
    if (g_knames[i] == name) { g_krecs[k].name = (int)i; return; }
  g_knames.push_back(name);
  g_krecs[k].name = (int)g_knames.size() - 1;
}

extern "C" int sard_prof_kernels_enable(int capacity) {
  g_sard_prof_all = capacity > 0 ? 1 : 0;
  g_kcapacity = capacity;
  if (capacity > 0) g_krecs.reserve(capacity);
  return 0;
}
// names: '\n'-separated kernel names in id order (truncated to names_len); rows[i*6..] = {name id, stream ordinal, start ms,
// end ms, algorithmic bytes, algorithmic flops} relative to the first record.  Returns the number of rows and clears the records.
extern "C" int sard_prof_kernels_collect(char* names, int names_len, double* rows, int capacity) {
  std::string all;
  for (const std::string& n : g_knames) { all += n; all += '\n'; }
  if (names && names_len > 0) { snprintf(names, (size_t)names_len, "%s", all.c_str()); }
  std::vector<cudaStream_t> streams;
  int n = 0;
  for (const KRec& r : g_krecs) {
    cudaEventSynchronize(r.b);
    if (n < capacity && r.name >= 0) {
      float t0 = 0.f, t1 = 0.f;
      cudaEventElapsedTime(&t0, g_krecs[0].a, r.a);
      cudaEventElapsedTime(&t1, g_krecs[0].a, r.b);
      int sid = -1;
      for (size_t k = 0; k < streams.size(); ++k) if (streams[k] == r.st) sid = (int)k;
      if (sid < 0) { streams.push_back(r.st); sid = (int)streams.size() - 1; }
      rows[n * 6 + 0] = r.name; rows[n * 6 + 1] = sid; rows[n * 6 + 2] = t0; rows[n * 6 + 3] = t1; rows[n * 6 + 4] = r.bytes; rows[n * 6 + 5] = r.flops;
      ++n;
    }
  }
  for (const KRec& r : g_krecs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_krecs.clear();
  return n;
}

extern "C" int sard_prof_enable(unsigned mask, int capacity) {
  g_mask = mask;
  g_capacity = capacity;
  if (capacity > 0) g_recs.reserve(capacity);
  return 0;
}

// rows[i*4 ..] = {site, stream ordinal (order of first use), start ms, end ms} relative to the first record;
// does not clear the records (call sard_prof_collect afterwards)
extern "C" int sard_prof_timeline(double* rows, int capacity) {
  std::vector<cudaStream_t> streams;
  int n = 0;
  for (const Rec& r : g_recs) {
    if (n >= capacity) break;
    SARD_CUDA(cudaEventSynchronize(r.b));
    float t0 = 0.f, t1 = 0.f;
    SARD_CUDA(cudaEventElapsedTime(&t0, g_recs[0].a, r.a));
    SARD_CUDA(cudaEventElapsedTime(&t1, g_recs[0].a, r.b));
    int sid = -1;
    for (size_t k = 0; k < streams.size(); ++k) if (streams[k] == r.st) sid = (int)k;
    if (sid < 0) { streams.push_back(r.st); sid = (int)streams.size() - 1; }
    rows[n * 4 + 0] = r.site; rows[n * 4 + 1] = sid; rows[n * 4 + 2] = t0; rows[n * 4 + 3] = t1;
    ++n;
  }
  return n < 0 ? 0 : n;
}

extern "C" int sard_prof_collect(double* rows, int n_sites) {
  for (int i = 0; i < n_sites * 4; ++i) rows[i] = 0.0;
  for (const Rec& r : g_recs) {
    SARD_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    SARD_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.site < n_sites) {
      rows[r.site * 4 + 0] += 1.0;
      rows[r.site * 4 + 1] += (double)ms;
      rows[r.site * 4 + 2] += r.flops;
      rows[r.site * 4 + 3] += r.bytes;
    }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  return 0;
}
