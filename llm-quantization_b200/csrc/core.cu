// Library plumbing: error strings, launch accounting, torch-CPU log2 step tables.
#include <atomic>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200q {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    return fail(B200Q_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
  }
  return B200Q_OK;
}

// ---- per-entry-point CUDA-event profiler --------------------------------------------------------
// bench.py turns this on for the timed steps: every C-ABI call then brackets its launches with two
// events on the launching stream, so kernel durations are measured inside the real pipeline rather
// than in a separate replay.  Off by default (two relaxed loads per call).
struct ProfRec {
  const char* name;
  cudaEvent_t e0, e1;
  double bytes, flops;
};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

KernelScope::KernelScope(const char* name, double bytes, double flops, cudaStream_t st)
    : name_(name), bytes_(bytes), flops_(flops), st_(st), e0_(nullptr), e1_(nullptr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) {
    e0_ = e1_ = nullptr;
    cudaGetLastError();
    return;
  }
  cudaEventRecord(e0_, st_);
}
KernelScope::~KernelScope() {
  if (e0_ == nullptr) return;
  cudaEventRecord(e1_, st_);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back({name_, e0_, e1_, bytes_, flops_});
}

// ---- torch-CPU log2 semantics ------------------------------------------------------------------
// torch.log2 on CPU (SLEEF u10) agrees with "double log2 rounded to float" at every point where
// rne()/floor() of the result changes value (verified exhaustively around the steps by
// tests/test_torch_semantics.py).  rne(log2f(r)) is then a step function of r whose steps sit
// next to sqrt(2)*2^e, shifted by a few ulps because several floats map to exactly e+0.5 and the
// tie goes to the even integer; floor(log2f(m)) steps a few ulps BELOW 2^e because log2f rounds up
// to exactly e there.  We tabulate the first float of each step.
static float log2f_model(uint32_t bits) {
  float r;
  std::memcpy(&r, &bits, 4);
  return static_cast<float>(std::log2(static_cast<double>(r)));
}

template <typename Pred>
static uint32_t first_bits_where(Pred pred) {
  // smallest positive finite float bit pattern with pred true (pred monotone); inf bits if none
  uint32_t lo = 1, hi = 0x7f800000u;
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo) / 2;
    if (pred(mid)) hi = mid; else lo = mid + 1;
  }
  return lo;
}

static uint32_t g_round_thr[255];
static uint32_t g_floor_thr[277];
static std::once_flag g_tables_once;

static void build_tables() {
  for (int e = -127; e <= 127; ++e) {
    const float target = static_cast<float>(e + 1);
    g_round_thr[e + 127] =
        first_bits_where([&](uint32_t b) { return std::nearbyintf(log2f_model(b)) >= target; });
  }
  for (int e = -149; e <= 127; ++e) {
    const float target = static_cast<float>(e);
    g_floor_thr[e + 149] =
        first_bits_where([&](uint32_t b) { return std::floor(log2f_model(b)) >= target; });
  }
}

const uint32_t* log2_round_thresholds() {
  std::call_once(g_tables_once, build_tables);
  return g_round_thr;
}
const uint32_t* log2_floor_thresholds() {
  std::call_once(g_tables_once, build_tables);
  return g_floor_thr;
}

}  // namespace b200q

extern "C" {

const char* b200q_last_error(void) { return b200q::g_last_error.c_str(); }
int b200q_version(void) { return 100; }
int64_t b200q_launch_count(void) { return b200q::g_launches.load(); }

void b200q_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(b200q::g_prof_mu);
  if (on) {
    for (auto& r : b200q::g_prof) {
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    b200q::g_prof.clear();
  }
  b200q::g_prof_on.store(on != 0);
}

int b200q_profile_query(const char* name, double* total_ms, int64_t* launches, double* bytes,
                        double* flops) {
  std::lock_guard<std::mutex> lock(b200q::g_prof_mu);
  double ms = 0, by = 0, fl = 0;
  int64_t n = 0;
  for (auto& r : b200q::g_prof) {
    if (name != nullptr && std::strcmp(name, r.name) != 0) continue;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) return b200q::fail(B200Q_ECUDA, "profile: sync");
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess)
      return b200q::fail(B200Q_ECUDA, "profile: elapsed");
    ms += t; by += r.bytes; fl += r.flops; ++n;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = n;
  if (bytes) *bytes = by;
  if (flops) *flops = fl;
  return B200Q_OK;
}

uint32_t b200q_log2_round_threshold_bits(int e) {
  if (e < -127 || e > 127) return 0;
  return b200q::log2_round_thresholds()[e + 127];
}
uint32_t b200q_log2_floor_threshold_bits(int e) {
  if (e < -149 || e > 127) return 0;
  return b200q::log2_floor_thresholds()[e + 149];
}

}  // extern "C"
