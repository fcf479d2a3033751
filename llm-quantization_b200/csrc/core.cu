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
int64_t launches_so_far() { return g_launches.load(std::memory_order_relaxed); }
static thread_local bool t_capture = false;
void set_stream_capture(bool on) { t_capture = on; }
bool in_stream_capture() { return t_capture; }
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
  if (!g_prof_on.load(std::memory_order_relaxed) || t_capture) return;
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

// For fp16 / bf16 tensors torch evaluates log2 in fp32 and rounds the result to the tensor's type
// before round()/floor() see it, and the argument itself is a 16-bit value: the steps sit
// elsewhere.  Their tables are built by walking every positive finite value of the type.
static float round_to_bf16(float f) {
  uint32_t b;
  std::memcpy(&b, &f, 4);
  if ((b & 0x7f800000u) == 0x7f800000u) return f;            // inf / nan
  b += 0x7fffu + ((b >> 16) & 1u);                             // round to nearest even
  b &= 0xffff0000u;
  std::memcpy(&f, &b, 4);
  return f;
}
static float f16_bits_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  const uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu;
  float v;
  if (exp == 0) v = std::ldexp((float)man, -24);
  else if (exp == 31) v = man ? NAN : INFINITY;
  else v = std::ldexp((float)(man | 0x400u), (int)exp - 25);
  uint32_t b;
  std::memcpy(&b, &v, 4);
  b |= sign;
  std::memcpy(&v, &b, 4);
  return v;
}
static float round_to_f16(float f) {
  if (!(std::fabs(f) <= 3.4e38f) || f == 0.f) return f;       // nan / inf / zero
  const float a = std::fabs(f);
  if (a >= 65520.f) return std::copysign(INFINITY, f);         // rounds past the largest half
  int e;
  std::frexp(a, &e);                                           // a = m * 2^e, m in [0.5, 1)
  const int ulp_exp = std::max(e - 11, -24);                   // 11 significant bits, subnormal floor
  const float q = std::nearbyint(std::ldexp(a, -ulp_exp));     // ties to even (default mode)
  return std::copysign(std::ldexp(q, ulp_exp), f);
}

static uint32_t g_round_thr[3][255];
static uint32_t g_floor_thr[3][277];
static std::once_flag g_tables_once;

static void build_tables() {
  for (int e = -127; e <= 127; ++e) {
    const float target = static_cast<float>(e + 1);
    g_round_thr[0][e + 127] =
        first_bits_where([&](uint32_t b) { return std::nearbyintf(log2f_model(b)) >= target; });
  }
  for (int e = -149; e <= 127; ++e) {
    const float target = static_cast<float>(e);
    g_floor_thr[0][e + 149] =
        first_bits_where([&](uint32_t b) { return std::floor(log2f_model(b)) >= target; });
  }
  // 16-bit types: one pass over all positive finite values in increasing order
  for (int dt = 1; dt <= 2; ++dt) {
    for (int i = 0; i < 255; ++i) g_round_thr[dt][i] = 0x7f800000u;
    for (int i = 0; i < 277; ++i) g_floor_thr[dt][i] = 0x7f800000u;
    const uint32_t last = dt == 1 ? 0x7bffu : 0x7f7fu;
    for (uint32_t v = 1; v <= last; ++v) {
      const float r = dt == 1 ? f16_bits_to_f32((uint16_t)v) : [&]() {
        uint32_t b = v << 16; float f; std::memcpy(&f, &b, 4); return f; }();
      uint32_t rbits;
      std::memcpy(&rbits, &r, 4);
      float y = log2f_model(rbits);
      y = dt == 1 ? round_to_f16(y) : round_to_bf16(y);
      const int ri = (int)std::nearbyintf(y), fi = (int)std::floor(y);
      // first value whose rounded log2 reaches e+1 / whose floored log2 reaches e, for every e
      // not yet claimed (values come in increasing order, the results are monotone)
      for (int e = ri - 1; e >= -127 && e <= 127 && g_round_thr[dt][e + 127] == 0x7f800000u; --e)
        g_round_thr[dt][e + 127] = rbits;
      for (int e = fi; e >= -149 && e <= 127 && g_floor_thr[dt][e + 149] == 0x7f800000u; --e)
        g_floor_thr[dt][e + 149] = rbits;
    }
  }
}

const uint32_t* log2_round_thresholds(int dtype) {
  std::call_once(g_tables_once, build_tables);
  return g_round_thr[dtype];
}
const uint32_t* log2_floor_thresholds(int dtype) {
  std::call_once(g_tables_once, build_tables);
  return g_floor_thr[dtype];
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

uint32_t b200q_log2_round_threshold_bits(int e) { return b200q_log2_round_threshold_bits_dt(e, 0); }
uint32_t b200q_log2_floor_threshold_bits(int e) { return b200q_log2_floor_threshold_bits_dt(e, 0); }
uint32_t b200q_log2_round_threshold_bits_dt(int e, int dtype) {
  if (e < -127 || e > 127 || dtype < 0 || dtype > 2) return 0;
  return b200q::log2_round_thresholds(dtype)[e + 127];
}
uint32_t b200q_log2_floor_threshold_bits_dt(int e, int dtype) {
  if (e < -149 || e > 127 || dtype < 0 || dtype > 2) return 0;
  return b200q::log2_floor_thresholds(dtype)[e + 149];
}

}  // extern "C"
