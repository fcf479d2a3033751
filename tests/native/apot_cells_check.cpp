// CPU check of llm-quantization_b200/csrc/apot_cells.h (test infrastructure, built by
// tests/test_apot_cells.py with g++): the cell-table lookup the APOT kernel evaluates per element
// must return the index torch.argmin(|x - levels|) returns (pot_apot_quantizer.py:294-297: fp32
// distances, first minimum; all-NaN distances -> index 0) for NaN and for every x with
// |x| <= 1e4.  (The kernels see x = w / s_b with |x| <= 1 / b_min = 100 for the reference's grids;
// far beyond that -- |x| > 2^23 x the level spacing -- every rounded distance ties and argmin
// degenerates to index 0, which neither the bisecting kernel nor the table reproduces.)
//
//   apot_cells_check <levels.bin> [--exhaustive]
// levels.bin: int32 n_sets, then per set int32 n_levels + n_levels float32 (ascending).
// Prints one line per set: "set <k> levels <n> eligible <0|1> checked <points> mismatches <m>".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "apot_cells.h"

using namespace b200q;

static int literal_argmin(const float* lv, int n, float x) {
  int best = 0;
  volatile float d = x - lv[0];
  float bd = fabsf(d);
  if (bd != bd) return 0;                       // NaN distance: torch.argmin returns the first NaN
  for (int i = 1; i < n; ++i) {
    d = x - lv[i];
    const float di = fabsf(d);
    if (di < bd) { bd = di; best = i; }
  }
  return best;
}

static float from_bits(uint32_t b) { float x; memcpy(&x, &b, 4); return x; }

struct Checker {
  const float* lv; int n; const ApotCells* cells;
  long long checked = 0, bad = 0;
  void point(float x) {
    if (x == x && !(fabsf(x) <= 1e4f)) return;  // outside the pinned domain (see the header)
    ++checked;
    const int got = apot_lookup(*cells, x), want = literal_argmin(lv, n, x);
    if (got != want) {
      if (bad < 10) fprintf(stderr, "  x=%.9g (0x%08x): table %d, argmin %d\n", x, apot_float_bits(x), got, want);
      ++bad;
    }
  }
  void around(float x, int ulps) {
    const int32_t k = apot_key(x);
    for (int d = -ulps; d <= ulps; ++d) point(apot_unkey(k + d));
  }
};

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rng() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double uni() { return (rng() >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s levels.bin [--exhaustive]\n", argv[0]); return 2; }
  const bool exhaustive = argc > 2 && !strcmp(argv[2], "--exhaustive");
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("levels.bin"); return 2; }
  int32_t n_sets = 0;
  if (fread(&n_sets, 4, 1, f) != 1) return 2;
  long long total_bad = 0;
  for (int s = 0; s < n_sets; ++s) {
    int32_t n = 0;
    if (fread(&n, 4, 1, f) != 1 || n < 1 || n > 64) return 2;
    std::vector<float> lv(n);
    if (fread(lv.data(), 4, n, f) != (size_t)n) return 2;
    ApotCells cells;
    const bool ok = apot_build_cells(lv.data(), n, cells);
    Checker c{lv.data(), n, &cells};
    if (ok) {
      for (int i = 0; i < cells.n_thr; ++i) c.around(cells.thr[i], 300);
      for (int i = 0; i < n; ++i) c.around(lv[i], 300);
      for (int k = -kApotCells / 2 - 2; k <= kApotCells / 2 + 2; ++k)        // cell boundaries
        c.around((float)(((double)k + 0.5) / (double)cells.scale), 300);
      c.around(cells.R, 300); c.around(-cells.R, 300); c.around(0.f, 300);
      const float special[] = {0.f, -0.f, 1e-45f, -1e-45f, 1.17549435e-38f, -1.17549435e-38f, 1e4f, -1e4f,
                               100.f, -100.f, 101.f, -101.f, NAN, -NAN};
      for (float x : special) c.point(x);
      for (int i = 0; i < 4000000; ++i) c.point((float)((uni() * 3.0 - 1.5) * cells.R));
      for (int i = 0; i < 1000000; ++i) {
        const double mag = exp(log(1e-30) + uni() * (log(1e3) - log(1e-30)));
        c.point((float)((rng() & 1) ? mag : -mag));
      }
      if (exhaustive) {
        // every fp32 bit pattern
        for (uint64_t b = 0; b <= 0xffffffffull; ++b) c.point(from_bits((uint32_t)b));
      }
    }
    printf("set %d levels %d eligible %d checked %lld mismatches %lld\n", s, n, ok ? 1 : 0, c.checked, c.bad);
    total_bad += c.bad;
  }
  fclose(f);
  return total_bad == 0 ? 0 : 1;
}
