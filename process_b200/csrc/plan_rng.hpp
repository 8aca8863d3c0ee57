// plan_rng.hpp -- the planner's random numbers.
//
// The planner splits the templates of a (sample, chromosome) over its tiles with chains of binomials: tens of
// thousands of Binomial(n, p) draws per call, in front of the first kernel launch.  Two things made that slow with
// the standard library: seeding a std::mt19937_64 from a std::seed_seq costs ~12 us per stream (one stream per 64
// tiles), and std::binomial_distribution ~0.3-0.6 us per draw.  Here:
//
//   PlanRng     a counter-based stream -- Philox4x32-10, the generator the kernels use -- keyed by what the stream
//               belongs to (seed, tag | sample, chromosome, block): nothing to seed, a stream starts in ~20 ns
//   binomial()  an exact Binomial(n, p) sampler: sequential inversion when n * min(p, 1 - p) < 10, else
//               Hoermann's transformed rejection with squeeze (BTRS, "The generation of binomial random
//               variates", J. Stat. Comput. Simul. 46, 1993): ~1.2 uniform pairs per draw, the acceptance test is
//               the exact ratio of probabilities (log form, Stirling tail corrections), so the law is Binomial(n, p)
//               up to double rounding.  tests/test_host_logic.py checks it against scipy's pmf (chi-square).
#pragma once
#include <cmath>
#include <cstdint>

namespace pcs {

struct PlanRng {
  uint32_t k0, k1;      // key: (seed, tag)
  uint32_t c1, c2, c3;  // counter words 1..3: what the stream belongs to
  uint32_t n = 0;       // counter word 0: blocks drawn so far
  uint32_t buf[4];
  int have = 0;

  PlanRng(uint32_t seed, uint32_t tag, uint32_t a, uint32_t b, uint32_t c) : k0(seed), k1(tag), c1(a), c2(b), c3(c) {}

  void refill() {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    uint32_t x0 = n++, x1 = c1, x2 = c2, x3 = c3, ka = k0, kb = k1;
    for (int r = 0; r < 10; ++r) {
      const uint64_t p0 = static_cast<uint64_t>(M0) * x0, p1 = static_cast<uint64_t>(M1) * x2;
      const uint32_t y0 = static_cast<uint32_t>(p1 >> 32) ^ x1 ^ ka, y1 = static_cast<uint32_t>(p1);
      const uint32_t y2 = static_cast<uint32_t>(p0 >> 32) ^ x3 ^ kb, y3 = static_cast<uint32_t>(p0);
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      ka += W0;
      kb += W1;
    }
    buf[0] = x0; buf[1] = x1; buf[2] = x2; buf[3] = x3;
    have = 4;
  }
  uint32_t u32() {
    if (have == 0) refill();
    return buf[--have];
  }
  // uniform on (0, 1): 53 random bits, never 0 or 1
  double uniform() {
    const uint64_t hi = u32(), lo = u32();
    const uint64_t bits = ((hi << 32) | lo) >> 11;
    return (static_cast<double>(bits) + 0.5) * (1.0 / 9007199254740992.0);
  }
};

namespace detail {

// log(k!) - [log(sqrt(2 pi)) + (k + 1/2) log(k + 1) - (k + 1)]: the tail of Stirling's series
inline double stirling_tail(double k) {
  static const double table[10] = {0.0810614667953272,  0.0413406959554092,  0.0276779256849983,  0.02079067210376509,
                                   0.0166446911898211,  0.0138761288230707,  0.0118967099458917,  0.0104112652619720,
                                   0.00925546218271273, 0.00833056343336287};
  if (k <= 9.0) return table[static_cast<int>(k)];
  const double kp1sq = (k + 1.0) * (k + 1.0);
  return (1.0 / 12 - (1.0 / 360 - 1.0 / 1260 / kp1sq) / kp1sq) / (k + 1.0);
}

// n * p < 10, p <= 1/2: walk the cumulative probabilities up from 0 (pmf(k+1) / pmf(k) = (n - k) / (k + 1) * p / q)
inline uint64_t binomial_inversion(PlanRng& rng, uint64_t n, double p) {
  const double q = 1.0 - p, s = p / q, a = (static_cast<double>(n) + 1.0) * s;
  const double f0 = std::exp(static_cast<double>(n) * std::log1p(-p));  // >= e^-10 (1 - p)^...: far from underflow
  for (;;) {
    double u = rng.uniform(), f = f0;
    uint64_t k = 0;
    // the mass beyond k = 200 is < 1e-150 for a mean below 10: a draw that gets there met rounding, try again
    while (u > f && k < 200 && k < n) {
      u -= f;
      ++k;
      f *= a / static_cast<double>(k) - s;
    }
    if (u <= f || k == n) return k;
  }
}

// n * p >= 10, p <= 1/2
inline uint64_t binomial_btrs(PlanRng& rng, uint64_t n_, double p) {
  const double n = static_cast<double>(n_);
  const double q = 1.0 - p, spq = std::sqrt(n * p * q);
  const double b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * p, c = n * p + 0.5;
  const double vr = 0.92 - 4.2 / b, r = p / q, alpha = (2.83 + 5.1 / b) * spq;
  const double m = std::floor((n + 1.0) * p);
  double h = 0.0;
  bool have_h = false;
  for (;;) {
    const double u = rng.uniform() - 0.5;
    double v = rng.uniform();
    const double us = 0.5 - std::fabs(u);
    const double k = std::floor((2.0 * a / us + b) * u + c);
    if (us >= 0.07 && v <= vr) return static_cast<uint64_t>(k);  // inside the box under the hat: accept at once
    if (k < 0.0 || k > n) continue;
    v = std::log(v * alpha / (a / (us * us) + b));
    if (!have_h) {
      h = (m + 0.5) * std::log((m + 1.0) / (r * (n - m + 1.0))) + stirling_tail(m) + stirling_tail(n - m);
      have_h = true;
    }
    const double bound = h + (n + 1.0) * std::log((n - m + 1.0) / (n - k + 1.0)) +
                         (k + 0.5) * std::log(r * (n - k + 1.0) / (k + 1.0)) - stirling_tail(k) - stirling_tail(n - k);
    if (v <= bound) return static_cast<uint64_t>(k);
  }
}

}  // namespace detail

// one draw of Binomial(n, p); n < 2^53
inline uint64_t binomial(PlanRng& rng, uint64_t n, double p) {
  if (n == 0 || !(p > 0.0)) return 0;
  if (p >= 1.0) return n;
  const bool flip = p > 0.5;
  const double pp = flip ? 1.0 - p : p;
  const uint64_t k = static_cast<double>(n) * pp < 10.0 ? detail::binomial_inversion(rng, n, pp) : detail::binomial_btrs(rng, n, pp);
  return flip ? n - k : k;
}

}  // namespace pcs
