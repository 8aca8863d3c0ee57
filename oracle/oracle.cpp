/*
 * oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call this file.  The product
 * (process_b200/, libpcs_seq.so) never links or imports it.
 *
 * PARITY UNPINNED.  The arithmetic of the simulate_seq() hot path lives in
 * albertocasagrande/RACES @ 1142937 (ProCESS `configure:9-11`), which is not
 * vendored under /root/reference and cannot be fetched; the reference ships
 * no tests, golden vectors or fixtures for this path.
 * This file therefore restates the contract visible from ProCESS
 *   - sample list / normal sample / purity wiring  src/seq_simulation.cpp:566-578, 650-657
 *   - same seed for simulator and sequencer        src/seq_simulation.cpp:368-369, 552-561
 *   - sequencer variants                           src/seq_simulation.cpp:386-428
 *   - insert size law Binomial(t=mean/p, p=1-sd^2/mean)
 *                                                  src/seq_simulation.cpp:431-451
 *   - outputs: occurrences, coverage at the locus  src/seq_simulation.cpp:92-140
 *   - genome model chromosome -> alleles -> fragments -> SIDs keyed by position
 *                                                  src/phylogenetic_forest.cpp:279-376
 * and freezes the RACES-internal rules as DESIGN.md "Semantics" lists them
 * (SURVEY.md Appendix A9-A17).  It is pinned only against hand-computed micro
 * forests (tests/golden/).
 *
 * Shape: the loop nest of the reference as recollected -- chromosome -> sample
 * -> cell -> allele -> fragment -> reads -- over EXPLICIT per-cell genomes
 * (std::map based, like RACES' GenomeMutations), a per-base coverage vector per
 * (sample, chromosome), std::mt19937_64 and std:: distributions, one thread by
 * default.  None of the product's data structures (haplotype intervals, tiles,
 * Philox) appear here, so agreement is evidence and not tautology.
 */
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../include/pcs_seq.h"

namespace {

thread_local std::string g_err;

// A11, read placement inside a fragment (UNPINNED: RACES-internal).  0 = the rule the product implements: the start
// is uniform over the fragment and a template that runs past the fragment's end is dropped (it falls off the
// molecule).  1 = SURVEY.md Appendix A as first written: the start is uniform over the starts from which the
// template fits, nothing is dropped.  Selectable so that the deviation can be measured (tests/).
std::atomic<int> g_placement_rule{0};

// where the time of the last oracle_simulate went, in CPU-seconds summed over its worker threads:
// [0] building explicit genomes, [1] per (sample, chromosome) fixed work (zeroing the per-base coverage vector,
// drawing the per-fragment template counts, gathering the tables), [2] the read loop, [3] wall-clock of the call
std::atomic<double> g_timing[4];
void add_time(int k, double sec) {
  double old = g_timing[k].load();
  while (!g_timing[k].compare_exchange_weak(old, old + sec)) {}
}
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- genome model
struct Fragment {
  uint32_t begin, end;                    // inclusive, 1-based
  std::map<uint32_t, uint32_t> sids;      // position -> mutation row (somatic only)
};

struct Allele {
  uint16_t origin;                        // germline allele this one descends from
  std::map<uint32_t, Fragment> fragments; // keyed by begin
};

struct ChrGenome {                        // one chromosome of one cell
  std::map<uint16_t, Allele> alleles;
  uint16_t next_id = 0;
};

struct Forest {
  const pcs_forest_desc* d;
  std::vector<std::vector<uint32_t>> children;
  std::vector<uint32_t> roots;
  std::vector<int64_t> node_leaf;         // node -> leaf index or -1
  // germline SIDs per chromosome and germline allele: sorted (pos,row)
  std::vector<std::array<std::vector<std::pair<uint32_t, uint32_t>>, 2>> germ;
  // mutation rows of a chromosome: [first,last)
  std::vector<uint32_t> chr_row_off;
};

void check(bool ok, const char* msg) {
  if (!ok) throw std::domain_error(msg);
}

Forest build_forest(const pcs_forest_desc* d) {
  Forest f;
  f.d = d;
  check(d->n_chr > 0 && d->n_chr < 65535, "n_chr out of range");
  f.children.resize(d->n_nodes);
  for (uint32_t v = 0; v < d->n_nodes; ++v) {
    int32_t p = d->node_parent[v];
    if (p < 0) {
      f.roots.push_back(v);
    } else {
      check(static_cast<uint32_t>(p) < v, "node_parent must precede the child");
      f.children[p].push_back(v);
    }
  }
  f.node_leaf.assign(d->n_nodes, -1);
  for (uint32_t l = 0; l < d->n_leaves; ++l) {
    check(d->leaf_node[l] < d->n_nodes, "leaf_node out of range");
    check(f.children[d->leaf_node[l]].empty(), "a sampled cell must be a leaf");
    check(d->leaf_sample[l] < d->n_samples, "leaf_sample out of range");
    f.node_leaf[d->leaf_node[l]] = l;
  }
  f.chr_row_off.assign(d->n_chr + 1, 0);
  for (uint32_t m = 0; m < d->n_mut; ++m) {
    check(d->mut_chr[m] < d->n_chr, "mut_chr out of range");
    if (m > 0) {
      check(d->mut_chr[m - 1] < d->mut_chr[m] ||
                (d->mut_chr[m - 1] == d->mut_chr[m] && d->mut_pos[m - 1] <= d->mut_pos[m]),
            "mutation table must be sorted by (chr, pos)");
    }
    check(d->mut_pos[m] >= 1 && d->mut_pos[m] <= d->chr_len[d->mut_chr[m]],
          "mutation position outside the chromosome");
    check(d->mut_ref_len[m] >= 1 && d->mut_alt_len[m] >= 1, "ref/alt must be non-empty");
    f.chr_row_off[d->mut_chr[m] + 1] = m + 1;
  }
  for (uint32_t c = 1; c <= d->n_chr; ++c)
    f.chr_row_off[c] = std::max(f.chr_row_off[c], f.chr_row_off[c - 1]);
  f.germ.resize(d->n_chr);
  for (uint64_t i = 0; i < d->n_germline; ++i) {
    uint32_t m = d->germ_mut[i];
    check(m < d->n_mut, "germ_mut out of range");
    uint32_t c = d->mut_chr[m];
    uint8_t mask = d->germ_allele_mask[i];
    check(mask != 0 && (mask >> d->chr_n_alleles[c]) == 0, "germ_allele_mask names a missing allele");
    for (int a = 0; a < 2; ++a)
      if (mask & (1u << a)) f.germ[c][a].push_back({d->mut_pos[m], m});
  }
  for (auto& g : f.germ)
    for (auto& v : g) {
      std::sort(v.begin(), v.end());
      for (size_t i = 1; i < v.size(); ++i)
        check(v[i].first != v[i - 1].first, "two germline SIDs at one position of one allele");
    }
  return f;
}

ChrGenome germline_genome(const Forest& f, uint32_t chr) {
  ChrGenome g;
  uint8_t n = f.d->chr_n_alleles[chr];
  check(n >= 1 && n <= 2, "chr_n_alleles must be 1 or 2");
  for (uint16_t a = 0; a < n; ++a) {
    Allele al;
    al.origin = a;
    Fragment fr;
    fr.begin = 1;
    fr.end = f.d->chr_len[chr];
    al.fragments.emplace(1u, std::move(fr));
    g.alleles.emplace(a, std::move(al));
  }
  g.next_id = n;
  return g;
}

// fragments of `src` clipped to [lo,hi]
std::map<uint32_t, Fragment> clip(const std::map<uint32_t, Fragment>& src, uint32_t lo, uint32_t hi) {
  std::map<uint32_t, Fragment> out;
  for (const auto& [b, fr] : src) {
    if (fr.end < lo || fr.begin > hi) continue;
    Fragment n;
    n.begin = std::max(fr.begin, lo);
    n.end = std::min(fr.end, hi);
    for (auto it = fr.sids.lower_bound(n.begin); it != fr.sids.end() && it->first <= n.end; ++it)
      n.sids.insert(*it);
    out.emplace(n.begin, std::move(n));
  }
  return out;
}

void apply_event(const Forest& f, ChrGenome& g, uint32_t chr, uint64_t e, bool preneo_only,
                 bool* stop) {
  const pcs_forest_desc* d = f.d;
  uint8_t kind = d->ev_kind[e];
  if (preneo_only && !(kind == PCS_EV_SID && d->ev_nature[e] == PCS_NATURE_PRENEOPLASTIC)) {
    *stop = true;  // pre-neoplastic SIDs are a prefix of the root's events
    return;
  }
  if (kind == PCS_EV_WGD) {
    std::vector<uint16_t> ids;
    for (const auto& [id, al] : g.alleles) ids.push_back(id);
    for (uint16_t id : ids) {
      Allele copy = g.alleles.at(id);
      g.alleles.emplace(g.next_id++, std::move(copy));
    }
    return;
  }
  if (d->ev_chr[e] != chr) return;
  uint16_t a = d->ev_allele[e];
  auto it = g.alleles.find(a);
  switch (kind) {
    case PCS_EV_SID: {
      if (it == g.alleles.end()) return;
      uint32_t m = d->ev_mut[e];
      check(m < d->n_mut && d->mut_chr[m] == chr, "SID event names a row of another chromosome");
      uint32_t pos = d->mut_pos[m];
      for (auto& [b, fr] : it->second.fragments) {
        if (pos >= fr.begin && pos <= fr.end) {
          check(fr.sids.emplace(pos, m).second, "two SIDs at one position of one allele");
          const auto& gv = f.germ[chr][it->second.origin];
          auto gi = std::lower_bound(gv.begin(), gv.end(), std::make_pair(pos, 0u));
          check(gi == gv.end() || gi->first != pos,
                "somatic SID at a germline SID position of the same allele");
          break;
        }
      }
      return;
    }
    case PCS_EV_CNA_AMP: {
      if (it == g.alleles.end()) return;
      uint32_t lo = d->ev_pos[e], hi = d->ev_pos[e] + d->ev_len[e] - 1;
      uint16_t dest = d->ev_dest[e];
      check(g.alleles.find(dest) == g.alleles.end(), "amplification destination allele exists");
      Allele n;
      n.origin = it->second.origin;
      n.fragments = clip(it->second.fragments, lo, hi);
      g.alleles.emplace(dest, std::move(n));
      g.next_id = std::max<uint16_t>(g.next_id, dest + 1);
      return;
    }
    case PCS_EV_CNA_DEL: {
      if (it == g.alleles.end()) return;
      uint32_t lo = d->ev_pos[e], hi = d->ev_pos[e] + d->ev_len[e] - 1;
      std::map<uint32_t, Fragment> left, right;
      if (lo > 1) left = clip(it->second.fragments, 1, lo - 1);
      if (hi < d->chr_len[chr]) right = clip(it->second.fragments, hi + 1, d->chr_len[chr]);
      left.merge(right);
      it->second.fragments = std::move(left);
      return;
    }
    default:
      throw std::domain_error("unknown event kind");
  }
}

// explicit genomes (one chromosome) of every leaf, by replaying events root -> leaf
struct ChrGenomes {
  std::vector<ChrGenome> leaf;          // [n_leaves]
  ChrGenome normal_plain;
  std::vector<ChrGenome> normal_preneo; // one per root
};

ChrGenomes build_chr_genomes(const Forest& f, uint32_t chr) {
  const pcs_forest_desc* d = f.d;
  ChrGenomes out;
  out.leaf.resize(d->n_leaves);
  out.normal_plain = germline_genome(f, chr);
  struct Item {
    uint32_t node;
    ChrGenome g;
  };
  for (uint32_t r : f.roots) {
    ChrGenome pn = germline_genome(f, chr);
    bool stop = false;
    for (uint64_t e = d->node_event_off[r]; e < d->node_event_off[r + 1] && !stop; ++e)
      apply_event(f, pn, chr, e, true, &stop);
    out.normal_preneo.push_back(std::move(pn));

    std::vector<Item> stack;
    stack.push_back({r, germline_genome(f, chr)});
    while (!stack.empty()) {
      Item it = std::move(stack.back());
      stack.pop_back();
      bool dummy = false;
      for (uint64_t e = d->node_event_off[it.node]; e < d->node_event_off[it.node + 1]; ++e)
        apply_event(f, it.g, chr, e, false, &dummy);
      if (f.node_leaf[it.node] >= 0) out.leaf[f.node_leaf[it.node]] = it.g;
      const auto& ch = f.children[it.node];
      for (size_t i = 0; i < ch.size(); ++i) {
        if (i + 1 == ch.size())
          stack.push_back({ch[i], std::move(it.g)});
        else
          stack.push_back({ch[i], it.g});
      }
    }
  }
  return out;
}

// ------------------------------------------------------------------ read walk
struct ErrMask {
  std::vector<uint64_t> w;
  void reset(uint32_t R) { w.assign((R + 63) / 64, 0); }
  void set(uint32_t i) { w[i >> 6] |= (1ull << (i & 63)); }
  bool any(uint32_t from, uint32_t n) const {
    for (uint32_t i = from; i < from + n; ++i)
      if ((i >> 6) < w.size() && (w[i >> 6] >> (i & 63)) & 1) return true;
    return false;
  }
};

struct SampleChrCounts {
  std::vector<uint32_t> cov;                // per base, index = position
  std::vector<uint32_t>* occ;               // [n_mut] of this sample (shared across chr)
};

// One read of R bases taken from `al`/`fr` starting at reference position x.
// Carried SIDs = somatic SIDs of the fragment U germline SIDs of the allele's
// origin lying inside the fragment.  A carried SID covers its own position
// only; the reference bases it replaces after the first are not covered.
void walk_read(const Forest& f, uint32_t chr, const Allele& al, const Fragment& fr, uint32_t x,
               uint32_t R, const ErrMask* err, SampleChrCounts& out) {
  const pcs_forest_desc* d = f.d;
  const auto& gv = f.germ[chr][al.origin];
  auto gi = std::lower_bound(gv.begin(), gv.end(), std::make_pair(x, 0u));
  auto si = fr.sids.lower_bound(x);
  uint32_t q = x, rem = R;
  while (rem > 0 && q <= fr.end) {
    // next carried SID at or after q
    while (gi != gv.end() && gi->first < q) ++gi;
    while (si != fr.sids.end() && si->first < q) ++si;
    bool has = false;
    uint32_t p = 0, m = 0;
    if (gi != gv.end() && gi->first <= fr.end) {
      has = true;
      p = gi->first;
      m = gi->second;
    }
    if (si != fr.sids.end() && si->first <= fr.end && (!has || si->first < p)) {
      has = true;
      p = si->first;
      m = si->second;
    }
    if (!has || p - q >= rem) {
      uint32_t last = std::min<uint64_t>(static_cast<uint64_t>(q) + rem - 1, fr.end);
      for (uint32_t b = q; b <= last; ++b) ++out.cov[b];
      return;
    }
    for (uint32_t b = q; b <= p; ++b) ++out.cov[b];
    rem -= (p - q);
    uint32_t offset = R - rem;
    uint32_t rl = d->mut_ref_len[m], alen = d->mut_alt_len[m];
    uint32_t consumed = std::min(alen, rem);
    if (!(err && err->any(offset, consumed))) ++(*out.occ)[m];
    rem -= consumed;
    q = p + rl;
  }
}

// --------------------------------------------------------------- error models
// error probability of read base i (DESIGN.md "Sequencer models")
inline double ramp(uint32_t i, uint32_t R) { return R > 1 ? 0.5 + static_cast<double>(i) / (R - 1) : 1.0; }
constexpr double kQualSigma = 0.5;

template <class RNG>
void draw_errors(RNG& rng, uint32_t sequencer, double rate, uint32_t R, ErrMask& mask) {
  mask.reset(R);
  std::uniform_real_distribution<double> unif(0.0, 1.0);
  if (sequencer == PCS_SEQ_BASIC_CONSTANT) {
    for (uint32_t i = 0; i < R; ++i)
      if (unif(rng) < rate) mask.set(i);
  } else if (sequencer == PCS_SEQ_BASIC_RANDOM) {
    std::normal_distribution<double> norm(0.0, 1.0);
    for (uint32_t i = 0; i < R; ++i) {
      double e = rate * ramp(i, R) * std::exp(kQualSigma * norm(rng) - 0.5 * kQualSigma * kQualSigma);
      if (unif(rng) < std::min(1.0, e)) mask.set(i);
    }
  }
}

// ------------------------------------------------------------------ sample set
struct CellRef {
  const ChrGenome* g;
  double weight;
  uint32_t cell;   // leaf index / root ordinal / 0
  uint16_t flags;  // PCS_PLACE_*
};

struct OutSample {
  std::vector<uint32_t> tumour_leaves;
  bool is_normal = false;
};

std::vector<OutSample> out_samples(const Forest& f, const pcs_seq_params& P, const uint32_t* leaf_group,
                                   uint32_t n_groups) {
  std::vector<OutSample> s;
  if (!P.normal_only) {
    s.resize(n_groups);
    for (uint32_t l = 0; l < f.d->n_leaves; ++l) {
      uint32_t g = leaf_group ? leaf_group[l] : f.d->leaf_sample[l];
      check(g < n_groups, "leaf group out of range");
      s[g].tumour_leaves.push_back(l);
    }
  }
  if (P.normal_only || P.with_normal_sample) {
    OutSample n;
    n.is_normal = true;
    s.push_back(n);
  }
  return s;
}

std::vector<CellRef> sample_cells(const ChrGenomes& G, const OutSample& s, const pcs_seq_params& P,
                                  bool preneo) {
  std::vector<CellRef> cells;
  double purity = s.is_normal ? 0.0 : P.purity;
  if (!s.is_normal && s.tumour_leaves.empty()) purity = 0.0;
  if (purity > 0)
    for (uint32_t l : s.tumour_leaves)
      cells.push_back({&G.leaf[l], purity / s.tumour_leaves.size(), l, PCS_PLACE_TUMOUR});
  if (purity < 1) {
    if (preneo) {
      for (uint32_t r = 0; r < G.normal_preneo.size(); ++r)
        cells.push_back({&G.normal_preneo[r], (1 - purity) / G.normal_preneo.size(), r,
                         PCS_PLACE_NORMAL_PRENEO});
    } else {
      cells.push_back({&G.normal_plain, 1 - purity, 0, PCS_PLACE_NORMAL_PLAIN});
    }
  }
  return cells;
}

struct Trace {
  pcs_read_placement* rec;
  uint32_t* masks;
  uint64_t cap;
  uint64_t n = 0;
};

struct FragRef {
  const Allele* al;
  const Fragment* fr;
  uint32_t cell;
  uint16_t allele, flags;
  double weight;
};

void simulate_sample_chr(const Forest& f, const ChrGenomes& G, uint32_t chr, uint32_t s_idx,
                         const OutSample& s, const pcs_seq_params& P, uint32_t n_threads,
                         std::vector<uint32_t>& occ_row, std::vector<uint32_t>& cov_row, Trace* trace,
                         uint64_t* n_reads_out) {
  const pcs_forest_desc* d = f.d;
  const double t_begin = now_s();
  std::atomic<double> t_loop{0.0};
  const uint32_t R = P.read_size;
  const bool paired = P.insert_size_mean > 0;
  const uint32_t mates = paired ? 2 : 1;
  const uint32_t clen = d->chr_len[chr];
  bool preneo = P.normal_only ? P.preneoplastic_in_normal : P.preneoplastic_in_normal;
  std::vector<CellRef> cells = sample_cells(G, s, P, preneo);

  std::vector<FragRef> frags;
  double total_w = 0;
  for (const auto& c : cells)
    for (const auto& [aid, al] : c.g->alleles)
      for (const auto& [b, fr] : al.fragments) {
        double w = c.weight * (static_cast<double>(fr.end) - fr.begin + 1);
        frags.push_back({&al, &fr, c.cell, aid, c.flags, w});
        total_w += w;
      }

  uint64_t N = static_cast<uint64_t>(std::llround(P.coverage * clen / (static_cast<double>(R) * mates)));
  std::seed_seq sq{static_cast<uint32_t>(P.seed), chr, s_idx, 0xC0FFEEu};
  std::mt19937_64 rng(sq);

  // templates per fragment: multinomial as sequential binomials
  std::vector<uint64_t> n_frag(frags.size(), 0);
  {
    uint64_t left = N;
    double wleft = total_w;
    for (size_t i = 0; i < frags.size() && left > 0; ++i) {
      double p = (i + 1 == frags.size()) ? 1.0 : std::min(1.0, std::max(0.0, frags[i].weight / wleft));
      uint64_t k = (p >= 1.0) ? left : std::binomial_distribution<uint64_t>(left, p)(rng);
      n_frag[i] = k;
      left -= k;
      wleft -= frags[i].weight;
    }
  }

  std::binomial_distribution<uint32_t> insert_dist;
  if (paired) {
    // get_bin_dist(): src/seq_simulation.cpp:431-451
    double q = static_cast<double>(P.insert_size_stddev) * P.insert_size_stddev / P.insert_size_mean;
    double p = 1 - q;
    check(p >= 0, "insert size mean must be >= its variance");
    uint32_t t = static_cast<uint32_t>(P.insert_size_mean / p);
    insert_dist = std::binomial_distribution<uint32_t>(t, p);
  }

  if (trace) n_threads = 1;
  n_threads = std::max<uint32_t>(1, std::min<uint32_t>(n_threads, frags.size() ? frags.size() : 1));
  std::vector<std::vector<uint32_t>> cov(n_threads), occ(n_threads);
  std::vector<uint64_t> placed(n_threads, 0);

  auto work = [&](uint32_t t) {
    cov[t].assign(static_cast<size_t>(clen) + 2, 0);
    occ[t].assign(d->n_mut, 0);
    SampleChrCounts out{std::move(cov[t]), &occ[t]};
    std::seed_seq tsq{static_cast<uint32_t>(P.seed), chr, s_idx, t + 1u};
    std::mt19937_64 trng(tsq);
    auto ins = insert_dist;
    ErrMask m1, m2;
    size_t lo = frags.size() * t / n_threads, hi = frags.size() * (t + 1) / n_threads;
    const double t_loop_begin = now_s();
    for (size_t i = lo; i < hi; ++i) {
      const FragRef& fr = frags[i];
      std::uniform_int_distribution<uint32_t> start(fr.fr->begin, fr.fr->end);
      const int rule = g_placement_rule.load();
      for (uint64_t k = 0; k < n_frag[i]; ++k) {
        uint32_t x = start(trng);
        uint32_t gap = paired ? ins(trng) : 0;
        uint64_t tlen = paired ? 2ull * R + gap : R;
        if (rule == 1) {  // uniform over the valid starts; a fragment shorter than the template yields nothing
          if (static_cast<uint64_t>(fr.fr->end) - fr.fr->begin + 1 < tlen) continue;
          x = std::uniform_int_distribution<uint32_t>(fr.fr->begin, static_cast<uint32_t>(fr.fr->end - tlen + 1))(trng);
        }
        if (static_cast<uint64_t>(x) + tlen - 1 > fr.fr->end) continue;  // falls off the molecule
        for (uint32_t mate = 0; mate < mates; ++mate) {
          uint32_t xs = mate == 0 ? x : x + R + gap;
          ErrMask* em = nullptr;
          if (P.sequencer != PCS_SEQ_ERRORLESS) {
            em = mate == 0 ? &m1 : &m2;
            draw_errors(trng, P.sequencer, P.error_rate, R, *em);
          }
          walk_read(f, chr, *fr.al, *fr.fr, xs, R, em, out);
          ++placed[t];
          if (trace && trace->n < trace->cap) {
            pcs_read_placement& r = trace->rec[trace->n];
            r.cell = fr.cell;
            r.start = xs;
            r.chr = static_cast<uint16_t>(chr);
            r.allele = fr.allele;
            r.sample = static_cast<uint16_t>(s_idx);
            r.flags = fr.flags;
            if (trace->masks) {
              uint32_t* mw = trace->masks + trace->n * PCS_ERRMASK_WORDS;
              std::memset(mw, 0, sizeof(uint32_t) * PCS_ERRMASK_WORDS);
              if (em)
                for (uint32_t b = 0; b < R && b < 32 * PCS_ERRMASK_WORDS; ++b)
                  if (em->any(b, 1)) mw[b >> 5] |= (1u << (b & 31));
            }
          }
          if (trace) ++trace->n;
        }
      }
    }
    {
      const double dt = now_s() - t_loop_begin;
      double old = t_loop.load();
      while (!t_loop.compare_exchange_weak(old, old + dt)) {}
    }
    cov[t] = std::move(out.cov);
  };

  if (n_threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < n_threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  for (uint32_t t = 0; t < n_threads; ++t) {
    for (uint32_t m = f.chr_row_off[chr]; m < f.chr_row_off[chr + 1]; ++m) {
      occ_row[m] += occ[t][m];
      cov_row[m] += cov[t][d->mut_pos[m]];
    }
    *n_reads_out += placed[t];
  }
  add_time(2, t_loop.load());
  // everything else of this task, in CPU-seconds: with inner threads the zeroing runs on all of them
  add_time(1, std::max(0.0, (now_s() - t_begin) * n_threads - t_loop.load()));
}

bool chr_selected(const pcs_seq_params& P, uint32_t c) { return !P.chr_mask || P.chr_mask[c]; }

void validate_params(const pcs_seq_params& P) {
  check(P.read_size >= 1, "read_size must be positive");
  check(P.coverage >= 0, "coverage must be non-negative");
  check(P.normal_only || (P.purity >= 0 && P.purity <= 1), "purity must belong to [0,1]");
  check(P.sequencer <= PCS_SEQ_BASIC_RANDOM, "Unsupported sequencer type");
  check(P.error_rate >= 0, "error_rate must be non-negative");
}

}  // namespace

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

/* A11 placement rule of the following oracle_simulate calls: 0 (default) uniform over the fragment + drop,
 * 1 uniform over the valid starts.  Returns the previous rule. */
int oracle_set_placement_rule(int rule) { return g_placement_rule.exchange(rule == 1 ? 1 : 0); }

/* CPU-seconds of the last oracle_simulate of the process: out[0] explicit genomes, out[1] fixed work per
 * (sample, chromosome), out[2] the read loop, out[3] wall-clock seconds of the call */
void oracle_last_timing(double out[4]) {
  for (int k = 0; k < 4; ++k) out[k] = g_timing[k].load();
}

/* free-running simulation.  occ/cov: [n_out_samples][n_mut].  trace_* may be NULL. */
int oracle_simulate(const pcs_forest_desc* desc, const pcs_seq_params* params, const uint32_t* leaf_group,
                    uint32_t n_groups, uint32_t n_threads, uint32_t* occ, uint32_t* cov,
                    pcs_read_placement* trace_rec, uint32_t* trace_masks, uint64_t trace_cap,
                    uint64_t* trace_n, uint64_t* n_reads) {
  try {
    Forest f = build_forest(desc);
    validate_params(*params);
    if (!leaf_group) n_groups = desc->n_samples;
    auto samples = out_samples(f, *params, leaf_group, n_groups);
    std::vector<std::vector<uint32_t>> occ_v(samples.size(), std::vector<uint32_t>(desc->n_mut, 0));
    auto cov_v = occ_v;
    Trace tr{trace_rec, trace_masks, trace_cap, 0};
    uint64_t placed = 0;
    for (auto& t : g_timing) t.store(0.0);
    const double t_call = now_s();
    std::vector<uint32_t> chrs;
    for (uint32_t c = 0; c < desc->n_chr; ++c)
      if (chr_selected(*params, c)) chrs.push_back(c);
    if (n_threads > 1 && chrs.size() >= 2 && !trace_rec) {
      // several chromosomes: one worker per chromosome (longest first), each single-threaded inside -- the
      // tables are those of a one-thread run (the RNG streams are keyed by (seed, chromosome, sample, thread 0)),
      // and no thread zeroes a per-base coverage vector it does not fill
      std::sort(chrs.begin(), chrs.end(), [&](uint32_t a, uint32_t b) {
        return desc->chr_len[a] != desc->chr_len[b] ? desc->chr_len[a] > desc->chr_len[b] : a < b;
      });
      std::atomic<size_t> next{0};
      std::atomic<uint64_t> placed_all{0};
      std::vector<std::string> errors(chrs.size());
      auto worker = [&] {
        for (size_t k = next.fetch_add(1); k < chrs.size(); k = next.fetch_add(1)) {
          try {
            const uint32_t c = chrs[k];
            const double t0 = now_s();
            ChrGenomes G = build_chr_genomes(f, c);
            add_time(0, now_s() - t0);
            uint64_t mine = 0;
            for (uint32_t s = 0; s < samples.size(); ++s)  // rows of chromosome c only: no two workers share a cell
              simulate_sample_chr(f, G, c, s, samples[s], *params, 1, occ_v[s], cov_v[s], nullptr, &mine);
            placed_all += mine;
          } catch (const std::exception& e) {
            errors[k] = e.what();
          }
        }
      };
      std::vector<std::thread> th;
      for (uint32_t w = 1; w < std::min<size_t>(n_threads, chrs.size()); ++w) th.emplace_back(worker);
      worker();
      for (auto& t : th) t.join();
      for (const auto& e : errors)
        if (!e.empty()) throw std::domain_error(e);
      placed = placed_all.load();
    } else {
      for (uint32_t c : chrs) {
        const double t0 = now_s();
        ChrGenomes G = build_chr_genomes(f, c);
        add_time(0, now_s() - t0);
        for (uint32_t s = 0; s < samples.size(); ++s)
          simulate_sample_chr(f, G, c, s, samples[s], *params, n_threads, occ_v[s], cov_v[s],
                              trace_rec ? &tr : nullptr, &placed);
      }
    }
    g_timing[3].store(now_s() - t_call);
    for (uint32_t s = 0; s < samples.size(); ++s) {
      std::memcpy(occ + static_cast<size_t>(s) * desc->n_mut, occ_v[s].data(), sizeof(uint32_t) * desc->n_mut);
      std::memcpy(cov + static_cast<size_t>(s) * desc->n_mut, cov_v[s].data(), sizeof(uint32_t) * desc->n_mut);
    }
    if (trace_n) *trace_n = tr.n;
    if (n_reads) *n_reads = placed;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

/* count a list of read placements.  Every placement must start inside a
 * fragment of the named allele of the named cell. */
int oracle_count_injected(const pcs_forest_desc* desc, uint32_t n_out_samples, uint32_t read_size,
                          const pcs_read_placement* rec, const uint32_t* err_masks, uint64_t n,
                          uint32_t* occ, uint32_t* cov) {
  try {
    Forest f = build_forest(desc);
    std::vector<std::vector<uint32_t>> occ_v(n_out_samples, std::vector<uint32_t>(desc->n_mut, 0));
    std::vector<std::vector<uint64_t>> by_chr(desc->n_chr);
    for (uint64_t i = 0; i < n; ++i) {
      check(rec[i].chr < desc->n_chr, "placement chromosome out of range");
      check(rec[i].sample < n_out_samples, "placement sample out of range");
      by_chr[rec[i].chr].push_back(i);
    }
    std::memset(cov, 0, sizeof(uint32_t) * static_cast<size_t>(n_out_samples) * desc->n_mut);
    for (uint32_t c = 0; c < desc->n_chr; ++c) {
      if (by_chr[c].empty()) continue;
      ChrGenomes G = build_chr_genomes(f, c);
      std::vector<SampleChrCounts> sc;
      for (uint32_t s = 0; s < n_out_samples; ++s)
        sc.push_back({std::vector<uint32_t>(static_cast<size_t>(desc->chr_len[c]) + 2, 0), &occ_v[s]});
      ErrMask em;
      for (uint64_t i : by_chr[c]) {
        const pcs_read_placement& r = rec[i];
        const ChrGenome* g = nullptr;
        if (r.flags == PCS_PLACE_TUMOUR) {
          check(r.cell < desc->n_leaves, "placement cell out of range");
          g = &G.leaf[r.cell];
        } else if (r.flags == PCS_PLACE_NORMAL_PLAIN) {
          g = &G.normal_plain;
        } else if (r.flags == PCS_PLACE_NORMAL_PRENEO) {
          check(r.cell < G.normal_preneo.size(), "placement root out of range");
          g = &G.normal_preneo[r.cell];
        } else {
          throw std::domain_error("unknown placement flags");
        }
        auto ai = g->alleles.find(r.allele);
        check(ai != g->alleles.end(), "placement names a missing allele");
        const Fragment* fr = nullptr;
        for (const auto& [b, x] : ai->second.fragments)
          if (r.start >= x.begin && r.start <= x.end) fr = &x;
        check(fr != nullptr, "placement starts outside every fragment of the allele");
        const ErrMask* emp = nullptr;
        if (err_masks) {
          em.reset(32 * PCS_ERRMASK_WORDS);
          const uint32_t* mw = err_masks + i * PCS_ERRMASK_WORDS;
          for (uint32_t b = 0; b < 32 * PCS_ERRMASK_WORDS; ++b)
            if ((mw[b >> 5] >> (b & 31)) & 1) em.set(b);
          emp = &em;
        }
        walk_read(f, c, ai->second, *fr, r.start, read_size, emp, sc[r.sample]);
      }
      for (uint32_t s = 0; s < n_out_samples; ++s)
        for (uint32_t m = f.chr_row_off[c]; m < f.chr_row_off[c + 1]; ++m)
          cov[static_cast<size_t>(s) * desc->n_mut + m] = sc[s].cov[desc->mut_pos[m]];
    }
    for (uint32_t s = 0; s < n_out_samples; ++s)
      std::memcpy(occ + static_cast<size_t>(s) * desc->n_mut, occ_v[s].data(), sizeof(uint32_t) * desc->n_mut);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

/* Materialise reads (SAM content) from a placement list: bases of the reference with the carried SIDs of
 * the named allele applied, CIGAR (length << 4 | op; 0 M, 1 I, 2 D; at most 16 runs), qualities for the
 * errorless ('I') and constant-quality (error '#', else Phred of error_rate) models; bases whose bit is set
 * in the read's error mask are substituted by (code + 1 + offset % 3) & 3 over ACGT.
 * ref_off[c] = offset of position 1 of chromosome c in ref_bases. */
int oracle_materialize(const pcs_forest_desc* desc, const uint64_t* ref_off, const char* ref_bases,
                       const uint32_t* alt_off, const char* alt_bytes, uint32_t read_size, uint32_t sequencer,
                       double error_rate, const pcs_read_placement* rec, const uint32_t* err_masks, uint64_t n,
                       uint8_t* seq, uint8_t* qual, uint32_t* cigar, uint32_t* n_cigar, uint32_t* lengths) {
  try {
    Forest f = build_forest(desc);
    const uint32_t R = read_size;
    std::vector<std::vector<uint64_t>> by_chr(desc->n_chr);
    for (uint64_t i = 0; i < n; ++i) {
      check(rec[i].chr < desc->n_chr, "placement chromosome out of range");
      by_chr[rec[i].chr].push_back(i);
    }
    auto subst = [](uint8_t b, uint32_t o) -> uint8_t {
      const char* acgt = "ACGT";
      const char* at = std::strchr(acgt, b);
      if (!at || !b) return b;
      return static_cast<uint8_t>(acgt[((at - acgt) + 1 + o % 3) & 3]);
    };
    int qc = static_cast<int>(-10.0 * std::log10(std::max(std::min(error_rate, 1.0), 1e-5)) + 0.5);
    qc = std::min(41, std::max(2, qc));
    for (uint32_t c = 0; c < desc->n_chr; ++c) {
      if (by_chr[c].empty()) continue;
      ChrGenomes G = build_chr_genomes(f, c);
      const char* ref = ref_bases + ref_off[c] - 1;  // ref[p]: base at 1-based position p
      for (uint64_t i : by_chr[c]) {
        const pcs_read_placement& r = rec[i];
        const ChrGenome* g = r.flags == PCS_PLACE_TUMOUR ? &G.leaf.at(r.cell)
                             : r.flags == PCS_PLACE_NORMAL_PLAIN ? &G.normal_plain : &G.normal_preneo.at(r.cell);
        auto ai = g->alleles.find(r.allele);
        check(ai != g->alleles.end(), "placement names a missing allele");
        const Fragment* fr = nullptr;
        for (const auto& [b, x] : ai->second.fragments)
          if (r.start >= x.begin && r.start <= x.end) fr = &x;
        check(fr != nullptr, "placement starts outside every fragment of the allele");
        const uint32_t* mw = err_masks ? err_masks + i * PCS_ERRMASK_WORDS : nullptr;
        uint8_t* sq = seq + i * R;
        uint8_t* ql = qual + i * R;
        uint32_t* cg = cigar + i * 16;
        uint32_t nc = 0, o = 0;
        auto push = [&](uint32_t op, uint32_t len) {
          if (!len) return;
          if (nc && (cg[nc - 1] & 15u) == op) cg[nc - 1] += len << 4;
          else if (nc < 16) cg[nc++] = (len << 4) | op;
        };
        auto put = [&](uint8_t b) {
          bool err = mw && o < 32 * PCS_ERRMASK_WORDS && ((mw[o >> 5] >> (o & 31)) & 1);
          sq[o] = err ? subst(b, o) : b;
          ql[o] = sequencer == PCS_SEQ_ERRORLESS ? 'I' : sequencer == PCS_SEQ_BASIC_CONSTANT ? (err ? '#' : static_cast<uint8_t>(33 + qc)) : 0;
          ++o;
        };
        const auto& gv = f.germ[c][ai->second.origin];
        auto gi = std::lower_bound(gv.begin(), gv.end(), std::make_pair(r.start, 0u));
        auto si = fr->sids.lower_bound(r.start);
        uint32_t q = r.start;
        while (o < R && q <= fr->end) {
          while (gi != gv.end() && gi->first < q) ++gi;
          while (si != fr->sids.end() && si->first < q) ++si;
          bool has = false;
          uint32_t p = 0, m = 0;
          if (gi != gv.end() && gi->first <= fr->end) { has = true; p = gi->first; m = gi->second; }
          if (si != fr->sids.end() && si->first <= fr->end && (!has || si->first < p)) { has = true; p = si->first; m = si->second; }
          if (!has || p - q >= R - o) {
            uint32_t last = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(q) + (R - o) - 1, fr->end));
            push(0, last - q + 1);
            for (uint32_t b = q; b <= last; ++b) put(static_cast<uint8_t>(ref[b]));
            break;
          }
          push(0, p - q);
          for (uint32_t b = q; b < p; ++b) put(static_cast<uint8_t>(ref[b]));
          const uint32_t rl = desc->mut_ref_len[m], al = desc->mut_alt_len[m];
          const uint32_t consumed = std::min(al, R - o);
          const char* alt = alt_bytes + alt_off[m];
          for (uint32_t b = 0; b < consumed; ++b) put(static_cast<uint8_t>(alt[b]));
          const uint32_t mm = std::min(consumed, rl);
          push(0, mm);
          push(1, consumed - mm);
          q = p + rl;
          if (o < R && rl > al) push(2, rl - al);
        }
        n_cigar[i] = nc;
        lengths[i] = o;
      }
    }
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

/* Every explicit genome of one chromosome at once, as CSR arrays: the sampled cells 0 .. n_leaves-1, then (if
 * with_preneo) the normal cells carrying the pre-neoplastic SIDs, one per root, as cells n_leaves + root ordinal.
 * Alleles without DNA are listed with no fragment.  Two calls: with every capacity 0 to size the arrays
 * (n_out = {alleles, fragments, SIDs}), then to fill them. */
int oracle_chr_genomes(const pcs_forest_desc* desc, uint32_t chr, int with_preneo, uint64_t cap_alleles,
                       uint64_t cap_frags, uint64_t cap_sids, uint32_t* allele_cell, uint16_t* allele_id,
                       uint8_t* allele_origin, uint64_t* allele_frag_off, uint32_t* frag_begin, uint32_t* frag_end,
                       uint64_t* allele_sid_off, uint32_t* sid_row, uint64_t n_out[3]) {
  try {
    Forest f = build_forest(desc);
    check(chr < desc->n_chr, "chromosome out of range");
    ChrGenomes G = build_chr_genomes(f, chr);
    uint64_t na = 0, nf = 0, ns = 0;
    auto emit = [&](const ChrGenome& g, uint32_t cell) {
      for (const auto& [aid, al] : g.alleles) {
        if (na < cap_alleles) {
          allele_cell[na] = cell;
          allele_id[na] = aid;
          allele_origin[na] = static_cast<uint8_t>(al.origin);
          allele_frag_off[na] = nf;
          allele_sid_off[na] = ns;
        }
        ++na;
        for (const auto& [b, fr] : al.fragments) {
          if (nf < cap_frags) {
            frag_begin[nf] = fr.begin;
            frag_end[nf] = fr.end;
          }
          ++nf;
          for (const auto& [pos, row] : fr.sids) {
            if (ns < cap_sids) sid_row[ns] = row;
            ++ns;
          }
        }
      }
    };
    for (uint32_t l = 0; l < desc->n_leaves; ++l) emit(G.leaf[l], l);
    if (with_preneo)
      for (uint32_t r = 0; r < G.normal_preneo.size(); ++r) emit(G.normal_preneo[r], desc->n_leaves + r);
    if (na <= cap_alleles && cap_alleles) {
      allele_frag_off[na] = nf;
      allele_sid_off[na] = ns;
    }
    n_out[0] = na;
    n_out[1] = nf;
    n_out[2] = ns;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

/* explicit genome of one cell on one chromosome, for hand-checked tests.
 * which: PCS_PLACE_*; cell: leaf index / root ordinal.
 * Output (capacity `cap` each, *n_frag / *n_sid receive the counts):
 *   fragments: frag_allele[], frag_origin[], frag_begin[], frag_end[]
 *   somatic SIDs: sid_allele[], sid_row[]                                  */
int oracle_cell_genome(const pcs_forest_desc* desc, uint32_t which, uint32_t cell, uint32_t chr,
                       uint32_t cap, uint16_t* frag_allele, uint16_t* frag_origin, uint32_t* frag_begin,
                       uint32_t* frag_end, uint32_t* n_frag, uint16_t* sid_allele, uint32_t* sid_row,
                       uint32_t* n_sid) {
  try {
    Forest f = build_forest(desc);
    check(chr < desc->n_chr, "chromosome out of range");
    ChrGenomes G = build_chr_genomes(f, chr);
    const ChrGenome* g = nullptr;
    if (which == PCS_PLACE_TUMOUR) {
      check(cell < desc->n_leaves, "cell out of range");
      g = &G.leaf[cell];
    } else if (which == PCS_PLACE_NORMAL_PLAIN) {
      g = &G.normal_plain;
    } else {
      check(cell < G.normal_preneo.size(), "root out of range");
      g = &G.normal_preneo[cell];
    }
    uint32_t nf = 0, ns = 0;
    for (const auto& [aid, al] : g->alleles) {
      if (al.fragments.empty() && nf < cap) {
        // an allele that lost all its DNA is still listed, with an empty fragment
        frag_allele[nf] = aid; frag_origin[nf] = al.origin; frag_begin[nf] = 0; frag_end[nf] = 0;
      }
      if (al.fragments.empty()) ++nf;
      for (const auto& [b, fr] : al.fragments) {
        if (nf < cap) {
          frag_allele[nf] = aid; frag_origin[nf] = al.origin; frag_begin[nf] = fr.begin; frag_end[nf] = fr.end;
        }
        ++nf;
        for (const auto& [pos, row] : fr.sids) {
          if (ns < cap) { sid_allele[ns] = aid; sid_row[ns] = row; }
          ++ns;
        }
      }
    }
    *n_frag = nf;
    *n_sid = ns;
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
