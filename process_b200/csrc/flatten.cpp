// flatten.cpp -- forest description -> haplotype-interval view (see flat.hpp).
//
// Event semantics replayed here (they are the ones the CPU oracle replays on
// explicit genomes; DESIGN.md "Semantics"):
//   SID  on allele a : placed iff the lineage has allele a and a fragment of a holds the position
//   AMP  a -> d      : allele d = fragments of a clipped to [pos, pos+len-1] (with the SIDs there)
//   DEL  on allele a : [pos, pos+len-1] removed from the fragments of a
//   WGD              : per chromosome, every allele in increasing id order is copied to id next_id++
#include "flat.hpp"
#include "host_pool.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace pcs {
namespace {

using FragKey = std::vector<std::pair<uint32_t, uint32_t>>;

void check(bool ok, const char* msg) {
  if (!ok) throw std::domain_error(msg);
}

FragKey clip(const FragKey& s, uint32_t lo, uint32_t hi) {
  FragKey out;
  if (lo > hi) return out;
  for (const auto& f : s) {
    if (f.second < lo || f.first > hi) continue;
    out.emplace_back(std::max(f.first, lo), std::min(f.second, hi));
  }
  return out;
}

FragKey remove_range(const FragKey& s, uint32_t lo, uint32_t hi, uint32_t chr_len) {
  FragKey out = lo > 1 ? clip(s, 1, lo - 1) : FragKey{};
  if (hi < chr_len) {
    FragKey r = clip(s, hi + 1, chr_len);
    out.insert(out.end(), r.begin(), r.end());
  }
  return out;
}

bool holds(const FragKey& s, uint32_t pos) {
  for (const auto& f : s)
    if (pos >= f.first && pos <= f.second) return true;
  return false;
}

struct Tree {
  std::vector<uint32_t> child_off, child_idx, roots;
  std::vector<int64_t> node_leaf;
  // The same tree in DFS PREORDER (roots in order, children by increasing node index): position i holds node
  // pre_node[i], its subtree is the range [i, pre_end[i]), its first child is i + 1 and the sibling after a
  // child c is pre_end[c].  The haplotype numbering walks the tree once per allele copy: laid out like this
  // every walk reads memory front to back instead of chasing node indices.
  std::vector<uint32_t> pre_node, pre_end, root_pos;
  std::vector<int64_t> pre_leaf;
  std::vector<uint32_t> leaf_pos, leaf_id;  // the sampled cells by increasing preorder position, and their leaf index
};

// one event of a chromosome, gathered next to its neighbours in walk order
struct Ev {
  uint32_t x;      // SID: mutation row; CNA: first position; WGD: index of the event in the caller's arrays
  uint32_t y;      // SID: position of the row; CNA: length
  uint32_t meta;   // SID: ref_len | alt_len << 8
  uint16_t allele, dest;
  uint8_t kind, nature;
};

struct ChrWork {
  uint32_t chr = 0;
  std::vector<Ev> ev;                 // events of this chromosome (+ WGD), by preorder position
  std::vector<uint32_t> ev_off;       // [n_nodes+1], indexed by preorder position
  std::vector<uint32_t> ev_nodes;     // the preorder positions that have events of this chromosome, increasing
  bool has_wgd = false;
  uint32_t row_lo = 0, row_hi = 0;    // rows of this chromosome in the mutation table
  uint32_t clen = 0;                  // its length
  std::vector<Inst> inst;
  std::vector<HapRec> haps;           // fragset is a LOCAL id until merge
  std::vector<FragKey> fragsets;
  std::vector<uint8_t> fs_empty;      // fragment set has no DNA left
  std::map<FragKey, uint32_t> intern;
  std::unordered_map<uint64_t, std::vector<std::pair<uint16_t, uint16_t>>> wgd_map;
  uint32_t germ_lo[2] = {0, 0}, germ_hi[2] = {0, 0};  // haplotype interval below each germline allele
  std::vector<Piece> pieces;          // cover_off is LOCAL until merge
  std::vector<Cover> covers;          // fragset is LOCAL until merge

  uint32_t intern_set(const FragKey& k) {
    auto it = intern.find(k);
    if (it != intern.end()) return it->second;
    uint32_t id = static_cast<uint32_t>(fragsets.size());
    fragsets.push_back(k);
    fs_empty.push_back(k.empty() ? 1 : 0);
    intern.emplace(k, id);
    return id;
  }
};

// which allele ids exist where: only needed to give WGD copies their ids
void wgd_prepass(const pcs_forest_desc& d, const Tree& t, ChrWork& w) {
  struct AState {
    std::vector<uint16_t> ids;
    uint32_t next;
  };
  std::vector<AState> pool;
  AState base;
  for (uint16_t a = 0; a < d.chr_n_alleles[w.chr]; ++a) base.ids.push_back(a);
  base.next = d.chr_n_alleles[w.chr];
  pool.push_back(base);
  std::vector<std::pair<uint32_t, uint32_t>> stack;  // (end of the subtree, allele state below it)
  for (uint32_t i = 0; i < d.n_nodes; ++i) {
    while (!stack.empty() && stack.back().first <= i) stack.pop_back();
    uint32_t cur = stack.empty() ? 0u : stack.back().second;
    bool own = false;
    auto make_own = [&]() {
      if (!own) {
        pool.push_back(pool[cur]);
        cur = static_cast<uint32_t>(pool.size() - 1);
        own = true;
      }
    };
    for (uint32_t k = w.ev_off[i]; k < w.ev_off[i + 1]; ++k) {
      const Ev& e = w.ev[k];
      if (e.kind == PCS_EV_CNA_AMP) {
        const auto& ids = pool[cur].ids;
        if (!std::binary_search(ids.begin(), ids.end(), e.allele)) continue;
        check(!std::binary_search(ids.begin(), ids.end(), e.dest), "amplification destination allele exists");
        make_own();
        auto& mid = pool[cur].ids;
        mid.insert(std::upper_bound(mid.begin(), mid.end(), e.dest), e.dest);
        pool[cur].next = std::max<uint32_t>(pool[cur].next, e.dest + 1u);
      } else if (e.kind == PCS_EV_WGD) {
        make_own();
        std::vector<uint16_t> snapshot = pool[cur].ids;
        auto& m = w.wgd_map[e.x];
        for (uint16_t a : snapshot) {
          check(pool[cur].next < 65535, "allele id overflow");
          uint16_t nd = static_cast<uint16_t>(pool[cur].next++);
          m.emplace_back(a, nd);
          auto& mid = pool[cur].ids;
          mid.insert(std::upper_bound(mid.begin(), mid.end(), nd), nd);
        }
      }
    }
    if (t.pre_end[i] > i + 1) stack.emplace_back(t.pre_end[i], cur);
  }
}

// what every numbering ends with: placements without carriers dropped, the rest sorted by row; the pieces the
// fragment sets in use cut the chromosome into, each with the fragment sets that cover it
void finish_chr(ChrWork& w) {
  const uint32_t chr = w.chr, clen = w.clen;
  // a SID no sampled haplotype inherited has an empty interval: drop it
  w.inst.erase(std::remove_if(w.inst.begin(), w.inst.end(), [](const Inst& in) { return in.span == 0; }), w.inst.end());
  // by row, placements of one row in DFS order: stable LSD radix sort on (row - smallest row), 11 bits a pass
  if (w.inst.size() > 1) {
    uint32_t row_min = 0xffffffffu, row_max = 0;
    for (const Inst& in : w.inst) {
      row_min = std::min(row_min, in.row);
      row_max = std::max(row_max, in.row);
    }
    std::vector<Inst> tmp(w.inst.size());
    Inst* src = w.inst.data();
    Inst* dst = tmp.data();
    const size_t n_inst = w.inst.size();
    for (uint32_t shift = 0; shift < 32 && ((row_max - row_min) >> shift) != 0; shift += 11) {
      uint32_t cnt[2049] = {0};
      for (size_t i = 0; i < n_inst; ++i) ++cnt[(((src[i].row - row_min) >> shift) & 2047u) + 1];
      for (uint32_t b = 0; b < 2048; ++b) cnt[b + 1] += cnt[b];
      for (size_t i = 0; i < n_inst; ++i) dst[cnt[((src[i].row - row_min) >> shift) & 2047u]++] = src[i];
      std::swap(src, dst);
    }
    if (src != w.inst.data()) w.inst.swap(tmp);
  }

  // pieces: maximal intervals on which the set of covering fragments is constant
  std::vector<uint8_t> used(w.fragsets.size(), 0);
  for (const auto& h : w.haps) used[h.fragset] = 1;
  std::vector<uint32_t> bp{1u, clen + 1};
  for (size_t k = 0; k < w.fragsets.size(); ++k)
    if (used[k])
      for (const auto& p : w.fragsets[k]) {
        bp.push_back(p.first);
        bp.push_back(p.second + 1);
      }
  std::sort(bp.begin(), bp.end());
  bp.erase(std::unique(bp.begin(), bp.end()), bp.end());
  for (size_t i = 0; i + 1 < bp.size(); ++i) {
    Piece pc{chr, bp[i], bp[i + 1] - 1, static_cast<uint32_t>(w.covers.size()), 0};
    for (size_t k = 0; k < w.fragsets.size(); ++k) {
      if (!used[k]) continue;
      for (const auto& p : w.fragsets[k])
        if (p.first <= pc.begin && pc.end <= p.second) {
          w.covers.push_back({static_cast<uint32_t>(k), p.second});
          ++pc.cover_n;
          break;
        }
    }
    if (pc.cover_n) w.pieces.push_back(pc);
  }
}

// haplotype numbering of one chromosome: leaves (w.haps), the SOMATIC placements sorted by row (w.inst),
// the interval below each germline allele, the pieces its fragment sets cut the chromosome into
//
// A haplotype carries at most ONE SID per position (the kernels' walk and the read materialiser rely on it; the
// oracle refuses anything else).  Rows of one position are neighbours in the mutation table (one locus), so the
// check looks at the few rows of the new SID's locus: a germline SID of the germline allele the walk descends
// from (row_mask), or a somatic instance that is still open (row_open) -- a second SID of the haplotype at that
// position: std::domain_error, as the oracle.
void flatten_chr(const pcs_forest_desc& d, const Tree& t, ChrWork& w, const std::atomic<uint8_t>* row_mask,
                 const uint32_t* row_locus, const uint32_t* locus_first_row) {
  const uint32_t chr = w.chr;
  const uint32_t clen = d.chr_len[chr];
  const uint8_t n0 = d.chr_n_alleles[chr];
  check(n0 >= 1 && n0 <= 2, "chr_n_alleles must be 1 or 2");
  // what the walk needs of the rows the SID events name; here and not where the events were dealt out, because
  // the rows of one chromosome are a few MB of the mutation table, and those of all of them are not
  for (Ev& e : w.ev)
    if (e.kind == PCS_EV_SID) {
      const uint32_t m = e.x;
      check(m < d.n_mut && d.mut_chr[m] == chr, "SID event names a row of another chromosome");
      e.y = d.mut_pos[m];
      e.meta = static_cast<uint32_t>(d.mut_ref_len[m]) | (static_cast<uint32_t>(d.mut_alt_len[m]) << 8);
    }
  // allele ids along every lineage: WGD copies get their ids, and an amplification into an id the lineage
  // already has is refused -- with or without a WGD in the forest
  wgd_prepass(d, t, w);

  const uint32_t full = w.intern_set(FragKey{{1u, clen}});
  uint32_t counter = 0;
  uint32_t* germ_lo = w.germ_lo;
  uint32_t* germ_hi = w.germ_hi;
  const uint32_t* ev_off = w.ev_off.data();
  const uint32_t* pre_end = t.pre_end.data();

  // The numbering is a recursion over (subtree, allele copy):
  //   walk(node, first event, fragment set, allele): the node's events from `first event` on -- a SID the allele
  //   still holds opens an instance, a deletion changes the fragment set, an amplification or a WGD walks the
  //   SAME subtree for the new copy from the next event on, before anything else happens -- then the node's leaf,
  //   then its children; every instance opened in a node is carried by the haplotypes numbered until its subtree
  //   ends.
  // In preorder a subtree is the range [i, pre_end[i]), so one walk is a scan over consecutive positions.  What a
  // node changed (fragment set, opened instances) is put on `undo` and taken back when the scan reaches the end
  // of its subtree.  Nothing happens at a node without events of this chromosome but its leaf, so the scan goes
  // from one node with events to the next (ev_nodes) and numbers the sampled cells in between (leaf_pos) in one
  // tight loop: the cost is O(events + haplotypes), not O(nodes) per allele copy.  Only copies (rare) start a
  // nested scan.
  struct Scan {
    uint32_t p, end;        // next position to look at, end of the subtree this scan covers
    uint32_t k;             // kNodeStart, or the next event of node p (the scan was interrupted by a copy)
    uint32_t fs;            // fragment set of the allele at p
    uint32_t undo_base;     // undo entries below this one belong to enclosing scans
    uint32_t node_fs, node_open;  // state when the events of node p began
    uint32_t li, ei;        // first sampled cell / first node with events at or after p
    uint32_t root;          // ordinal of the root this scan started at, if it is the root's own scan
    uint16_t allele;
    bool root_base, preneo_done;
  };
  struct Undo {
    uint32_t end, fs, open_size;
  };
  constexpr uint32_t kNodeStart = 0xffffffffu, kNever = 0xffffffffu;
  std::vector<Scan> scans;
  std::vector<Undo> undo;
  std::vector<uint32_t> open;  // instances whose interval is still growing
  std::vector<uint8_t> row_open(w.row_hi - w.row_lo, 0);  // the row has an open somatic instance
  // two germline rows of one position must not share an allele
  if (w.row_hi > w.row_lo)
    for (uint32_t l = row_locus[w.row_lo]; l <= row_locus[w.row_hi - 1]; ++l) {
      uint32_t seen = 0;
      for (uint32_t m = locus_first_row[l]; m < locus_first_row[l + 1]; ++m) {
        const uint32_t mask = row_mask[m].load(std::memory_order_relaxed);
        check((seen & mask) == 0, "two germline SIDs at one position of one allele");
        seen |= mask;
      }
    }
  const uint32_t* leaf_pos = t.leaf_pos.data();
  const uint32_t* leaf_id = t.leaf_id.data();
  const uint32_t n_leaf_pos = static_cast<uint32_t>(t.leaf_pos.size());
  const uint32_t* ev_nodes = w.ev_nodes.data();
  const uint32_t n_ev_nodes = static_cast<uint32_t>(w.ev_nodes.size());
  auto close_to = [&](uint32_t open_size) {  // instances opened since are carried by [lo, counter)
    for (size_t k = open_size; k < open.size(); ++k) {
      Inst& in = w.inst[open[k]];
      in.span = counter - in.lo;
      row_open[in.row - w.row_lo] = 0;
    }
    open.resize(open_size);
  };

  w.inst.reserve(w.ev.size());
  for (uint16_t g = 0; g < n0; ++g) {
    germ_lo[g] = counter;
    w.haps.push_back({0u, full, g, HAP_NORMAL_PLAIN});
    ++counter;
    for (uint32_t ri = 0; ri < t.root_pos.size(); ++ri) {
      const uint32_t r = t.root_pos[ri];
      scans.push_back({r, pre_end[r], kNodeStart, full, static_cast<uint32_t>(undo.size()), full,
                       static_cast<uint32_t>(open.size()),
                       static_cast<uint32_t>(std::lower_bound(leaf_pos, leaf_pos + n_leaf_pos, r) - leaf_pos),
                       static_cast<uint32_t>(std::lower_bound(ev_nodes, ev_nodes + n_ev_nodes, r) - ev_nodes), ri, g, true,
                       false});
      while (!scans.empty()) {
        Scan& sc = scans.back();
        uint32_t k = sc.k;
        if (k == kNodeStart) {
          while (undo.size() > sc.undo_base && undo.back().end <= sc.p) {  // subtrees that ended at or before p
            close_to(undo.back().open_size);
            sc.fs = undo.back().fs;
            undo.pop_back();
          }
          if (sc.p >= sc.end) {  // this subtree is done (its undo entries all ended at or before sc.end)
            scans.pop_back();
            continue;
          }
          const uint32_t next_ev = sc.ei < n_ev_nodes ? ev_nodes[sc.ei] : kNever;
          if (next_ev != sc.p && !(sc.root_base && sc.p == r)) {
            // nodes without events: number their sampled cells up to the next stop
            uint32_t stop = std::min(sc.end, next_ev);
            if (undo.size() > sc.undo_base) stop = std::min(stop, undo.back().end);
            uint32_t li = sc.li;
            const uint32_t fs = sc.fs;
            const uint16_t allele = sc.allele;
            for (; li < n_leaf_pos && leaf_pos[li] < stop; ++li) w.haps.push_back({leaf_id[li], fs, allele, HAP_TUMOUR});
            counter += li - sc.li;
            sc.li = li;
            sc.p = stop;
            continue;
          }
          k = ev_off[sc.p];
          sc.node_fs = sc.fs;
          sc.node_open = static_cast<uint32_t>(open.size());
        } else {
          sc.k = kNodeStart;
        }
        const uint32_t p = sc.p;
        const bool root_node = sc.root_base && p == r;
        auto preneo_leaf = [&]() {
          sc.preneo_done = true;
          w.haps.push_back({sc.root, sc.fs, sc.allele, HAP_NORMAL_PRENEO});
          ++counter;
        };
        bool copied = false;
        for (const uint32_t k_end = ev_off[p + 1]; k < k_end;) {
          const Ev e = w.ev[k++];
          if (root_node && !sc.preneo_done && !(e.kind == PCS_EV_SID && e.nature == PCS_NATURE_PRENEOPLASTIC))
            preneo_leaf();
          uint32_t copy_fs = 0;
          uint16_t copy_allele = 0;
          if (e.kind == PCS_EV_WGD) {
            auto it = w.wgd_map.find(e.x);
            if (it == w.wgd_map.end()) continue;
            bool mine = false;
            for (const auto& [a, nd] : it->second)
              if (a == sc.allele) {
                mine = true;
                copy_fs = sc.fs;
                copy_allele = nd;
                break;
              }
            if (!mine) continue;
          } else {
            if (e.allele != sc.allele) continue;
            if (e.kind == PCS_EV_SID) {
              if (holds(w.fragsets[sc.fs], e.y)) {
                const uint32_t l = row_locus[e.x];
                for (uint32_t m = locus_first_row[l]; m < locus_first_row[l + 1]; ++m) {
                  check(!((row_mask[m].load(std::memory_order_relaxed) >> g) & 1u),
                        "somatic SID at a germline SID position of the same allele");
                  check(!row_open[m - w.row_lo], "two SIDs at one position of one allele");
                }
                row_open[e.x - w.row_lo] = 1;
                open.push_back(static_cast<uint32_t>(w.inst.size()));
                w.inst.push_back({counter, 0u, e.x, e.meta});
              }
              continue;
            }
            if (e.kind == PCS_EV_CNA_DEL) {
              sc.fs = w.intern_set(remove_range(w.fragsets[sc.fs], e.x, e.x + e.y - 1, clen));
              continue;
            }
            if (e.kind != PCS_EV_CNA_AMP) throw std::domain_error("unknown event kind");
            copy_fs = w.intern_set(clip(w.fragsets[sc.fs], e.x, e.x + e.y - 1));
            copy_allele = e.dest;
          }
          // the copy: this subtree once more, from the next event on, before this scan goes on
          sc.k = k;
          const Scan copy{p, pre_end[p], k, copy_fs, static_cast<uint32_t>(undo.size()), copy_fs,
                          static_cast<uint32_t>(open.size()), sc.li, sc.ei, 0u, copy_allele, false, true};
          scans.push_back(copy);  // invalidates sc
          copied = true;
          break;
        }
        if (copied) continue;
        if (root_node && !sc.preneo_done) preneo_leaf();
        if (sc.fs != sc.node_fs || open.size() != sc.node_open)
          undo.push_back({pre_end[p], sc.node_fs, sc.node_open});
        const bool dead = w.fs_empty[sc.fs] != 0;  // no DNA left: nothing below can be read
        if (sc.ei < n_ev_nodes && ev_nodes[sc.ei] == p) ++sc.ei;
        if (sc.li < n_leaf_pos && leaf_pos[sc.li] == p) {  // a sampled cell (always a leaf of the tree)
          if (!dead) {
            w.haps.push_back({leaf_id[sc.li], sc.fs, sc.allele, HAP_TUMOUR});
            ++counter;
          }
          ++sc.li;
        }
        if (dead && pre_end[p] != p + 1) {  // skip the subtree
          sc.p = pre_end[p];
          sc.li = static_cast<uint32_t>(std::lower_bound(leaf_pos + sc.li, leaf_pos + n_leaf_pos, sc.p) - leaf_pos);
          sc.ei = static_cast<uint32_t>(std::lower_bound(ev_nodes + sc.ei, ev_nodes + n_ev_nodes, sc.p) - ev_nodes);
        } else {
          sc.p = p + 1;
        }
      }
    }
    germ_hi[g] = counter;
  }

  finish_chr(w);
}

}  // namespace

namespace {
struct PhaseTimer {
  bool on = std::getenv("PCS_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    auto n = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[pcs flatten] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};
}  // namespace

// what the path-specific haplotype numbering of flatten_with() works on
struct Numbering {
  const pcs_forest_desc& d;
  FlatForest& out;
  std::vector<ChrWork>& work;              // [n_chr]: chr, row_lo, row_hi set; the numbering fills the rest
  const std::atomic<uint8_t>* row_mask;    // [n_mut] germline allele mask of every row
  const std::function<void(uint32_t, const std::function<void(uint32_t)>&)>& parallel_for;
  PhaseTimer& timer;
};

// The flattened view of a forest given in any form: the mutation table and the germline of `d` are turned into
// loci, germline masks and -- after `number` has produced, per chromosome, the haplotype leaves, the somatic
// placements (sorted by row), the interval below each germline allele, fragment sets and pieces -- the merged
// instance table.  `number` is the only part that depends on how the forest is described (an event-labelled
// tree: flatten_forest; explicit per-cell genomes: flatten_cell_genomes).
template <class NumberFn>
void flatten_with(const pcs_forest_desc& d, FlatForest& out, unsigned n_threads, const std::function<void()>& loci_ready,
                  bool defer_instances, NumberFn&& number) {
  PhaseTimer timer;
  check(d.n_chr >= 1 && d.n_chr < 65535, "n_chr out of range");
  {
    FlatStore keep = std::move(out.store);  // a block lent by the caller survives the reset
    keep.used = 0;
    keep.heap.clear();
    out = FlatForest{};
    out.store = std::move(keep);
  }
  out.n_chr = d.n_chr;
  out.n_mut = d.n_mut;
  out.n_leaves = d.n_leaves;
  out.n_samples = d.n_samples;
  out.chr_len.assign(d.chr_len, d.chr_len + d.n_chr);
  out.chr_n_alleles.assign(d.chr_n_alleles, d.chr_n_alleles + d.n_chr);
  out.leaf_sample.assign(d.leaf_sample, d.leaf_sample + d.n_leaves);
  for (uint32_t c = 0; c < d.n_chr; ++c) {
    check(d.chr_len[c] >= 1, "chromosome length must be positive");
    check(d.chr_len[c] < (1u << 31), "chromosome length must be below 2^31");  // 32-bit positions on the device
  }

  n_threads = std::max(1u, n_threads);
  // run fn(task) for task in [0, n_tasks) on up to n_threads threads; the first exception is rethrown
  auto parallel_for = [&](uint32_t n_tasks, const std::function<void(uint32_t)>& fn) {
    std::vector<std::string> errors(n_tasks);
    HostPool::get().run(n_tasks, [&](size_t k) {
      try {
        fn(static_cast<uint32_t>(k));
      } catch (const std::exception& e) {
        errors[k] = e.what();
        if (errors[k].empty()) errors[k] = "flatten failed";
      }
    }, n_threads);
    for (const auto& e : errors)
      if (!e.empty()) throw std::domain_error(e);
  };

  // ---- mutation table -> loci (validated and numbered in row chunks)
  const uint32_t n_chunks = d.n_mut ? std::min<uint32_t>(4 * n_threads, (d.n_mut + 65535) / 65536) : 0;
  auto chunk_lo = [&](uint32_t k) { return static_cast<uint32_t>(static_cast<uint64_t>(d.n_mut) * k / n_chunks); };
  std::vector<uint32_t> chunk_loci(n_chunks + 1, 0);
  {
    const uint16_t* mc = d.mut_chr;
    const uint32_t* mp = d.mut_pos;
    const uint32_t* clen = d.chr_len;
    const uint32_t n_chr = d.n_chr;
    parallel_for(n_chunks, [&, mc, mp, clen, n_chr](uint32_t k) {
      const uint32_t lo = chunk_lo(k), hi = chunk_lo(k + 1);
      uint32_t cnt = 0;
      bool ok_chr = true, ok_pos = true, ok_len = true, ok_order = true;
      // row lo is compared with its predecessor too (chunk edges are ordinary rows)
      uint32_t pc = lo ? mc[lo - 1] : 0, pp = lo ? mp[lo - 1] : 0;
      for (uint32_t m = lo; m < hi; ++m) {
        const uint32_t c = mc[m], p = mp[m];
        ok_chr &= c < n_chr;
        ok_pos &= p >= 1 && p <= clen[c < n_chr ? c : 0];
        ok_len &= d.mut_ref_len[m] >= 1 && d.mut_alt_len[m] >= 1;
        ok_order &= m == 0 || pc < c || (pc == c && pp <= p);
        cnt += (m == 0 || pc != c || pp != p) ? 1u : 0u;
        pc = c;
        pp = p;
      }
      check(ok_chr, "mut_chr out of range");
      check(ok_pos, "mutation position outside the chromosome");
      check(ok_len, "ref/alt must be non-empty");
      check(ok_order, "mutation table must be sorted by (chr, pos)");
      chunk_loci[k + 1] = cnt;
    });
  }
  for (uint32_t k = 0; k < n_chunks; ++k) chunk_loci[k + 1] += chunk_loci[k];
  const uint32_t n_loci = n_chunks ? chunk_loci[n_chunks] : 0;
  out.locus_pos = out.store.table<uint32_t>(n_loci);
  out.row_locus = out.store.table<uint32_t>(d.n_mut);
  out.locus_first_row = out.store.table<uint32_t>(static_cast<size_t>(n_loci) + 1);
  out.locus_first_row[n_loci] = d.n_mut;
  out.locus_inst_off = out.store.table<uint32_t>(static_cast<size_t>(n_loci) + 1);
  if (defer_instances) {  // what the device builds the instance table from (FlatForest::inst_deferred)
    out.row_meta = out.store.table<uint16_t>(d.n_mut);
    out.germ_mask = out.store.table<uint8_t>(static_cast<size_t>(d.n_mut) + 1);
  }
  {
    const uint16_t* mc = d.mut_chr;
    const uint32_t* mp = d.mut_pos;
    uint32_t* locus_pos = out.locus_pos.data();
    uint32_t* first_row = out.locus_first_row.data();
    uint32_t* row_locus = out.row_locus.data();
    uint16_t* row_meta = out.row_meta.data();  // null unless deferred
    uint8_t* germ_mask = out.germ_mask.data();
    parallel_for(n_chunks, [&, mc, mp, locus_pos, first_row, row_locus, row_meta, germ_mask](uint32_t k) {
      const uint32_t lo = chunk_lo(k), hi = chunk_lo(k + 1);
      uint32_t l = chunk_loci[k];  // loci before this chunk
      uint32_t pc = lo ? mc[lo - 1] : 0, pp = lo ? mp[lo - 1] : 0;
      for (uint32_t m = lo; m < hi; ++m) {
        const uint32_t c = mc[m], p = mp[m];
        if (m == 0 || pc != c || pp != p) {
          locus_pos[l] = p;
          first_row[l] = m;
          ++l;
        }
        row_locus[m] = l - 1;
        pc = c;
        pp = p;
      }
      if (row_meta) {
        for (uint32_t m = lo; m < hi; ++m)
          row_meta[m] = static_cast<uint16_t>(d.mut_ref_len[m] | (static_cast<uint16_t>(d.mut_alt_len[m]) << 8));
        std::memset(germ_mask + lo, 0, (k + 1 == n_chunks ? hi + 1 : hi) - lo);  // the mask bytes start out "not germline"
      }
    });
  }
  // rows and loci of every chromosome (the table is sorted by chromosome)
  std::vector<uint32_t> chr_row_off(d.n_chr + 1, d.n_mut);
  for (uint32_t c = 0; c <= d.n_chr; ++c)
    chr_row_off[c] = static_cast<uint32_t>(std::lower_bound(d.mut_chr, d.mut_chr + d.n_mut, static_cast<uint16_t>(c)) - d.mut_chr);
  chr_row_off[d.n_chr] = d.n_mut;
  out.chr_locus_off.assign(d.n_chr + 1, n_loci);
  for (uint32_t c = 0; c < d.n_chr; ++c)
    out.chr_locus_off[c] = chr_row_off[c] < d.n_mut ? out.row_locus[chr_row_off[c]] : n_loci;

  timer.lap("loci");
  if (loci_ready) loci_ready();  // locus_pos, row_locus, chr_locus_off are final: the caller may start copying them
  // ---- germline SIDs by row.  The caller's list comes in any order; a SID is normally listed once, so the
  // list is scattered into one allele-mask byte per row (0 = not germline; a row listed twice gets the union
  // of its masks).  Should a row be listed twice, a stably sorted copy of the list is walked for the instances
  // instead (same result, slower).  Built before the haplotype numbering, which checks somatic SIDs against it.
  const uint64_t G = d.n_germline;
  check(G <= 0xffffffffull, "too many germline SIDs");
  const uint32_t g_chunks = G ? static_cast<uint32_t>(std::min<uint64_t>(4 * n_threads, (G + 65535) / 65536)) : 0;
  // relaxed atomic ORs: a row listed twice may be written by two threads
  static_assert(sizeof(std::atomic<uint8_t>) == 1, "the mask bytes are sent to the device as they lie");
  std::unique_ptr<std::atomic<uint8_t>[]> row_mask_heap;
  std::atomic<uint8_t>* row_mask;
  if (defer_instances) {
    row_mask = reinterpret_cast<std::atomic<uint8_t>*>(out.germ_mask.data());  // zeroed by the loci pass
  } else {
    row_mask_heap.reset(new std::atomic<uint8_t>[static_cast<size_t>(d.n_mut) + 1]());
    row_mask = row_mask_heap.get();
  }
  {
    // Threads must not write mask bytes of the same cache line, or the line bounces between cores for every
    // entry.  A list sorted by row is scattered chunk by chunk as it lies; any other order is first partitioned
    // by row range (bucket = row >> shift, at most 64 of them), then every bucket is scattered by one thread.
    uint32_t shift = 6;
    while ((d.n_mut >> shift) >= 64) ++shift;
    const uint32_t n_buckets = (d.n_mut >> shift) + 1;
    // A list sorted by row (the order a shim's std::map<SID, ...> gives) is validated, checked for its order and
    // scattered in ONE pass: the listings of one row are neighbours, so a chunk gathers a row's masks in a register
    // and writes the byte once, with a plain store -- except for its first and its last row, which a neighbouring
    // chunk may hold listings of as well (atomic OR).  Should a chunk find the list out of order, the masks are
    // cleared and the general path below starts over.
    std::vector<uint8_t> chunk_sorted(g_chunks, 1);
    parallel_for(g_chunks, [&](uint32_t k) {
      const uint64_t lo = G * k / g_chunks, hi = G * (k + 1) / g_chunks;
      if (lo == hi) return;
      const uint32_t first_row = d.germ_mut[lo], last_row = d.germ_mut[hi - 1];
      check(first_row < d.n_mut && last_row < d.n_mut, "germ_mut out of range");
      uint32_t prev = lo ? d.germ_mut[lo - 1] : 0;
      bool sorted = true, in_range = true, named = true;
      uint32_t row = first_row;
      uint8_t acc = 0;
      auto flush = [&]() {
        if (row == first_row || row == last_row)
          row_mask[row].fetch_or(acc, std::memory_order_relaxed);
        else
          row_mask[row].store(acc, std::memory_order_relaxed);
      };
      for (uint64_t i = lo; i < hi; ++i) {
        const uint32_t m = d.germ_mut[i];
        if (m >= d.n_mut) {
          in_range = false;
          break;
        }
        named &= d.germ_allele_mask[i] != 0;
        sorted &= prev <= m;
        prev = m;
        if (m != row) {
          flush();
          row = m;
          acc = 0;
        }
        acc |= d.germ_allele_mask[i];
      }
      check(in_range, "germ_mut out of range");
      check(named, "germ_allele_mask names a missing allele");
      flush();
      chunk_sorted[k] = sorted;
    });
    if (std::find(chunk_sorted.begin(), chunk_sorted.end(), 0) != chunk_sorted.end()) {
      // any other order: start over -- count the entries of every bucket per chunk, partition, scatter by bucket
      parallel_for(g_chunks, [&](uint32_t k) {
        const size_t lo = (static_cast<size_t>(d.n_mut) + 1) * k / g_chunks, hi = (static_cast<size_t>(d.n_mut) + 1) * (k + 1) / g_chunks;
        for (size_t m = lo; m < hi; ++m) row_mask[m].store(0, std::memory_order_relaxed);
      });
      std::vector<uint32_t> hist(static_cast<size_t>(g_chunks) * n_buckets, 0);
      parallel_for(g_chunks, [&](uint32_t k) {
        const uint64_t lo = G * k / g_chunks, hi = G * (k + 1) / g_chunks;
        uint32_t h[64] = {0};  // on the stack: neighbouring rows of `hist` share cache lines
        for (uint64_t i = lo; i < hi; ++i) ++h[d.germ_mut[i] >> shift];
        std::copy(h, h + n_buckets, hist.data() + static_cast<size_t>(k) * n_buckets);
      });
      // hist -> where chunk k writes its entries of bucket b: buckets in order, chunks in order inside a bucket
      std::vector<uint64_t> bucket_off(n_buckets + 1, 0);
      {
        uint64_t run = 0;
        for (uint32_t b = 0; b < n_buckets; ++b) {
          bucket_off[b] = run;
          for (uint32_t k = 0; k < g_chunks; ++k) {
            uint32_t& h = hist[static_cast<size_t>(k) * n_buckets + b];
            const uint32_t cnt = h;
            h = static_cast<uint32_t>(run);
            run += cnt;
          }
        }
        bucket_off[n_buckets] = run;
      }
      check(d.n_mut <= (1u << 30), "too many rows for an unsorted germline list: sort it by row");
      std::unique_ptr<uint32_t[]> part(new uint32_t[G]);  // row | mask << 30
      parallel_for(g_chunks, [&](uint32_t k) {
        const uint64_t lo = G * k / g_chunks, hi = G * (k + 1) / g_chunks;
        uint32_t at[64];
        std::copy(hist.data() + static_cast<size_t>(k) * n_buckets, hist.data() + static_cast<size_t>(k + 1) * n_buckets, at);
        for (uint64_t i = lo; i < hi; ++i) {
          const uint32_t m = d.germ_mut[i];
          const uint32_t q = at[m >> shift]++;
          part[q] = m | static_cast<uint32_t>(d.germ_allele_mask[i]) << 30;
        }
      });
      parallel_for(n_buckets, [&](uint32_t b) {  // bucket = 2^shift rows, shift >= 6: whole cache lines
        for (uint64_t i = bucket_off[b]; i < bucket_off[b + 1]; ++i)
          row_mask[part[i] & 0x3fffffffu].fetch_or(static_cast<uint8_t>(part[i] >> 30), std::memory_order_relaxed);
      });
    }
  }
  timer.lap("germline masks");
  std::vector<ChrWork> work(d.n_chr);
  for (uint32_t c = 0; c < d.n_chr; ++c) {
    work[c].chr = c;
    work[c].clen = d.chr_len[c];
    work[c].row_lo = chr_row_off[c];
    work[c].row_hi = chr_row_off[c + 1];
  }
  {
    const std::function<void(uint32_t, const std::function<void(uint32_t)>&)> pf = parallel_for;
    Numbering nb{d, out, work, row_mask, pf, timer};
    number(nb);
  }
  timer.lap("haplotype numbering");

  // chunks of loci (a locus never straddles two chunks); germline rows per chunk
  const uint32_t m_chunks = n_loci ? std::min<uint32_t>(4 * n_threads, (n_loci + 65535) / 65536) : 0;
  auto chunk_locus = [&](uint32_t k) { return static_cast<uint32_t>(static_cast<uint64_t>(n_loci) * k / m_chunks); };
  std::vector<uint64_t> germ_before(m_chunks + 1, 0);
  parallel_for(m_chunks, [&](uint32_t k) {
    uint64_t cnt = 0;
    const uint32_t r0 = out.locus_first_row[chunk_locus(k)], r1 = out.locus_first_row[chunk_locus(k + 1)];
    for (uint32_t m = r0; m < r1; ++m) cnt += row_mask[m].load(std::memory_order_relaxed) != 0;
    germ_before[k + 1] = cnt;
  });
  for (uint32_t k = 0; k < m_chunks; ++k) germ_before[k + 1] += germ_before[k];
  const bool listed_once = germ_before[m_chunks] == G;
  out.chr_row_off = chr_row_off;
  out.germ_iv.assign(static_cast<size_t>(d.n_chr) * 8, 0);
  for (uint32_t c = 0; c < d.n_chr; ++c) {  // haplotype interval of every germline allele mask: 1, 2, 3 = both
    const ChrWork& w = work[c];
    uint32_t* iv = out.germ_iv.data() + static_cast<size_t>(c) * 8;
    iv[1] = w.germ_lo[0]; iv[2] = w.germ_lo[1]; iv[3] = w.germ_lo[0];
    iv[5] = w.germ_hi[0] - w.germ_lo[0]; iv[6] = w.germ_hi[1] - w.germ_lo[1]; iv[7] = w.germ_hi[1] - w.germ_lo[0];
  }
  std::vector<uint32_t> g_mut;   // only when a row is listed more than once
  std::vector<uint8_t> g_mask;
  if (!listed_once) {
    std::vector<uint64_t> key(G);  // (row, list index): a plain sort keeps the list order inside a row
    for (uint64_t i = 0; i < G; ++i) key[i] = (static_cast<uint64_t>(d.germ_mut[i]) << 32) | i;
    std::sort(key.begin(), key.end());
    g_mut.resize(G);
    g_mask.resize(G);
    for (uint64_t i = 0; i < G; ++i) {
      g_mut[i] = static_cast<uint32_t>(key[i] >> 32);
      g_mask[i] = d.germ_allele_mask[key[i] & 0xffffffffull];
      if (i > 0 && g_mut[i] == g_mut[i - 1])  // masks of one row must name different alleles
        for (uint64_t j = i; j-- > 0 && g_mut[j] == g_mut[i];)
          check((g_mask[j] & g_mask[i]) == 0, "a germline SID is listed twice for one allele");
    }
    for (uint32_t k = 0; k <= m_chunks; ++k)
      germ_before[k] = std::lower_bound(g_mut.begin(), g_mut.end(), out.locus_first_row[chunk_locus(k)]) - g_mut.begin();
  }
  timer.lap("germline by row");

  // ---- instances: somatic placements (few, already sorted by row inside each chromosome) merged with the
  // germline, written once, in place, by chunks of loci; every chunk knows where its output starts
  std::vector<Inst> som;
  {
    size_t n_som = 0;
    for (const auto& w : work) n_som += w.inst.size();
    som.reserve(n_som);
    for (const auto& w : work) som.insert(som.end(), w.inst.begin(), w.inst.end());  // chromosome-major = row-major
  }
  check(som.size() + G <= 0xffffffffull, "too many SID placements");
  const size_t n_inst = som.size() + G;
  out.n_inst = n_inst;
  if (defer_instances && listed_once) {
    // the device builds inst / locus_inst_off (FlatForest::inst_deferred); what the merge below would have refused
    // is refused here: a mask naming an allele the chromosome does not have (or one past the two germline alleles)
    parallel_for(d.n_chr, [&](uint32_t c) {
      const uint32_t bad = (0xffu << d.chr_n_alleles[c]) | 0xfcu;
      uint32_t any = 0;
      for (uint32_t m = chr_row_off[c]; m < chr_row_off[c + 1]; ++m) any |= row_mask[m].load(std::memory_order_relaxed) & bad;
      check(any == 0, "germ_allele_mask names a missing allele");
    });
    out.inst_deferred = true;
    out.som = std::move(som);
    out.locus_inst_off = Table<uint32_t>{};
    timer.lap("instances (deferred)");
  } else {
  out.inst = out.store.table<Inst>(n_inst);
  out.locus_inst_off[n_loci] = static_cast<uint32_t>(n_inst);
  auto row_less = [](const Inst& in, uint32_t row) { return in.row < row; };
  parallel_for(m_chunks, [&](uint32_t k) {
    const uint32_t l0 = chunk_locus(k), l1 = chunk_locus(k + 1);
    size_t si = std::lower_bound(som.begin(), som.end(), out.locus_first_row[l0], row_less) - som.begin();
    const size_t s1 = std::lower_bound(som.begin(), som.end(), out.locus_first_row[l1], row_less) - som.begin();
    size_t gi = germ_before[k];
    const size_t g1 = germ_before[k + 1];
    size_t pos = gi + si;  // the instances of every row before this chunk
    Inst* inst = out.inst.data();
    auto germline = [&](uint32_t m, uint8_t mask) {
      const ChrWork& w = work[d.mut_chr[m]];
      check((mask >> d.chr_n_alleles[w.chr]) == 0, "germ_allele_mask names a missing allele");
      const uint32_t meta = static_cast<uint32_t>(d.mut_ref_len[m]) | (static_cast<uint32_t>(d.mut_alt_len[m]) << 8);
      if (mask == 3) {
        inst[pos++] = {w.germ_lo[0], w.germ_hi[1] - w.germ_lo[0], m, meta};
      } else {
        const uint32_t g = mask == 1 ? 0 : 1;
        inst[pos++] = {w.germ_lo[g], w.germ_hi[g] - w.germ_lo[g], m, meta};
      }
    };
    if (listed_once) {
      // the common case, a row at a time with everything a chromosome fixes hoisted out of the loop: the interval
      // of each allele mask (1, 2, 3 = both), the next somatic row, the row of the next locus
      const uint32_t* first_row = out.locus_first_row.data();
      uint32_t* inst_off = out.locus_inst_off.data();
      const uint8_t* ref_len = d.mut_ref_len;
      const uint8_t* alt_len = d.mut_alt_len;
      uint32_t l = l0;
      const uint32_t m_end = first_row[l1];
      for (uint32_t m = first_row[l0]; m < m_end;) {
        const ChrWork& w = work[d.mut_chr[m]];
        const uint32_t chr_end = std::min(w.row_hi, m_end);
        const uint32_t bad = 0xffu << d.chr_n_alleles[w.chr];  // mask bits of alleles the chromosome does not have
        const uint32_t iv_lo[4] = {0u, w.germ_lo[0], w.germ_lo[1], w.germ_lo[0]};
        const uint32_t iv_n[4] = {0u, w.germ_hi[0] - w.germ_lo[0], w.germ_hi[1] - w.germ_lo[1], w.germ_hi[1] - w.germ_lo[0]};
        uint32_t any_bad = 0;
        uint32_t next_som = si < s1 ? som[si].row : 0xffffffffu;
        uint32_t next_locus_row = first_row[l];
        for (; m < chr_end; ++m) {
          if (m == next_locus_row) {
            inst_off[l] = static_cast<uint32_t>(pos);
            next_locus_row = first_row[++l];
          }
          if (m == next_som) {
            do inst[pos++] = som[si++]; while (si < s1 && som[si].row == m);
            next_som = si < s1 ? som[si].row : 0xffffffffu;
          }
          const uint32_t mask = row_mask[m].load(std::memory_order_relaxed);
          if (mask) {
            any_bad |= mask & bad;
            const uint32_t k = mask & 3u;  // bits above the two germline alleles are refused below
            any_bad |= mask & ~3u;
            inst[pos++] = {iv_lo[k], iv_n[k], m, static_cast<uint32_t>(ref_len[m]) | (static_cast<uint32_t>(alt_len[m]) << 8)};
            ++gi;
          }
        }
        check(any_bad == 0, "germ_allele_mask names a missing allele");
      }
    } else {
      for (uint32_t l = l0; l < l1; ++l) {
        out.locus_inst_off[l] = static_cast<uint32_t>(pos);
        const uint32_t row_end = out.locus_first_row[l + 1];
        for (uint32_t m = out.locus_first_row[l]; m < row_end; ++m) {
          while (si < s1 && som[si].row == m) inst[pos++] = som[si++];
          while (gi < g1 && g_mut[gi] == m) germline(m, g_mask[gi++]);
        }
      }
    }
    check(gi == g1 && si == s1 && pos == g1 + s1, "internal: instance merge out of step");
  });
  timer.lap("instances");
  }

  // ---- merge the per-chromosome haplotypes, fragment sets and pieces
  out.chr_haps.resize(d.n_chr);
  out.full_fragset.resize(d.n_chr);
  out.chr_piece_off.assign(d.n_chr + 1, 0);
  std::vector<uint32_t> fs_base(d.n_chr + 1, 0), cover_base(d.n_chr + 1, 0);
  for (uint32_t c = 0; c < d.n_chr; ++c) {
    fs_base[c + 1] = fs_base[c] + static_cast<uint32_t>(work[c].fragsets.size());
    cover_base[c + 1] = cover_base[c] + static_cast<uint32_t>(work[c].covers.size());
    out.chr_piece_off[c + 1] = out.chr_piece_off[c] + static_cast<uint32_t>(work[c].pieces.size());
  }
  out.fragsets.resize(fs_base[d.n_chr]);
  out.covers.resize(cover_base[d.n_chr]);
  out.pieces.resize(out.chr_piece_off[d.n_chr]);
  parallel_for(d.n_chr, [&](uint32_t c) {
    ChrWork& w = work[c];
    const uint32_t base = fs_base[c];
    for (auto& h : w.haps) h.fragset += base;
    out.chr_haps[c] = std::move(w.haps);
    for (size_t k = 0; k < w.fragsets.size(); ++k) {
      std::vector<Frag>& fr = out.fragsets[base + k];
      fr.reserve(w.fragsets[k].size());
      for (const auto& p : w.fragsets[k]) fr.push_back({p.first, p.second});
    }
    out.full_fragset[c] = base;  // interned first in flatten_chr
    for (size_t k = 0; k < w.covers.size(); ++k)
      out.covers[cover_base[c] + k] = {w.covers[k].fragset + base, w.covers[k].frag_end};
    for (size_t k = 0; k < w.pieces.size(); ++k) {
      Piece pc = w.pieces[k];
      pc.cover_off += cover_base[c];
      out.pieces[out.chr_piece_off[c] + k] = pc;
    }
  });
  timer.lap("merge");
}

void flatten_forest(const pcs_forest_desc& d, FlatForest& out, unsigned n_threads, const std::function<void()>& loci_ready,
                    bool defer_instances) {
  check(d.n_nodes >= 1, "the forest has no nodes");
  flatten_with(d, out, n_threads, loci_ready, defer_instances, [&](Numbering& nb) {
    FlatForest& out = nb.out;
    std::vector<ChrWork>& work = nb.work;
    const std::atomic<uint8_t>* row_mask = nb.row_mask;
    const auto& parallel_for = nb.parallel_for;
    PhaseTimer& timer = nb.timer;
    // ---- cell tree
    Tree t;
    t.child_off.assign(d.n_nodes + 1, 0);
    for (uint32_t v = 0; v < d.n_nodes; ++v) {
      int32_t p = d.node_parent[v];
      if (p < 0) {
        t.roots.push_back(v);
      } else {
        check(static_cast<uint32_t>(p) < v, "node_parent must precede the child");
        ++t.child_off[p + 1];
      }
    }
    for (uint32_t v = 0; v < d.n_nodes; ++v) t.child_off[v + 1] += t.child_off[v];
    t.child_idx.resize(t.child_off[d.n_nodes]);
    {
      std::vector<uint32_t> fill(t.child_off.begin(), t.child_off.end() - 1);
      for (uint32_t v = 0; v < d.n_nodes; ++v)
        if (d.node_parent[v] >= 0) t.child_idx[fill[d.node_parent[v]]++] = v;
    }
    out.n_roots = static_cast<uint32_t>(t.roots.size());
    t.node_leaf.assign(d.n_nodes, -1);
    for (uint32_t l = 0; l < d.n_leaves; ++l) {
      check(d.leaf_node[l] < d.n_nodes, "leaf_node out of range");
      check(t.child_off[d.leaf_node[l] + 1] == t.child_off[d.leaf_node[l]], "a sampled cell must be a leaf");
      check(d.leaf_sample[l] < d.n_samples, "leaf_sample out of range");
      t.node_leaf[d.leaf_node[l]] = l;
    }

    // preorder layout
    {
      const uint32_t n = d.n_nodes;
      t.pre_node.resize(n);
      t.pre_end.resize(n);
      t.pre_leaf.resize(n);
      std::vector<uint32_t> pos_of(n);
      std::vector<uint32_t> stack;
      uint32_t next = 0;
      for (uint32_t r : t.roots) {
        t.root_pos.push_back(next);
        stack.push_back(r);
        while (!stack.empty()) {
          const uint32_t v = stack.back();
          stack.pop_back();
          pos_of[v] = next;
          t.pre_node[next] = v;
          t.pre_leaf[next] = t.node_leaf[v];
          ++next;
          for (uint32_t c = t.child_off[v + 1]; c > t.child_off[v]; --c) stack.push_back(t.child_idx[c - 1]);  // smallest on top
        }
      }
      check(next == n, "internal: preorder does not cover the tree");
      // subtree ends: a node's subtree ends where its last child's does; children come after their parent
      for (uint32_t i = n; i-- > 0;) {
        const uint32_t v = t.pre_node[i];
        const uint32_t nc = t.child_off[v + 1] - t.child_off[v];
        t.pre_end[i] = nc ? t.pre_end[pos_of[t.child_idx[t.child_off[v + 1] - 1]]] : i + 1;
      }
      t.leaf_pos.reserve(d.n_leaves);
      t.leaf_id.reserve(d.n_leaves);
      for (uint32_t i = 0; i < n; ++i)
        if (t.pre_leaf[i] >= 0) {
          t.leaf_pos.push_back(i);
          t.leaf_id.push_back(static_cast<uint32_t>(t.pre_leaf[i]));
        }
    }
    timer.lap("cell tree");
    // ---- events: validated and dealt to their chromosomes in WALK order (by preorder position of the node, then
    // in the node's own order), with what the walk needs of the rows they name.  One counting pass and one filling
    // pass over chunks of preorder positions: every event is read twice, whatever the number of chromosomes; a
    // WGD goes to every chromosome.
    check(d.node_event_off[0] == 0 && d.node_event_off[d.n_nodes] == d.n_events, "node_event_off is not a CSR of the events");
    check(d.n_events <= 0xffffffffull, "too many events");
    for (uint32_t v = 0; v < d.n_nodes; ++v)
      check(d.node_event_off[v] <= d.node_event_off[v + 1], "node_event_off must be non-decreasing");
    std::vector<uint64_t> chr_load(d.n_chr, 0);
    {
      const uint32_t n = d.n_nodes, n_chr = d.n_chr;
      const uint32_t p_chunks = std::max(1u, std::min<uint32_t>(4 * n_threads, (n + 4095) / 4096));
      auto pos_lo = [&](uint32_t k) { return static_cast<uint32_t>(static_cast<uint64_t>(n) * k / p_chunks); };
      parallel_for(n_chr, [&](uint32_t c) { work[c].ev_off.assign(static_cast<size_t>(n) + 1, 0); });
      std::atomic<bool> any_wgd{false};
      parallel_for(p_chunks, [&](uint32_t k) {  // ev_off[c][i + 1] = events of chromosome c in the node at position i
        bool wgd = false;
        for (uint32_t i = pos_lo(k); i < pos_lo(k + 1); ++i) {
          const uint32_t v = t.pre_node[i];
          for (uint64_t e = d.node_event_off[v]; e < d.node_event_off[v + 1]; ++e) {
            const uint8_t kind = d.ev_kind[e];
            check(kind <= PCS_EV_WGD, "unknown event kind");
            if (kind == PCS_EV_WGD) {
              wgd = true;
              for (uint32_t c = 0; c < n_chr; ++c) ++work[c].ev_off[i + 1];
            } else {
              check(d.ev_chr[e] < n_chr, "event chromosome out of range");
              ++work[d.ev_chr[e]].ev_off[i + 1];
            }
          }
        }
        if (wgd) any_wgd.store(true, std::memory_order_relaxed);
      });
      parallel_for(n_chr, [&](uint32_t c) {
        ChrWork& w = work[c];
        for (uint32_t i = 0; i < n; ++i) {
          if (w.ev_off[i + 1]) w.ev_nodes.push_back(i);
          w.ev_off[i + 1] += w.ev_off[i];
        }
        w.ev.resize(w.ev_off[n]);
        w.has_wgd = any_wgd.load(std::memory_order_relaxed);
        chr_load[c] = w.ev_off[n];
      });
      parallel_for(p_chunks, [&](uint32_t k) {
        const uint32_t i0 = pos_lo(k), i1 = pos_lo(k + 1);
        std::vector<uint32_t> at(n_chr);  // where the next event of each chromosome goes
        for (uint32_t c = 0; c < n_chr; ++c) at[c] = work[c].ev_off[i0];
        for (uint32_t i = i0; i < i1; ++i) {
          const uint32_t v = t.pre_node[i];
          for (uint64_t e = d.node_event_off[v]; e < d.node_event_off[v + 1]; ++e) {
            const uint8_t kind = d.ev_kind[e];
            if (kind == PCS_EV_WGD) {
              for (uint32_t c = 0; c < n_chr; ++c)
                work[c].ev[at[c]++] = Ev{static_cast<uint32_t>(e), 0u, 0u, 0, 0, kind, d.ev_nature[e]};
              continue;
            }
            const uint32_t c = d.ev_chr[e];
            if (kind == PCS_EV_SID) {  // position and lengths of the row: looked up per chromosome (flatten_chr)
              work[c].ev[at[c]++] = Ev{d.ev_mut[e], 0u, 0u, d.ev_allele[e], 0, kind, d.ev_nature[e]};
            } else {
              check(d.ev_len[e] >= 1, "CNA length must be positive");
              work[c].ev[at[c]++] = Ev{d.ev_pos[e], d.ev_len[e], 0u, d.ev_allele[e], d.ev_dest[e], kind, d.ev_nature[e]};
            }
          }
        }
        for (uint32_t c = 0; c < n_chr; ++c)
          check(at[c] == work[c].ev_off[i1], "internal: events dealt out of step");
      });
    }
    timer.lap("events by chromosome");
    // ---- per-chromosome haplotype numbering, chromosomes in parallel (most events first)
    std::vector<uint32_t> chr_order(d.n_chr);
    for (uint32_t c = 0; c < d.n_chr; ++c) chr_order[c] = c;
    std::sort(chr_order.begin(), chr_order.end(), [&](uint32_t a, uint32_t b) {
      return chr_load[a] != chr_load[b] ? chr_load[a] > chr_load[b] : a < b;
    });
    parallel_for(d.n_chr, [&](uint32_t k) { flatten_chr(d, t, work[chr_order[k]], row_mask, out.row_locus.data(), out.locus_first_row.data()); });

  });
}

// ---------------------------------------------------------------- explicit per-cell genomes
// What the seam of the reference actually hands over (src/seq_simulation.cpp:566-575): per sample a list of per-cell
// genomes, chromosome -> allele -> fragment -> SID (traversal idiom src/phylogenetic_forest.cpp:279-290, 332-335),
// plus the normal sample's cells.  No tree comes with them, so the haplotype numbering is recovered from the
// contents: the carriers of a SID form a clade of the (unknown) haplotype tree, clades are nested, and sorting
// the haplotypes lexicographically by their SID lists -- every list ordered by decreasing number of carriers --
// lays every clade out as one run of consecutive haplotypes: a haplotype's list then starts with the chain of
// clades it belongs to, outermost first, so two members of a clade share the whole prefix down to it.  A SID whose
// carriers are NOT one clade (a later CNA deletion took it from part of one) simply gets one placement per run.
namespace {

void number_from_genomes(const pcs_cell_genomes_desc& g, const std::vector<uint64_t>& alleles, ChrWork& w,
                         const pcs_forest_desc& d, const std::atomic<uint8_t>* row_mask, const uint32_t* row_locus,
                         const uint32_t* locus_first_row) {
  const uint32_t chr = w.chr, clen = w.clen;
  const uint8_t n0 = d.chr_n_alleles[chr];
  check(n0 >= 1 && n0 <= 2, "chr_n_alleles must be 1 or 2");
  const uint32_t n_rows = w.row_hi - w.row_lo;
  const uint32_t full = w.intern_set(FragKey{{1u, clen}});
  // two germline rows of one position must not share an allele
  if (n_rows)
    for (uint32_t l = row_locus[w.row_lo]; l <= row_locus[w.row_hi - 1]; ++l) {
      uint32_t seen = 0;
      for (uint32_t m = locus_first_row[l]; m < locus_first_row[l + 1]; ++m) {
        const uint32_t mask = row_mask[m].load(std::memory_order_relaxed);
        check((seen & mask) == 0, "two germline SIDs at one position of one allele");
        seen |= mask;
      }
    }
  // carriers per row, and the rank of every row when rows are ordered by (carriers descending, row)
  std::vector<uint32_t> carriers(n_rows, 0);
  for (uint64_t a : alleles)
    for (uint64_t k = g.allele_sid_off[a]; k < g.allele_sid_off[a + 1]; ++k) {
      const uint32_t m = g.sid_row[k];
      check(m >= w.row_lo && m < w.row_hi, "an allele carries a SID of another chromosome");
      ++carriers[m - w.row_lo];
    }
  std::vector<uint32_t> by_count(n_rows), rank(n_rows);
  for (uint32_t i = 0; i < n_rows; ++i) by_count[i] = i;
  std::stable_sort(by_count.begin(), by_count.end(), [&](uint32_t x, uint32_t y) { return carriers[x] > carriers[y]; });
  for (uint32_t i = 0; i < n_rows; ++i) rank[by_count[i]] = i;
  // haplotypes: the normal cell's germline alleles (no somatic SID: they sort first in their block) and every
  // allele that still has DNA
  struct Hap {
    uint32_t cell, fragset;
    uint16_t allele;
    uint8_t kind, origin;
    uint64_t key_off;
    uint32_t key_n;
  };
  std::vector<Hap> haps;
  std::vector<uint32_t> keys;  // ranks of the SIDs of every haplotype, ascending
  for (uint16_t a0 = 0; a0 < n0; ++a0) haps.push_back({0u, full, a0, HAP_NORMAL_PLAIN, static_cast<uint8_t>(a0), 0, 0});
  std::vector<uint32_t> positions;
  for (uint64_t a : alleles) {
    check(g.allele_origin[a] < n0, "allele_origin names a missing germline allele");
    FragKey fk;
    for (uint64_t k = g.allele_frag_off[a]; k < g.allele_frag_off[a + 1]; ++k) {
      check(g.frag_begin[k] >= 1 && g.frag_begin[k] <= g.frag_end[k] && g.frag_end[k] <= clen, "fragment outside the chromosome");
      check(fk.empty() || fk.back().second < g.frag_begin[k], "the fragments of an allele must be disjoint and sorted");
      fk.emplace_back(g.frag_begin[k], g.frag_end[k]);
    }
    if (fk.empty()) continue;  // an allele that lost all its DNA is never read
    const uint32_t cell = g.allele_cell[a];
    check(cell < g.n_cells + g.n_normal_preneo, "allele_cell out of range");
    Hap h{cell < g.n_cells ? cell : cell - g.n_cells, w.intern_set(fk), g.allele_id[a],
          static_cast<uint8_t>(cell < g.n_cells ? HAP_TUMOUR : HAP_NORMAL_PRENEO), g.allele_origin[a], keys.size(), 0};
    positions.clear();
    for (uint64_t k = g.allele_sid_off[a]; k < g.allele_sid_off[a + 1]; ++k) {
      const uint32_t m = g.sid_row[k], pos = d.mut_pos[m];
      check(holds(fk, pos), "an allele carries a SID outside its fragments");
      // one SID per position of a haplotype: not on a germline SID of the allele it descends from ...
      const uint32_t l = row_locus[m];
      for (uint32_t q = locus_first_row[l]; q < locus_first_row[l + 1]; ++q)
        check(!((row_mask[q].load(std::memory_order_relaxed) >> h.origin) & 1u),
              "somatic SID at a germline SID position of the same allele");
      positions.push_back(pos);
      keys.push_back(rank[m - w.row_lo]);
    }
    std::sort(positions.begin(), positions.end());  // ... and no two somatic ones
    check(std::adjacent_find(positions.begin(), positions.end()) == positions.end(), "two SIDs at one position of one allele");
    h.key_n = static_cast<uint32_t>(keys.size() - h.key_off);
    std::sort(keys.begin() + static_cast<std::ptrdiff_t>(h.key_off), keys.end());
    haps.push_back(h);
  }
  std::vector<uint32_t> order(haps.size());
  for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    const Hap &hx = haps[x], &hy = haps[y];
    if (hx.origin != hy.origin) return hx.origin < hy.origin;
    return std::lexicographical_compare(keys.begin() + static_cast<std::ptrdiff_t>(hx.key_off),
                                        keys.begin() + static_cast<std::ptrdiff_t>(hx.key_off + hx.key_n),
                                        keys.begin() + static_cast<std::ptrdiff_t>(hy.key_off),
                                        keys.begin() + static_cast<std::ptrdiff_t>(hy.key_off + hy.key_n));
  });
  // leaves in that order; the interval below each germline allele
  for (uint16_t a0 = 0; a0 < 2; ++a0) w.germ_lo[a0] = w.germ_hi[a0] = 0;
  w.haps.reserve(order.size());
  for (uint32_t i = 0; i < order.size(); ++i) {
    const Hap& h = haps[order[i]];
    if (i == 0 || haps[order[i - 1]].origin != h.origin) w.germ_lo[h.origin] = i;
    w.germ_hi[h.origin] = i + 1;
    w.haps.push_back({h.cell, h.fragset, h.allele, h.kind});
  }
  // placements: the runs of consecutive haplotypes that carry a row
  std::vector<uint32_t> last(n_rows, 0xffffffffu);  // index in w.inst of the row's run that ends at the previous haplotype
  for (uint32_t i = 0; i < order.size(); ++i) {
    const Hap& h = haps[order[i]];
    for (uint32_t k = 0; k < h.key_n; ++k) {
      const uint32_t r = by_count[keys[h.key_off + k]];  // row - row_lo
      if (last[r] != 0xffffffffu && w.inst[last[r]].lo + w.inst[last[r]].span == i) {
        ++w.inst[last[r]].span;
      } else {
        const uint32_t m = w.row_lo + r;
        last[r] = static_cast<uint32_t>(w.inst.size());
        w.inst.push_back({i, 1u, m, static_cast<uint32_t>(d.mut_ref_len[m]) | (static_cast<uint32_t>(d.mut_alt_len[m]) << 8)});
      }
    }
  }
  finish_chr(w);
}

pcs_forest_desc common_desc(const pcs_cell_genomes_desc& g) {
  pcs_forest_desc d{};
  d.n_chr = g.n_chr;
  d.chr_len = g.chr_len;
  d.chr_n_alleles = g.chr_n_alleles;
  d.n_samples = g.n_samples;
  d.n_leaves = g.n_cells;
  d.leaf_sample = g.cell_sample;
  d.n_mut = g.n_mut;
  d.mut_chr = g.mut_chr;
  d.mut_pos = g.mut_pos;
  d.mut_ref_len = g.mut_ref_len;
  d.mut_alt_len = g.mut_alt_len;
  d.n_germline = g.n_germline;
  d.germ_mut = g.germ_mut;
  d.germ_allele_mask = g.germ_allele_mask;
  return d;
}

}  // namespace

void flatten_cell_genomes(const pcs_cell_genomes_desc& g, FlatForest& out, unsigned n_threads,
                          const std::function<void()>& loci_ready, bool defer_instances) {
  const pcs_forest_desc d = common_desc(g);
  for (uint32_t c = 0; c < g.n_cells; ++c) check(g.cell_sample[c] < g.n_samples, "cell_sample out of range");
  check(g.n_alleles == 0 || (g.allele_frag_off[0] == 0 && g.allele_sid_off[0] == 0), "allele offsets are not CSR arrays");
  flatten_with(d, out, n_threads, loci_ready, defer_instances, [&](Numbering& nb) {
    nb.out.n_roots = g.n_normal_preneo;
    std::vector<std::vector<uint64_t>> by_chr(g.n_chr);
    for (uint64_t a = 0; a < g.n_alleles; ++a) {
      check(g.allele_chr[a] < g.n_chr, "allele_chr out of range");
      check(g.allele_frag_off[a] <= g.allele_frag_off[a + 1] && g.allele_sid_off[a] <= g.allele_sid_off[a + 1],
            "allele offsets must be non-decreasing");
      by_chr[g.allele_chr[a]].push_back(a);
    }
    nb.parallel_for(g.n_chr, [&](uint32_t c) {
      number_from_genomes(g, by_chr[c], nb.work[c], d, nb.row_mask, nb.out.row_locus.data(), nb.out.locus_first_row.data());
    });
  });
}

size_t flat_store_bytes(const pcs_cell_genomes_desc& g) {
  auto padded = [](size_t n, size_t elem) { return (std::max<size_t>(n * elem, 1) + 255) & ~static_cast<size_t>(255); };
  const size_t n_sid = g.n_alleles ? static_cast<size_t>(g.allele_sid_off[g.n_alleles]) : 0;
  return 2 * padded(g.n_mut, 4) + 2 * padded(static_cast<size_t>(g.n_mut) + 1, 4) + padded(g.n_mut, 2) +
         padded(static_cast<size_t>(g.n_mut) + 1, 1) + padded(n_sid + static_cast<size_t>(g.n_germline), sizeof(Inst));
}

size_t flat_store_bytes(const pcs_forest_desc& d) {
  auto padded = [](size_t n, size_t elem) { return (std::max<size_t>(n * elem, 1) + 255) & ~static_cast<size_t>(255); };
  // loci <= rows; a SID event or a germline entry places at most one instance
  return 2 * padded(d.n_mut, 4) + 2 * padded(static_cast<size_t>(d.n_mut) + 1, 4) + padded(d.n_mut, 2) +
         padded(static_cast<size_t>(d.n_mut) + 1, 1) + padded(static_cast<size_t>(d.n_events + d.n_germline), sizeof(Inst));
}

}  // namespace pcs
