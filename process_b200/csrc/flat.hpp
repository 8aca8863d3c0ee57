// flat.hpp -- host-side flattened ("sequencing") view of a phylogenetic forest.
//
// The reference walks explicit per-cell genomes (chromosome -> allele -> fragment
// -> SID, src/phylogenetic_forest.cpp:279-376).  Here every (sampled cell, allele)
// pair of a chromosome is a leaf of a HAPLOTYPE TREE: the cell tree with a branch
// added wherever an amplification or WGD copies an allele.  Leaves are numbered in
// DFS order, so the carriers of a SID are one contiguous interval [lo, lo+span) of
// haplotype indices and "does this read carry the SID" is one unsigned compare.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pcs_seq.h"

namespace pcs {

// A big table of the flattened view: a typed window into memory owned by the forest's FlatStore.
template <class T>
struct Table {
  T* p = nullptr;
  size_t n = 0;
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  T* begin() { return p; }
  T* end() { return p + n; }
  const T* begin() const { return p; }
  const T* end() const { return p + n; }
};

// Where the big tables live.  Either one block the caller lends for the forest's lifetime -- pinned host
// memory, so that the H2D DMA reads the tables where the flattener wrote them, no staging copy and no page
// faults on fresh heap pages -- or the heap (also the fallback once the block is used up).  Memory comes
// back uninitialised: every table is written in full by flatten_forest.
struct FlatStore {
  char* base = nullptr;
  size_t capacity = 0, used = 0;
  std::vector<std::unique_ptr<char[]>> heap;

  void* take(size_t bytes) {
    bytes = (std::max<size_t>(bytes, 1) + 255) & ~static_cast<size_t>(255);
    if (base && used + bytes <= capacity) {
      void* q = base + used;
      used += bytes;
      return q;
    }
    heap.emplace_back(new char[bytes]);
    return heap.back().get();
  }
  template <class T>
  Table<T> table(size_t count) {
    return Table<T>{static_cast<T*>(take(count * sizeof(T))), count};
  }
  // is [q, q + bytes) inside the lent block?
  bool lent(const void* q, size_t bytes) const {
    const char* c = static_cast<const char*>(q);
    return base && c >= base && c + bytes <= base + used;
  }
};

struct Inst {          // one placement of a SID on a haplotype subtree (device layout: uint4)
  uint32_t lo;         // first haplotype index carrying it
  uint32_t span;       // number of haplotype indices
  uint32_t row;        // mutation table row
  uint32_t meta;       // ref_len | alt_len << 8
};

enum HapKind : uint8_t { HAP_TUMOUR = 0, HAP_NORMAL_PLAIN = 1, HAP_NORMAL_PRENEO = 2 };

struct HapRec {        // haplotype leaf: index h is its position in chr_haps
  uint32_t cell;       // leaf index / 0 / root ordinal
  uint32_t fragset;    // global fragment-set id
  uint16_t allele;
  uint8_t kind;
};

struct Frag { uint32_t b, e; };                 // inclusive
struct Cover { uint32_t fragset, frag_end; };   // a fragment set covering a piece
struct Piece { uint32_t chr, begin, end, cover_off, cover_n; };

struct FlatForest {
  uint32_t n_chr = 0, n_mut = 0, n_leaves = 0, n_samples = 0, n_roots = 0;
  std::vector<uint32_t> chr_len;
  std::vector<uint8_t> chr_n_alleles;
  std::vector<uint32_t> leaf_sample;

  // loci: distinct (chr, pos) of the mutation table
  FlatStore store;                       // backing memory of the five tables below
  Table<uint32_t> locus_pos;             // [L]
  std::vector<uint32_t> chr_locus_off;   // [n_chr+1]
  Table<uint32_t> locus_inst_off;        // [L+1]
  Table<uint32_t> row_locus;             // [n_mut]
  Table<uint32_t> locus_first_row;       // [L+1]
  Table<Inst> inst;                      // sorted by row; inside a row: somatic placements in DFS order, then germline

  // Deferred instance table (uploads).  The germline's share of `inst` -- 99 % of it on a WGS forest -- is a function
  // of one allele-mask byte per row and three haplotype intervals per chromosome.  A forest flattened with
  // defer_instances keeps those instead of `inst` / `locus_inst_off` (both stay empty): the device builds the two
  // tables from them (kernels.cu: build_instances_kernel) -- 3 bytes per row to write and send instead of 20.
  bool inst_deferred = false;
  size_t n_inst = 0;                     // instances of the forest, deferred or not
  Table<uint8_t> germ_mask;              // [n_mut] germline allele mask of every row (0: not germline)
  Table<uint16_t> row_meta;              // [n_mut] ref_len | alt_len << 8
  std::vector<Inst> som;                 // somatic placements, sorted by row (inside a row: DFS order)
  std::vector<uint32_t> germ_iv;         // [n_chr][8]: first haplotype of allele mask 0..3, haplotypes of mask 0..3
  std::vector<uint32_t> chr_row_off;     // [n_chr+1] rows of every chromosome

  // haplotype leaves, per chromosome (index inside a chromosome = haplotype index)
  std::vector<std::vector<HapRec>> chr_haps;

  // fragment sets (interned lists of fragments) and the pieces they cut chromosomes into
  std::vector<std::vector<Frag>> fragsets;
  std::vector<uint32_t> full_fragset;    // [n_chr] id of {[1, chr_len]}
  std::vector<Piece> pieces;             // sorted by (chr, begin)
  std::vector<uint32_t> chr_piece_off;   // [n_chr+1]
  std::vector<Cover> covers;
};

// throws std::domain_error on malformed input.  A block lent through out.store before the call is kept and used.
// loci_ready (optional) is called, on the calling thread, once locus_pos / row_locus / chr_locus_off are final.
// defer_instances: see FlatForest::inst_deferred (ignored -- the host builds the table -- when a germline SID is
// listed more than once).
void flatten_forest(const pcs_forest_desc& d, FlatForest& out, unsigned n_threads,
                    const std::function<void()>& loci_ready = nullptr, bool defer_instances = false);
// the same view from explicit per-cell genomes (what get_sample_mutations_list() / get_normal_sample() hand over)
void flatten_cell_genomes(const pcs_cell_genomes_desc& g, FlatForest& out, unsigned n_threads,
                          const std::function<void()>& loci_ready = nullptr, bool defer_instances = false);
// bytes of lent memory that are always enough for the tables of `d` (FlatStore::capacity)
size_t flat_store_bytes(const pcs_forest_desc& d);
size_t flat_store_bytes(const pcs_cell_genomes_desc& g);

}  // namespace pcs
