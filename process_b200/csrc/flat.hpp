// flat.hpp -- host-side flattened ("sequencing") view of a phylogenetic forest.
//
// The reference walks explicit per-cell genomes (chromosome -> allele -> fragment
// -> SID, src/phylogenetic_forest.cpp:279-376).  Here every (sampled cell, allele)
// pair of a chromosome is a leaf of a HAPLOTYPE TREE: the cell tree with a branch
// added wherever an amplification or WGD copies an allele.  Leaves are numbered in
// DFS order, so the carriers of a SID are one contiguous interval [lo, lo+span) of
// haplotype indices and "does this read carry the SID" is one unsigned compare.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pcs_seq.h"

namespace pcs {

// allocator that default-initialises: vector::resize() of the multi-megabyte tables below
// must not spend time (and page faults on one thread) zeroing what is overwritten right away
template <class T, class A = std::allocator<T>>
struct default_init_allocator : A {
  using A::A;
  template <class U>
  struct rebind {
    using other = default_init_allocator<U, typename std::allocator_traits<A>::template rebind_alloc<U>>;
  };
  template <class U>
  void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) {
    ::new (static_cast<void*>(p)) U;
  }
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    std::allocator_traits<A>::construct(static_cast<A&>(*this), p, std::forward<Args>(args)...);
  }
};
template <class T>
using BigVec = std::vector<T, default_init_allocator<T>>;

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
  BigVec<uint32_t> locus_pos;            // [L]
  std::vector<uint32_t> chr_locus_off;   // [n_chr+1]
  BigVec<uint32_t> locus_inst_off;       // [L+1]
  BigVec<uint32_t> row_locus;            // [n_mut]
  BigVec<uint32_t> locus_first_row;      // [L+1]
  BigVec<Inst> inst;                     // sorted by row

  // haplotype leaves, per chromosome (index inside a chromosome = haplotype index)
  std::vector<std::vector<HapRec>> chr_haps;

  // fragment sets (interned lists of fragments) and the pieces they cut chromosomes into
  std::vector<std::vector<Frag>> fragsets;
  std::vector<uint32_t> full_fragset;    // [n_chr] id of {[1, chr_len]}
  std::vector<Piece> pieces;             // sorted by (chr, begin)
  std::vector<uint32_t> chr_piece_off;   // [n_chr+1]
  std::vector<Cover> covers;
};

// throws std::domain_error on malformed input
void flatten_forest(const pcs_forest_desc& d, FlatForest& out, unsigned n_threads);

}  // namespace pcs
