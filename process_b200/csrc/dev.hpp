// dev.hpp -- POD records shared by the host planner and the CUDA kernels.
#pragma once
#include <cstdint>

#include <vector_types.h>

namespace pcs {

// One way of drawing a haplotype for a read that starts inside a tile: the
// haplotype leaves of one (sample group | normal cells, fragment set) list.
// A single 32-bit draw u picks the entry (first with u <= thr) and the leaf inside
// it: leaf = umulhi(u - base, scale), base = previous entry's thr + 1 (0 for the first),
// scale = floor(list_n * 2^32 / (thr - base + 1)).  16 bytes: one LDS.128 per read.
struct Entry {
  uint32_t thr;       // last draw value belonging to this entry (cumulative)
  uint32_t scale;
  uint32_t list_off;  // into hap_list
  uint32_t frag_end;  // last position of the fragment the tile lies in (reads never cross it)
};
static_assert(sizeof(Entry) == 16, "Entry layout");

constexpr uint32_t kMaxStagedEntries = 16;  // tiles drawing from more lists use the global kernel

// A tile: a stretch of one piece of one chromosome for one output sample.  Every
// template whose start falls in [begin, begin+len) is drawn by the CTA owning it.
struct Tile {
  uint32_t chr;
  uint32_t begin;
  uint32_t len;
  uint32_t n_templates;
  uint32_t entry_off;
  uint32_t n_entries;
  uint32_t sample;   // output sample index
  uint32_t id;       // global tile number: Philox key, independent of sharding
  uint32_t l0, l1;   // loci with position in [begin, begin+len+reach): staged range
  uint32_t r0;       // first mutation row of locus l0
  uint32_t n_rows;   // rows of the loci [l0, l1)
};
static_assert(sizeof(Tile) == 48, "Tile layout");

struct DevForest {
  const uint32_t* locus_pos;       // [L]
  const uint32_t* chr_locus_off;   // [n_chr+1]
  const uint32_t* locus_inst_off;  // [L+1]
  const uint4* inst;               // [I] {lo, span, row, ref_len | alt_len<<8}
  const uint32_t* hap_list;        // concatenated haplotype lists
  uint32_t n_loci;
  uint32_t n_mut;
};

struct SeqModel {
  uint32_t read_size;
  uint32_t paired;          // 0/1
  uint32_t sequencer;       // PCS_SEQ_*
  uint32_t err_thr;         // floor(error_rate * 2^32), constant-quality model
  float error_rate;         // random-quality model
  uint32_t insert_n;        // columns of the insert-size alias table (paired)
  uint32_t insert_min;      // smallest insert with non-zero probability
  const uint32_t* insert_alias;  // [insert_n][2] {keep threshold over the u32 range, alias column}
  uint32_t seed;
  uint32_t reach;           // bases past a tile's last start position a template can span (without deletions)
  uint32_t dir_shift;       // staged kernel: log2 of the bucket size of the in-tile locus directory
};

// shared-memory capacities of the staged sampler kernel for one plan (maxima over its staged tiles)
struct StageDims {
  uint32_t max_loci;
  uint32_t max_rows;
  uint32_t max_buckets;
};

// reference bases and alt strings, needed only when reads are materialised (SAM output)
struct SeqData {
  const uint8_t* ref;            // ASCII bases of all loaded chromosomes, concatenated
  const uint64_t* chr_ref_off;   // [n_chr+1] offset of position 1 of each chromosome; equal offsets = not loaded
  const uint8_t* alt;            // alt strings of all rows, concatenated
  const uint32_t* alt_off;       // [n_mut+1]
};

constexpr uint32_t kMaxCigar = 16;

// one materialised read (device layout, 112 bytes); bases and qualities live in separate [n][R] arrays
struct SamHeader {
  uint32_t hap, start, frag_end, chr_sample;   // as DevPlacement
  uint32_t read_id, tile_id, flags, mate_start; // flags: bit0 paired, bit1 second mate, bit2 CIGAR truncated
  int32_t tlen;
  uint32_t n_cigar, len, pad0;                  // len: bases actually written (R unless clipped by a fragment end)
  uint32_t cigar[kMaxCigar];                    // length << 4 | op (0 M, 1 I, 2 D)
};
static_assert(sizeof(SamHeader) == 112, "SamHeader layout");

// injected placement after host translation (cell, allele) -> haplotype index
struct DevPlacement {
  uint32_t hap;
  uint32_t start;
  uint32_t frag_end;
  uint32_t chr_sample;  // chr | sample << 16
};

}  // namespace pcs
