// dev.hpp -- POD records shared by the host planner and the CUDA kernels.
#pragma once
#include <cstdint>

#include <vector_types.h>

namespace pcs {

#if defined(__CUDACC__)
#define PCS_HD __host__ __device__ __forceinline__
#else
#define PCS_HD inline
#endif

// One way of drawing a haplotype for a read that starts inside a tile: the
// haplotype leaves of one (sample group | normal cells, fragment set) list.
// A single 32-bit draw u picks the entry (first with u <= thr) and the leaf inside
// it.  With x = u - base (base = previous entry's thr + 1, 0 for the first) and
// width = thr - base + 1 the leaf is EXACTLY floor(x * list_n / width): every leaf of
// an entry owns floor(width / list_n) or ceil(width / list_n) of the entry's draw
// values -- as uniform as 32 bits allow (every cell / allele of a class is
// equiprobable: src/sequencing.cpp:155-163, src/seq_simulation.cpp:572-578).
// The quotient is a 64-bit fixed-point multiply by scale64 = ceil(list_n * 2^64 / width)
// (exact_leaf_scale / exact_leaf below): `scale` holds its high word here, the low
// word lives in the parallel array entry_lo[] (one extra LDS.32 per read that reaches
// the walk; the 16-byte entry stays one LDS.128).
struct Entry {
  uint32_t thr;       // last draw value belonging to this entry (cumulative)
  uint32_t scale;     // high word of scale64
  uint32_t list_off;  // into hap_list
  uint32_t frag_end;  // last position of the fragment the tile lies in (reads never cross it)
};
static_assert(sizeof(Entry) == 16, "Entry layout");

// floor(x * list_n / width) for every 0 <= x < width <= 2^32 when list_n < width:
// x * scale64 / 2^64 = x * list_n / width + x * delta / 2^64 with 0 <= delta < 1, and the excess
// (< 2^-32) cannot reach the next integer because frac(x * list_n / width) <= 1 - 1 / width.
PCS_HD uint32_t exact_leaf(uint32_t x, uint32_t scale_hi, uint32_t scale_lo) {
#if defined(__CUDA_ARCH__)
  const uint32_t carry = __umulhi(x, scale_lo);
#else
  const uint32_t carry = static_cast<uint32_t>((static_cast<uint64_t>(x) * scale_lo) >> 32);
#endif
  return static_cast<uint32_t>((static_cast<uint64_t>(x) * scale_hi + carry) >> 32);
}

// host: scale64 = ceil(list_n * 2^64 / width).  list_n >= width (a class so light that it owns fewer
// draw values than it has haplotypes) saturates: leaf = x - 1 (0 for x = 0), inside the list.
inline void exact_leaf_scale(uint64_t list_n, uint64_t width, uint32_t& hi, uint32_t& lo) {
  if (width == 0 || list_n >= width) {
    hi = lo = 0xffffffffu;
    return;
  }
  const unsigned __int128 s = ((static_cast<unsigned __int128>(list_n) << 64) + (width - 1)) / width;
  hi = static_cast<uint32_t>(static_cast<uint64_t>(s >> 32));
  lo = static_cast<uint32_t>(static_cast<uint64_t>(s));
}

constexpr uint32_t kMaxStagedEntries = 16;  // tiles drawing from more lists use the global kernel
constexpr uint32_t kMaxThinLoci = 1024;     // a thinned tile keeps one cumulative count per locus in shared memory

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
  // Thinning (single-end reads, staged tiles).  A read changes the tables only if it spans a locus, i.e. if its
  // start offset lies in U = the union over the tile's loci p of [p - R + 1, p] (plus the tail zone: the offsets
  // from which a read may run past the end of its fragment, tail_off on).  Of the tile's n_templates templates,
  // n_useful ~ Binomial(n_templates, u_len / len) start in U (drawn by the host: a multinomial split), uniformly
  // over U; the others are never drawn by the counting kernels -- same law for the tables, a fifth of the work.
  // thin == 0: the tile is not thinned (every template is drawn, start uniform over the tile).
  uint32_t n_useful;
  uint32_t u_len;     // |U|, in offsets
  uint32_t tail_off;  // first offset of the tail zone (== len: none)
  uint32_t thin;
};
static_assert(sizeof(Tile) == 64, "Tile layout");

// The useful offsets of a thinned tile, from its sorted locus positions.  The tile is cut at `limit`, the last
// position before the tail zone: below it, locus i adds the offsets of its window [max(p_i - R + 1, begin),
// min(p_i, limit)] that no earlier locus covered -- gain_i of them, the LAST gain_i of the window; from limit + 1
// on every offset is useful (the tail zone: one more window after the last locus).  Host (planner: u_len) and
// device (staging: the cumulative gains the draws are mapped through) use this one rule.
struct UsefulScan {
  uint32_t begin, last, limit, R;  // tile begin, last position of the tile, last position before the tail zone, read size
  uint32_t covered_to;             // every position <= covered_to is in U or before the tile
  PCS_HD void init(uint32_t tile_begin, uint32_t tile_len, uint32_t tail_off, uint32_t read_size) {
    begin = tile_begin;
    last = tile_begin + tile_len - 1u;
    limit = tile_begin + tail_off - 1u;
    R = read_size;
    covered_to = tile_begin - 1u;
  }
  PCS_HD uint32_t add_window(uint32_t s, uint32_t e) {
    if (s <= covered_to) s = covered_to + 1u;
    if (s > e) return 0u;
    covered_to = e;
    return e - s + 1u;
  }
  PCS_HD uint32_t add_locus(uint32_t p) { return add_window(p + 1u > R ? p + 1u - R : 0u, p < limit ? p : limit); }
  PCS_HD uint32_t add_tail() { return add_window(limit + 1u, last); }
};

// |U| of a tile in closed form (host): the loci are sorted and every window has the same length, so what is covered
// when locus i comes is exactly what locus i - 1 reached -- gain_i depends on p_{i-1} and p_i only (the rule the
// device scans in parallel, kernels.cu: build_useful_cum).  No branch in the loop: it runs at memory speed.
inline uint32_t useful_offsets(const uint32_t* pos, uint32_t n, uint32_t tile_begin, uint32_t tile_len, uint32_t tail_off,
                               uint32_t R) {
  const uint32_t last = tile_begin + tile_len - 1u, limit = tile_begin + tail_off - 1u;
  auto gain = [=](uint32_t prev_e, uint32_t p) {
    const uint32_t e = p < limit ? p : limit;
    const uint32_t s0 = p + 1u > R ? p + 1u - R : 0u;
    const uint32_t s = s0 > prev_e + 1u ? s0 : prev_e + 1u;
    return s <= e ? e - s + 1u : 0u;
  };
  uint32_t u = n ? gain(tile_begin - 1u, pos[0]) : 0u;  // |U| <= tile_len: no overflow
  for (uint32_t i = 1; i < n; ++i) u += gain(pos[i - 1] < limit ? pos[i - 1] : limit, pos[i]);
  u += last > limit ? last - limit : 0u;  // the tail zone: every offset past `limit`
  return u;
}

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
  uint32_t err_thr;         // constant quality: floor(error_rate * 2^32); random quality: a bound of every base's error probability
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
