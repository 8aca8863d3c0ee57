// kernels.cu -- sm_100a kernels of the read sampler.
//
//   sample_tiles_staged_kernel  the hot kernel.  One CTA per tile: the tile's sorted loci
//                         are staged in shared memory as 16-byte records {position,
//                         carrier interval, SID lengths, row}.  Single-end reads, THINNED tile
//                         (dev.hpp: Tile): only the templates whose read can span a locus are
//                         drawn, start mapped through the cumulative widths of the loci's
//                         windows, and every one of them is walked at once.  Otherwise every
//                         template is drawn (Philox4x32-10, one block = two single-end reads
//                         or one paired template), probes a bucket directory, and the reads
//                         that may span a locus wait in a per-warp queue and are walked 32
//                         at a time.  The walk resolves the haplotype and counts depth /
//                         occurrences with shared-memory atomics; the error models test a
//                         carried SID against two bits of an error block drawn, converged,
//                         before the walk; one coalesced red.global per touched counter at
//                         the end.
//   sample_tiles_global_kernel  same walk straight from global memory (tiles too dense
//                         to stage) and the read-tracing debug mode.
//   count_injected_kernel the same locus walk over a caller-supplied placement list
//   finalize_kernel       coverage[s][row] = depth[s][locus(row)]
//   sum_u32_kernel        table checksums (k_bar, k_alt of the roofline byte model)
//
// The locus walk is the device restatement of what the reference does per read:
// apply the allele's SIDs to the reference stretch the read spans, count the read
// in the coverage of every position it spans and in the occurrences of every SID
// it carries (ReadSimulator<>::operator(), call sites src/seq_simulation.cpp:371,
// 413,423; outputs consumed at src/seq_simulation.cpp:92-140).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "../../include/pcs_seq.h"
#include "dev.hpp"
#include "kernels.hpp"

namespace pcs {

// ------------------------------------------------------------------- Philox
// Philox4x32-10 as a keyed bijection of its 128-bit counter.  The key is a
// compile-time constant, so the ten round keys fold into immediates; everything
// that varies -- (index, tile, purpose, seed) -- rides in the counter, and distinct
// tuples are distinct counters.
constexpr uint32_t kPhiloxKey0 = 0x5EEDC0DEu, kPhiloxKey1 = 0xCA11AB1Eu;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint2 k = make_uint2(kPhiloxKey0, kPhiloxKey1);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float u01(uint32_t w) { return (static_cast<float>(w) + 0.5f) * 2.3283064365386963e-10f; }

constexpr float kQualSigma = 0.5f;

// error probability ramp of the random-quality model: 0.5 at the first base, 1.5 at the last
__device__ __forceinline__ float ramp(uint32_t i, uint32_t R) {
  return R > 1 ? 0.5f + static_cast<float>(i) / static_cast<float>(R - 1) : 1.0f;
}

// -------------------------------------------------- sequencing errors on SID bases
// Does a sequencing error hide the read's hit-th carried SID?  It does iff one of the SID's bases in the read is
// an error.  Base b of that SID at read offset o is an error iff  t < thr  (constant quality) or
// u01(t) < min(1, error_rate * ramp(o) * exp(sigma * z - sigma^2 / 2))  (random quality), where
//   t  the TEST WORD:  b == 0 and hit < 2: word 2 * (read & 1) + hit of the ERROR BLOCK of the read's pair,
//                      Philox(read >> 1, tile, 1, seed) -- one block serves the first two SIDs of both reads
//                      of a draw block (or both mates of a template), and the samplers compute it once, converged,
//                      before the walk;  every other base: word x of the base's own block
//                      Q = Philox(read, tile, 0x40000000 | b << 20 | hit, seed);
//   z  the quality deviate (random quality only): Box-Muller of words y, z of Q.
// All kernels (staged / global samplers, trace, SAM records) draw through these functions, so the tables, the
// traces and the SAM files of one call describe the same reads with the same errors, whatever the scheduling.
constexpr uint32_t kErrFastHits = 2;
constexpr uint32_t kPurposeErrorBlock = 1u, kPurposeSidBase = 0x40000000u;
constexpr uint32_t kPurposeOutside = 2u;  // thinned tiles: the templates the samplers never draw (TileReads)

struct ErrModel {  // what the error draw needs of the sequencer model, by value (registers, also across a call)
  uint32_t sequencer, err_thr, read_size, seed;
  float error_rate;
};
__device__ __forceinline__ ErrModel err_model(const SeqModel& M) {
  return ErrModel{M.sequencer, M.err_thr, M.read_size, M.seed, M.error_rate};
}

__device__ __forceinline__ uint4 error_block(uint32_t read, uint32_t tile, uint32_t seed) {
  return philox4x32_10(make_uint4(read >> 1, tile, kPurposeErrorBlock, seed));
}

// The error block of a pair as the walk wants it: two bits per carried SID, in the order a read meets them --
// 0: no error, count it; 1: an error hides it; 2: ask the queue (ErrDefer).  Bits 0-3: the first two SIDs of the
// pair's even read, bits 4-7: of its odd read.  A test word at or above the model's threshold is "no error";
// below it, constant quality: an error (code 1); random quality: err_thr is only an upper bound of every base's
// error probability (set_model), so the exact test decides (code 2).
__device__ __forceinline__ uint32_t pair_error_codes(const uint4& e, uint32_t thr, uint32_t below_code) {
  return (e.x < thr ? below_code : 0u) | (e.y < thr ? below_code << 2 : 0u) | (e.z < thr ? below_code << 4 : 0u) |
         (e.w < thr ? below_code << 6 : 0u);
}
__device__ __forceinline__ uint32_t below_threshold_code(const SeqModel& M) {
  return M.sequencer == PCS_SEQ_BASIC_CONSTANT ? 1u : 2u;
}

// the exact test for one base (any base; recomputes the blocks it needs).  *quality: the base's error probability
__device__ __forceinline__ bool sid_base_error(const ErrModel& E, uint32_t read, uint32_t tile, uint32_t hit, uint32_t b,
                                               uint32_t o, float* quality = nullptr) {
  const bool fast = b == 0u && hit < kErrFastHits;
  uint4 q = make_uint4(0u, 0u, 0u, 0u);
  if (!fast || E.sequencer != PCS_SEQ_BASIC_CONSTANT) q = philox4x32_10(make_uint4(read, tile, kPurposeSidBase | (b << 20) | hit, E.seed));
  uint32_t t = q.x;
  if (fast) {
    const uint4 e = error_block(read, tile, E.seed);
    t = (read & 1u) ? (hit ? e.w : e.z) : (hit ? e.y : e.x);
  }
  if (E.sequencer == PCS_SEQ_BASIC_CONSTANT) {
    if (quality) *quality = E.error_rate;
    return t < E.err_thr;
  }
  const float z = sqrtf(-2.0f * __logf(u01(q.y))) * cospif(2.0f * u01(q.z));
  const float p = E.error_rate * ramp(o, E.read_size) * __expf(kQualSigma * z - 0.5f * kQualSigma * kQualSigma);
  if (quality) *quality = p;
  return u01(t) < fminf(p, 1.0f);
}

// is any of the `n` read bases of the hit-th SID, starting at read offset `off`, a sequencing error?  Exact, from
// scratch.  Kept out of line on purpose: it is the cold path of the samplers (a third SID in a read, an insertion,
// a random-quality test word below the bound) and must not cost the walk its registers.
__device__ __noinline__ bool sid_errors_slow(ErrModel E, uint32_t read, uint32_t tile, uint32_t hit, uint32_t off,
                                             uint32_t n, uint32_t* mask) {
  bool any = false;
  for (uint32_t b = 0; b < n; ++b) {
    if (sid_base_error(E, read, tile, hit, b, off + b)) {
      any = true;
      const uint32_t i = off + b;
      if (mask && i < 32u * PCS_ERRMASK_WORDS) mask[i >> 5] |= 1u << (i & 31);
      if (!mask) break;
    }
  }
  return any;
}

// Error policies of the locus walk.  count(V, row, abs_row, hit, off, n): the read carries the SID of `row`
// as its hit-th carried SID on read bases [off, off+n); add the occurrence unless a sequencing error hides it.
template <class View>
__device__ __forceinline__ void add_alt_row(const View& V, uint32_t row, bool abs_row) {
  if (abs_row) V.add_alt_abs(row); else V.add_alt(row);
}

struct ErrDraw {  // global-memory kernels and trace mode: every SID tested from scratch
  static constexpr bool kSnvFirst = false;
  const SeqModel& M;
  uint32_t read, tile;
  uint32_t* mask;  // trace mode: error bits found, else nullptr
  template <class View>
  __device__ __forceinline__ void count(const View& V, uint32_t row, bool abs_row, uint32_t hit, uint32_t off,
                                        uint32_t n) const {
    if (M.sequencer != PCS_SEQ_ERRORLESS && sid_errors_slow(err_model(M), read, tile, hit, off, n, mask)) return;
    add_alt_row(V, row, abs_row);
  }
  template <class View, class W, class OffFn>
  __device__ __forceinline__ void settle(const View& V, uint32_t row, bool abs_row, W& w, uint32_t n, OffFn off) const {
    count(V, row, abs_row, w.hit, off(), n);
  }
};

struct NoErr {
  static constexpr bool kSnvFirst = false;
  template <class View, class W, class OffFn>
  __device__ __forceinline__ void settle(const View& V, uint32_t row, bool abs_row, W&, uint32_t, OffFn) const {
    add_alt_row(V, row, abs_row);
  }
};

struct ErrMaskLookup {
  static constexpr bool kSnvFirst = false;
  const uint32_t* mask;  // nullptr: no errors
  template <class View>
  __device__ void count(const View& V, uint32_t row, bool abs_row, uint32_t, uint32_t off, uint32_t n) const {
    if (mask)
      for (uint32_t i = off; i < off + n; ++i)
        if (i < 32u * PCS_ERRMASK_WORDS && ((mask[i >> 5] >> (i & 31)) & 1u)) return;
    add_alt_row(V, row, abs_row);
  }
  template <class View, class W, class OffFn>
  __device__ __forceinline__ void settle(const View& V, uint32_t row, bool abs_row, W& w, uint32_t n, OffFn off) const {
    count(V, row, abs_row, w.hit, off(), n);
  }
};

// ---------------------------------------------------------------- locus walk
// One read of R bases from haplotype h.  (q, rem): next reference position and
// bases still to place as of the last carried indel; between indels the read
// advances one reference base per read base, so the state need not be touched
// at loci that carry nothing or an SNV.  stop = first position the read cannot reach.
struct Walk {
  uint32_t q, rem, stop, hit;
  uint32_t codes;  // staged kernel, error models (ErrDefer): two bits per carried SID still to come
  __device__ __forceinline__ void init(uint32_t x, uint32_t R, uint32_t frag_end) {
    q = x;
    rem = R;
    hit = 0;
    stop = min(x + R, frag_end + 1u);
  }
};

// a locus as the walk sees it: one SID instance inline (the common case) or a list
struct Locus {
  uint32_t lo, span, lens, row, k0, n_multi;
};

struct GlobalView {
  const uint32_t* pos;       // locus_pos
  const uint32_t* ioff;      // locus_inst_off
  const uint4* inst;
  uint32_t* depth;           // [L] of the sample (nullptr: trace mode)
  uint32_t* alt;             // [M] of the sample
  __device__ __forceinline__ uint32_t position(uint32_t i) const { return __ldg(pos + i); }
  __device__ __forceinline__ void load(uint32_t i, Locus& L) const {
    L.k0 = __ldg(ioff + i);
    const uint32_t n = __ldg(ioff + i + 1) - L.k0;
    L.span = 0;
    L.n_multi = n;
    if (n == 1) {
      const uint4 in = __ldg(inst + L.k0);
      L.lo = in.x; L.span = in.y; L.row = in.z; L.lens = in.w; L.n_multi = 0;
    }
  }
  __device__ __forceinline__ uint4 instance(uint32_t k) const { return __ldg(inst + k); }
  __device__ __forceinline__ void add_depth(uint32_t i) const { if (depth) atomicAdd(depth + i, 1u); }
  __device__ __forceinline__ void add_alt(uint32_t row) const { if (alt) atomicAdd(alt + row, 1u); }
  __device__ __forceinline__ void add_alt_abs(uint32_t row) const { add_alt(row); }
};

// staged record: x = position, y = lo, z = span, w = ref_len | alt_len << 8 | relative row << 16;
// z == 0: no inline instance, y = first instance (absolute index), w = how many
// The three arrays are addressed with 32-bit shared-window addresses computed once
// per CTA (ld.shared / red.shared on a register address).
struct SharedView {
  uint32_t rec;          // shared address of uint4 [n]
  uint32_t depth;        // shared address of uint32 [n]
  uint32_t alt;          // shared address of uint32 [rows]
  const uint4* inst;     // global instances, for the rare multi-instance loci
  uint32_t r0;
  __device__ __forceinline__ uint4 record(uint32_t i) const {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(rec + i * 16u));
    return r;
  }
  __device__ __forceinline__ void add_depth(uint32_t i) const {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(depth + i * 4u) : "memory");
  }
  __device__ __forceinline__ void add_alt(uint32_t row) const {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(alt + row * 4u) : "memory");
  }
  __device__ __forceinline__ void add_alt_abs(uint32_t row) const { add_alt(row - r0); }
  __device__ __forceinline__ uint4 instance(uint32_t k) const { return __ldg(inst + k); }
};

// Staged kernel, error models.  The first two SIDs of a read are settled inside the walk from the two bits of its
// pair's error block (pair_error_codes) when they are SNVs -- nine carried SIDs in ten.  Whatever needs more draws (a
// third SID, an insertion's further bases, a random-quality test word below the bound) would run inside the walk
// with the one or two lanes concerned while the others wait, so the walk counts that occurrence at once and
// queues the carried SID
// {read, row, hit, read offset | base << 16 | bases left << 24}; the warp settles 32 queued bases at a time,
// one per lane (settle_carried): a lane draws for the base of its item, puts the item back for the next base
// if there is one, and the FIRST erroneous base of a SID takes the occurrence back.  The draws are
// sid_base_error's: an occurrence survives iff none of its bases is an error, whatever the order.
constexpr uint32_t kCarriedSlots = 96;  // per warp; a SID that finds the queue full is settled on the spot

// out of line: the queue-full path inside the walk is cold and must not cost the walk registers
__device__ __noinline__ void settle_sid_cold(ErrModel E, uint32_t alt_addr, uint32_t read, uint32_t tile, uint32_t row,
                                             uint32_t hit, uint32_t off, uint32_t n) {
  for (uint32_t b = 0; b < n; ++b)
    if (sid_base_error(E, read, tile, hit, b, off + b)) return;
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(alt_addr + row * 4u) : "memory");
}

// Walk::codes: the read's four bits of pair_error_codes, then "ask the queue" for every later SID; a SID that is
// not one read base long (an insertion) asks the queue whatever its code.
struct ErrDefer {
  static constexpr bool kSnvFirst = true;
  const SeqModel& M;
  uint32_t read, tile;
  uint32_t slots0;   // shared address of warp 0's kCarriedSlots uint4 slots (warp w: + w * kCarriedSlots * 16)
  uint32_t counts0;  // shared address of warp 0's number of waiting items (warp w: + 4 w)
  template <class View, class OffFn>
  __device__ __forceinline__ void settle(const View& V, uint32_t row, bool abs_row, Walk& w, uint32_t n, OffFn off) const {
    ErrDraw{M, read, tile, nullptr}.count(V, row, abs_row, w.hit, off(), n);  // past the staged loci: from scratch
  }
  template <class OffFn>
  __device__ __forceinline__ void settle(const SharedView& V, uint32_t row, bool abs_row, Walk& w, uint32_t n, OffFn off) const {
    const uint32_t rel = abs_row ? row - V.r0 : row;
    uint32_t c = w.codes & 3u;
    w.codes = (w.codes >> 2) | 0x80000000u;
    if (n != 1u) c = 2u;
    if (c == 0u) {  // nine carried SIDs in ten leave here
      V.add_alt(rel);
      return;
    }
    if (c == 1u) return;
    defer(V, rel, w.hit, off(), n);
  }
  __device__ __forceinline__ void defer(const SharedView& V, uint32_t rel, uint32_t hit, uint32_t off, uint32_t n) const {
    if (n == 0u) {  // nothing of the SID is read: no base to get wrong (cannot happen: alt_len >= 1 and rem_p >= 1)
      V.add_alt(rel);
      return;
    }
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t slots = slots0 + warp * (kCarriedSlots * 16u), count_addr = counts0 + warp * 4u;
    // one slot per carried SID, allocated for all the lanes that are here together with ONE shared atomic
    const uint32_t together = __activemask();
    uint32_t below;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(below));
    const uint32_t rank = __popc(together & below);
    uint32_t slot = 0;
    if (rank == 0)
      asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(slot) : "r"(count_addr), "r"(__popc(together)) : "memory");
    slot = __shfl_sync(together, slot, __ffs(together) - 1) + rank;
    if (slot < kCarriedSlots) {
      V.add_alt(rel);
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slots + slot * 16u), "r"(read), "r"(rel),
                   "r"(hit), "r"(off | ((n - 1u) << 24))
                   : "memory");
    } else {
      settle_sid_cold(err_model(M), V.alt, read, tile, rel, hit, off, n);
    }
  }
};

// a carried SID at position p: count it unless a sequencing error hides it, then let an
// indel move the read's frame.  Returns false when the read is used up.
template <class View, class Err>
__device__ __forceinline__ bool carried_sid(const View& V, uint32_t p, uint32_t lens, uint32_t row, bool abs_row,
                                            uint32_t R, uint32_t frag_end, Walk& w, const Err& err) {
  if (Err::kSnvFirst) {
    // the staged error-model walk: SNVs (nine carried SIDs in ten) leave through the shortest path
    if ((lens & 0xffffu) == 0x0101u) {  // one read base (the read reaches p, so it has one left), the frame stays
      err.settle(V, row, abs_row, w, 1u, [&] { return R - (w.rem - (p - w.q)); });
      ++w.hit;
      return true;
    }
    const uint32_t rem_p = w.rem - (p - w.q);  // bases left when the read reaches p (>= 1)
    const uint32_t consumed = min((lens >> 8) & 0xffu, rem_p);
    err.settle(V, row, abs_row, w, consumed, [&] { return R - rem_p; });
    ++w.hit;
    w.rem = rem_p - consumed;
    w.q = p + (lens & 0xffu);
    w.stop = min(w.q + w.rem, frag_end + 1u);
    return w.rem != 0;
  }
  // One settle() for SNVs and indels alike: lanes that carry an SNV and lanes that carry an indel at the same
  // trip of the walk run it together (errorless: it is the one red.shared of the occurrence).
  const uint32_t ref_len = lens & 0xffu, alt_len = (lens >> 8) & 0xffu;
  const uint32_t rem_p = w.rem - (p - w.q);  // bases left when the read reaches p (>= 1)
  const uint32_t consumed = min(alt_len, rem_p);
  err.settle(V, row, abs_row, w, consumed, [&] { return R - rem_p; });
  ++w.hit;
  if ((lens & 0xffffu) != 0x0101u) {  // an indel moves the read's frame
    w.rem = rem_p - consumed;
    w.q = p + ref_len;
    w.stop = min(w.q + w.rem, frag_end + 1u);
    return w.rem != 0;
  }
  return true;
}

// Walk loci [i, end) of the global arrays.  Returns false if the view ran out
// before the read did.
template <class Err>
__device__ __forceinline__ bool walk_global(const GlobalView& V, uint32_t i, uint32_t end, uint32_t h, uint32_t R,
                                            uint32_t frag_end, Walk& w, const Err& err) {
  for (; i < end; ++i) {
    const uint32_t p = V.position(i);
    if (p >= w.stop) return true;
    if (p < w.q) continue;  // inside the reference bases a carried SID replaced
    V.add_depth(i);
    Locus L;
    V.load(i, L);
    if (L.span != 0) {
      if (h - L.lo < L.span && !carried_sid(V, p, L.lens, L.row, false, R, frag_end, w, err)) return true;
    } else {
      for (uint32_t k = L.k0; k < L.k0 + L.n_multi; ++k) {
        const uint4 in = V.instance(k);
        if (h - in.x < in.y) {  // a haplotype carries at most one SID per position (enforced by the flattener)
          if (!carried_sid(V, p, in.w, in.z, false, R, frag_end, w, err)) return true;
          break;
        }
      }
    }
  }
  return false;
}

template <class Err>
__device__ __forceinline__ bool walk_shared(const SharedView& V, uint32_t i, uint32_t end, uint32_t h, uint32_t R,
                                            uint32_t frag_end, Walk& w, const Err& err) {
  for (;; ++i) {
    const uint4 r = V.record(i);  // record `end` is a sentinel at position 0xffffffff
    if (r.x >= w.stop) return i < end;
    if (r.x < w.q) continue;
    V.add_depth(i);
    if (r.z != 0) {
      if (h - r.y < r.z && !carried_sid(V, r.x, r.w & 0xffffu, r.w >> 16, false, R, frag_end, w, err)) return true;
    } else {
      for (uint32_t k = r.y; k < r.y + r.w; ++k) {
        const uint4 in = V.instance(k);
        if (h - in.x < in.y) {
          if (!carried_sid(V, r.x, in.w, in.z, true, R, frag_end, w, err)) return true;
          break;
        }
      }
    }
  }
}

__device__ __forceinline__ uint32_t lower_bound_pos(const uint32_t* pos, uint32_t lo, uint32_t hi, uint32_t x) {
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(pos + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------ template draws
// Single-end: Philox block j of a tile yields templates 2j and 2j+1 (two words
// each: start, haplotype).  Paired-end: block j yields template j (start,
// haplotype, insert).
struct Template {
  uint32_t x, h, frag_end;
};

template <class EntryPtr>
__device__ __forceinline__ bool place_at(const Tile& T, EntryPtr ent, const uint32_t* ent_lo, const DevForest& F,
                                         uint32_t off, uint32_t u_hap, uint32_t tlen, Template& out);

template <class EntryPtr>
__device__ __forceinline__ bool place(const Tile& T, EntryPtr ent, const uint32_t* ent_lo, const DevForest& F,
                                      uint32_t u_start, uint32_t u_hap, uint32_t tlen, Template& out) {
  return place_at(T, ent, ent_lo, F, __umulhi(u_start, T.len), u_hap, tlen, out);
}

// the template that starts at tile offset `off`: haplotype from its draw word, fragment end, does it fit
template <class EntryPtr>
__device__ __forceinline__ bool place_at(const Tile& T, EntryPtr ent, const uint32_t* ent_lo, const DevForest& F,
                                         uint32_t off, uint32_t u_hap, uint32_t tlen, Template& out) {
  out.x = T.begin + off;
  uint32_t e = 0, base = 0;
  while (u_hap > ent[e].thr) {  // the last entry's thr is 0xffffffff
    base = ent[e].thr + 1u;
    ++e;
  }
  const uint32_t leaf = exact_leaf(u_hap - base, ent[e].scale, __ldg(ent_lo + e));
  out.h = __ldg(F.hap_list + (ent[e].list_off + leaf));
  out.frag_end = ent[e].frag_end;
  return out.x + (tlen - 1u) <= out.frag_end;  // else the template falls off its molecule
}

// insert size ~ Binomial(t, p) (get_bin_dist, src/seq_simulation.cpp:431-451) by Walker's alias method:
// the high part of u * n picks a column, the low part is the coin -- one 8-byte load, no search
__device__ __forceinline__ uint32_t draw_insert(const SeqModel& M, uint32_t u) {
  const uint64_t prod = static_cast<uint64_t>(u) * M.insert_n;
  const uint32_t col = static_cast<uint32_t>(prod >> 32), coin = static_cast<uint32_t>(prod);
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(M.insert_alias) + col);  // {keep threshold, alias}
  return M.insert_min + (coin < a.x ? col : a.y);
}

__device__ __forceinline__ void block_add_u64(uint32_t v, unsigned long long* dst) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ uint32_t s_part[32];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
    for (uint32_t i = 0; i < (blockDim.x + 31) / 32; ++i) tot += s_part[i];
    if (tot) atomicAdd(dst, tot);
  }
}

// *dst += total - (sum of `minus` over the block)
__device__ __forceinline__ void block_sub_u64(unsigned long long total, uint32_t minus, unsigned long long* dst) {
  for (int o = 16; o > 0; o >>= 1) minus += __shfl_xor_sync(0xffffffffu, minus, o);
  __shared__ uint32_t s_minus[32];
  if ((threadIdx.x & 31) == 0) s_minus[threadIdx.x >> 5] = minus;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long m = 0;
    for (uint32_t i = 0; i < (blockDim.x + 31) / 32; ++i) m += s_minus[i];
    if (total > m) atomicAdd(dst, total - m);
  }
}

// ------------------------------------------------------------ thinned tiles
// The useful offsets of a thinned tile (dev.hpp: Tile, UsefulScan), as the device sees them: cum[i] = offsets of
// U contributed by the loci 0..i; cum[n] adds the tail zone, so cum[n] == T.u_len.  A gain depends on nothing but
// the previous locus' position, so the gains are computed in parallel and scanned.
// pos(i): position of the tile's i-th locus.  Block-wide (any block size that is a multiple of 32, <= 1024);
// ends with __syncthreads().
template <class PosFn>
__device__ void build_useful_cum(const Tile& T, uint32_t R, uint32_t n_u, PosFn pos, uint32_t* cum) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const uint32_t begin = T.begin, last = T.begin + T.len - 1u, limit = T.begin + T.tail_off - 1u;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t i0 = 0; i0 <= n_u; i0 += blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    uint32_t g = 0;
    if (i <= n_u) {
      const uint32_t prev_e = i ? min(pos(i - 1u), limit) : begin - 1u;
      uint32_t s_, e_;
      if (i < n_u) {
        const uint32_t p = pos(i);
        s_ = p + 1u > R ? p + 1u - R : 0u;
        e_ = min(p, limit);
      } else {
        s_ = limit + 1u;
        e_ = last;
      }
      s_ = max(s_, prev_e + 1u);
      g = s_ <= e_ ? e_ - s_ + 1u : 0u;
    }
    uint32_t x = g;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= static_cast<uint32_t>(o)) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = lane < n_warps ? s_w[lane] : 0u;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= static_cast<uint32_t>(o)) w += y;
      }
      s_w[lane] = w;
    }
    __syncthreads();
    const uint32_t incl = s_carry + (warp ? s_w[warp - 1] : 0u) + x;
    if (i <= n_u) cum[i] = incl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1u) s_carry = incl;
    __syncthreads();
  }
}

// the i-th window's end, and the offset a useful-draw t (< T.u_len) maps to once i = the window it falls in
// (smallest i with cum[i] > t): the new offsets of window i are its LAST cum[i] - cum[i-1] positions
__device__ __forceinline__ uint32_t useful_position(uint32_t t, uint32_t cum_i, uint32_t window_end) {
  return window_end - (cum_i - 1u - t);
}

// Enumeration of a tile's reads for the kernels that need ALL of them (trace, SAM records, coverage tracks), from
// global memory.  Thinned tile: read r < n_useful is the sampler's read r -- Philox block r / 2 of purpose 0, start
// drawn over U; read r >= n_useful is one of the templates the sampler never draws: block (r - n_useful) / 2 of
// purpose 2, start drawn over the offsets NOT in U (it spans no locus and cannot fall off its fragment).  A tile
// that is not thinned: block r / 2 of purpose 0, start uniform over the tile.
struct TileReads {
  const uint32_t* pos;  // locus positions of the tile (global)
  const uint32_t* cum;  // shared: cumulative useful offsets, n_u + 1 entries (thinned tiles)
  uint32_t n_u, begin, last, limit, len, u_len, n_useful, thin, tile_id, seed;
  template <class CumBuf>
  __device__ void init(const Tile& T, const DevForest& F, const SeqModel& M, CumBuf* cum_buf) {
    pos = F.locus_pos + T.l0;
    cum = cum_buf;
    begin = T.begin;
    len = T.len;
    last = T.begin + T.len - 1u;
    limit = T.begin + T.tail_off - 1u;
    u_len = T.u_len;
    n_useful = T.n_useful;
    thin = T.thin && !M.paired;
    tile_id = T.id;
    seed = M.seed;
    n_u = 0;
    if (thin) {  // block-wide
      n_u = T.l1 - T.l0;
      const uint32_t* p = pos;
      build_useful_cum(T, M.read_size, n_u, [p](uint32_t i) { return __ldg(p + i); }, cum_buf);
    }
  }
  __device__ __forceinline__ uint32_t window_end(uint32_t i) const { return i < n_u ? min(__ldg(pos + i), limit) : last; }
  // single-end read r of the tile: start offset, haplotype draw; first: index of the first locus the read spans
  // (n_u: none)
  __device__ void read(uint32_t r, uint32_t& off, uint32_t& u_hap, uint32_t& first) const {
    if (!thin) {
      const uint4 u = philox4x32_10(make_uint4(r >> 1, tile_id, 0u, seed));
      off = __umulhi((r & 1u) ? u.z : u.x, len);
      u_hap = (r & 1u) ? u.w : u.y;
      first = 0xffffffffu;  // unknown: search
      return;
    }
    const bool useful = r < n_useful;
    const uint32_t q = useful ? r : r - n_useful;
    const uint4 u = philox4x32_10(make_uint4(q >> 1, tile_id, useful ? 0u : kPurposeOutside, seed));
    const uint32_t us = (q & 1u) ? u.z : u.x;
    u_hap = (q & 1u) ? u.w : u.y;
    if (useful) {
      const uint32_t t = __umulhi(us, u_len);
      uint32_t lo = 0, hi = n_u;  // smallest i with cum[i] > t
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cum[mid] > t) hi = mid; else lo = mid + 1u;
      }
      off = useful_position(t, cum[lo], window_end(lo)) - begin;
      first = lo;
    } else {
      // the t-th offset outside U: outside[i] = offsets of the tile up to window i's end that are not in U
      const uint32_t t = __umulhi(us, len - u_len);
      uint32_t lo = 0, hi = n_u;  // smallest i with outside[i] > t
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (window_end(mid) - begin + 1u - cum[mid] > t) hi = mid; else lo = mid + 1u;
      }
      const uint32_t prev_e = lo ? window_end(lo - 1u) : begin - 1u;
      const uint32_t prev_out = lo ? prev_e - begin + 1u - cum[lo - 1u] : 0u;
      off = prev_e + 1u + (t - prev_out) - begin;
      first = n_u;
    }
  }
};

// ---------------------------------------------------- staged sampler kernel
// Error models: the two reads of a draw block are walked by ONE copy of the walk (a rolled loop): unrolled, the
// second copy costs the kernel ~60 bytes of spills in the loop (constant quality on C3: 19.5 ms unrolled, 15.7 ms
// rolled -- profiles/r02_v5_thin_loop_variants.md).  The errorless kernel has the registers and stays unrolled.  Keeping the loop's invariants in shared
// memory instead of registers (PCS_THIN_STATE_IN_SMEM=1) was measured too and does not pay.
#ifndef PCS_THIN_UNROLL
#define PCS_THIN_UNROLL 1
#endif
#ifndef PCS_THIN_STATE_IN_SMEM
#define PCS_THIN_STATE_IN_SMEM 0
#endif
constexpr int kThinUnroll = PCS_THIN_UNROLL;
constexpr int kStagedThreads = 256;
constexpr int kDefaultMinCtas = 3;
constexpr uint32_t kQueueSlots = 96;  // per warp: < 32 waiting + <= 64 pushed by one Philox block per lane

// hide where a shared-window address came from, so the compiler keeps it in a register
// instead of rebuilding it from SR_CgaCtaId inside every loop
__device__ __forceinline__ uint32_t opaque(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
  return r;
}

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct StagedTile {
  SharedView SV;
  uint32_t dir;    // shared address of uint2 [buckets]: {index, position} of the first locus at/after the bucket
  uint32_t ent;    // shared address of the tile's entries (16 bytes each)
  uint32_t carried, carried_n;  // error models: shared addresses of warp 0's carried-SID queue and its count
  uint32_t n, stage_end, chr_l1;
  __device__ __forceinline__ uint2 first_locus(uint32_t bucket) const {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(dir + bucket * 8u));
    return v;
  }
  // the sampling entry a haplotype draw falls in: {thr, scale high word, list_off, frag_end}; base = first draw
  // value of the entry, lo_addr = shared address of the low word of its scale
  __device__ __forceinline__ uint4 entry_of(uint32_t u_hap, uint32_t& base, uint32_t& lo_addr) const {
    uint32_t addr = ent;
    base = 0;
    uint4 a = lds128(addr);
    while (u_hap > a.x) {  // the last entry's thr is 0xffffffff
      base = a.x + 1u;
      addr += 16u;
      a = lds128(addr);
    }
    lo_addr = ent + kMaxStagedEntries * 16u + ((addr - ent) >> 2);
    return a;
  }
};

// one queued read {start offset in the tile, haplotype draw, read id, first staged locus to look at} through the
// staged loci (and past them, if a carried deletion stretches it that far).  The haplotype is resolved here, not
// where the read was drawn: three reads out of four never get this far and never need it.
// (Measured alternative, round 2: one locus per visit to the queue, a read with a further locus going back to
// it, so that every visit runs with a full warp -- 21.9 ms against 19.5 ms for this loop on C3: the pop, the
// push and their ballots cost more than the idle lanes of a loop that makes 2.4 trips per drain.)
template <bool ERRORS>
__device__ __forceinline__ void staged_read(const StagedTile& S, const Tile& T, const DevForest& F, const SeqModel& M,
                                            uint32_t* depth, uint32_t* alt, uint4 item, uint32_t err_codes) {
  const uint32_t R = M.read_size;
  uint32_t base, lo_addr;
  const uint4 a = S.entry_of(item.y, base, lo_addr);
  const uint32_t h = __ldg(F.hap_list + (a.z + exact_leaf(item.y - base, a.y, lds32(lo_addr))));
  const uint32_t xs = T.begin + item.x, read_id = item.z, i = item.w, frag_end = a.w;
  Walk w;
  w.init(xs, R, frag_end);
  bool done;
  if (ERRORS) {
    const ErrDefer err{M, read_id, T.id, S.carried, S.carried_n};
    w.codes = 0xAAAAAAA0u | err_codes;
    done = walk_shared(S.SV, i, S.n, h, R, frag_end, w, err);
    if (!done && w.stop > S.stage_end && T.l1 < S.chr_l1) {
      const GlobalView GV{F.locus_pos, F.locus_inst_off, F.inst, depth + static_cast<size_t>(T.sample) * F.n_loci,
                          alt + static_cast<size_t>(T.sample) * F.n_mut};
      walk_global(GV, T.l1, S.chr_l1, h, R, frag_end, w, err);
    }
  } else {
    const NoErr err;
    done = walk_shared(S.SV, i, S.n, h, R, frag_end, w, err);
    if (!done && w.stop > S.stage_end && T.l1 < S.chr_l1) {
      const GlobalView GV{F.locus_pos, F.locus_inst_off, F.inst, depth + static_cast<size_t>(T.sample) * F.n_loci,
                          alt + static_cast<size_t>(T.sample) * F.n_mut};
      walk_global(GV, T.l1, S.chr_l1, h, R, frag_end, w, err);
    }
  }
}

// a queued read: its pair's error block is drawn here, with the whole warp (the reads of a drain come from
// anywhere in the tile)
template <bool ERRORS>
__device__ __forceinline__ void staged_read_queued(const StagedTile& S, const Tile& T, const DevForest& F, const SeqModel& M,
                                                   uint32_t* depth, uint32_t* alt, uint4 item) {
  uint32_t codes = 0;
  if (ERRORS)
    codes = (pair_error_codes(error_block(item.z, T.id, M.seed), M.err_thr, below_threshold_code(M)) >> (4u * (item.z & 1u))) & 15u;
  staged_read<ERRORS>(S, T, F, M, depth, alt, item, codes);
}

// Per-warp queue of reads that may span a locus.  Drawing and probing stay converged
// (every lane works on its own reads); the divergent part -- the walk, which three reads
// out of four never need -- runs only when 32 reads are waiting, one per lane.
struct HitQueue {
  uint32_t base;  // shared address of this warp's kQueueSlots uint4 slots
  uint32_t tail;  // shared address one past the last waiting read (warp-uniform)
  __device__ __forceinline__ void push(bool has, uint4 item, uint32_t lanes_below) {
    const uint32_t mask = __ballot_sync(0xffffffffu, has);
    if (has) sts128(tail + __popc(mask & lanes_below) * 16u, item);
    tail += __popc(mask) * 16u;
  }
};

// Error models: settle the SID bases the walks of this warp queued, 32 at a time (all of them when `all`).
// Called converged.  Out of line, so that its registers (a Philox block and the quality model's float math)
// are not the sampling loop's.
__device__ __noinline__ void settle_carried(ErrModel E, uint32_t slots, uint32_t count_addr, uint32_t alt_addr,
                                            uint32_t tile, bool all) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t lanes_below;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanes_below));
  __syncwarp();
  uint32_t waiting = min(lds32(count_addr), kCarriedSlots);
  const uint32_t before = waiting;
  while (waiting >= 32u || (all && waiting != 0u)) {
    const uint32_t take = min(waiting, 32u);
    waiting -= take;
    const uint32_t first = slots + waiting * 16u;
    const bool have = lane < take;
    uint4 c = make_uint4(0u, 0u, 0u, 0u);  // {read, row, hit, read offset | base << 16 | bases left << 24}
    if (have) c = lds128(first + lane * 16u);
    const uint32_t o = c.w & 0xffffu, b = (c.w >> 16) & 0xffu;
    if (have && sid_base_error(E, c.x, tile, c.z, b, o)) {
      bool first_error = true;  // rare: was an earlier base of this SID already an error?
      for (uint32_t e = 0; e < b && first_error; ++e) first_error = !sid_base_error(E, c.x, tile, c.z, e, o - b + e);
      if (first_error) asm volatile("red.shared.add.u32 [%0], 0xffffffff;" ::"r"(alt_addr + c.y * 4u) : "memory");
    }
    // insertions: the item goes back for its next base, into a slot another lane has just read: the barrier orders
    // every lane's read before any lane's write (a vote alone synchronises the lanes, not their memory accesses)
    const bool more = (c.w >> 24) != 0u;
    const uint32_t mask = __ballot_sync(0xffffffffu, more);
    __syncwarp();
    if (more) sts128(first + __popc(mask & lanes_below) * 16u, make_uint4(c.x, c.y, c.z, c.w + 0x00010001u - 0x01000000u));
    waiting += __popc(mask);
    __syncwarp();
  }
  if (lane == 0 && waiting != before) sts32(count_addr, waiting);
  __syncwarp();
}

template <bool PAIRED, bool ERRORS, int MIN_CTAS>
__global__ void __launch_bounds__(kStagedThreads, MIN_CTAS)
sample_tiles_staged_kernel(const Tile* __restrict__ tiles, const Entry* __restrict__ entries,
                           const uint32_t* __restrict__ entry_lo, DevForest F, SeqModel M, StageDims D,
                           uint32_t* __restrict__ depth, uint32_t* __restrict__ alt,
                           unsigned long long* __restrict__ n_reads) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint4* s_rec = reinterpret_cast<uint4*>(smem);
  uint4* s_queue = s_rec + D.max_loci + 1;
  uint4* s_carried = s_queue + (kStagedThreads / 32) * kQueueSlots;
  uint4* s_ent = s_carried + (ERRORS ? (kStagedThreads / 32) * kCarriedSlots : 0u);
  uint32_t* s_ent_lo = reinterpret_cast<uint32_t*>(s_ent + kMaxStagedEntries);  // [kMaxStagedEntries]
  uint2* s_dir = reinterpret_cast<uint2*>(s_ent_lo + kMaxStagedEntries);
  uint32_t* s_depth = reinterpret_cast<uint32_t*>(s_dir + D.max_buckets);
  uint32_t* s_alt = s_depth + D.max_loci;
  __shared__ uint32_t s_safe;
  __shared__ uint32_t s_dropped;  // thinned tiles: templates that fell off their molecule
  if (threadIdx.x == 0) s_dropped = 0;
  __shared__ uint32_t s_carried_n[kStagedThreads / 32];
  if (threadIdx.x < kStagedThreads / 32) s_carried_n[threadIdx.x] = 0;

  const Tile T = tiles[blockIdx.x];
  const uint32_t n = T.l1 - T.l0;
  const uint32_t shift = M.dir_shift;
  const uint32_t n_buckets = ((T.len + M.reach) >> shift) + 1;

  // ---- stage the tile: locus records, zeroed counters, entries
  for (uint32_t i = threadIdx.x; i < n; i += kStagedThreads) {
    const uint32_t l = T.l0 + i;
    const uint32_t k0 = __ldg(F.locus_inst_off + l), k1 = __ldg(F.locus_inst_off + l + 1);
    uint4 r = make_uint4(__ldg(F.locus_pos + l), k0, 0u, k1 - k0);
    if (k1 - k0 == 1u) {
      const uint4 in = __ldg(F.inst + k0);
      r.y = in.x;
      r.z = in.y;
      r.w = (in.w & 0xffffu) | ((in.z - T.r0) << 16);
    }
    s_rec[i] = r;
    s_depth[i] = 0;
  }
  if (threadIdx.x == 0) s_rec[n] = make_uint4(0xffffffffu, 0u, 0u, 0u);  // sentinel: past every read
  for (uint32_t r = threadIdx.x; r < T.n_rows; r += kStagedThreads) s_alt[r] = 0;
  if (threadIdx.x < T.n_entries) {
    s_ent[threadIdx.x] = __ldg(reinterpret_cast<const uint4*>(entries + T.entry_off) + threadIdx.x);
    s_ent_lo[threadIdx.x] = __ldg(entry_lo + T.entry_off + threadIdx.x);
  }
  if (threadIdx.x == 32) {
    // start offsets below s_safe fit their molecule whatever haplotype is drawn: the longest template
    // (M.reach bases) ends at or before the nearest fragment end of any entry
    uint32_t min_fe = 0xffffffffu;
    for (uint32_t e = 0; e < T.n_entries; ++e) min_fe = min(min_fe, __ldg(&entries[T.entry_off + e].frag_end));
    const long long lim = static_cast<long long>(min_fe) + 2 - static_cast<long long>(M.reach) - T.begin;
    s_safe = lim <= 0 ? 0u : (lim >= static_cast<long long>(T.len) ? T.len : static_cast<uint32_t>(lim));
  }
  __syncthreads();
  const bool thin = !PAIRED && T.thin != 0u;  // the same for the whole CTA
  uint32_t* s_cum = reinterpret_cast<uint32_t*>(s_queue);    // thinned tile: cumulative useful offsets [n_u + 1]
  uint16_t* s_slot = reinterpret_cast<uint16_t*>(s_dir);     //               window of every 64th useful offset
  uint32_t n_u = 0;
  if (!thin) {
    // directory: first staged locus at or after the start of each bucket
    for (uint32_t b = threadIdx.x; b < n_buckets; b += kStagedThreads) {
      const uint32_t x = T.begin + (b << shift);
      uint32_t lo = 0, hi = n;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (s_rec[mid].x < x) lo = mid + 1; else hi = mid;
      }
      s_dir[b] = make_uint2(lo, s_rec[lo].x - T.begin);  // s_rec[n] is the sentinel at 0xffffffff; offsets from T.begin
    }
  } else {
    // the useful offsets (dev.hpp: Tile): cumulative gains of the loci's windows, and for every 64th useful offset
    // the window it falls in -- a draw finds its window with one load and, one time in three, one step
    n_u = n;
    build_useful_cum(T, M.read_size, n_u, [s_rec](uint32_t i) { return s_rec[i].x; }, s_cum);
    const uint32_t n_slots = (T.u_len >> 6) + 1u;
    for (uint32_t b = threadIdx.x; b < n_slots; b += kStagedThreads) {
      const uint32_t t = b << 6;
      uint32_t lo = 0, hi = n_u;  // smallest i with cum[i] > t (cum[n_u] = u_len > t for every slot but a last empty one)
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_cum[mid] > t) hi = mid; else lo = mid + 1u;
      }
      s_slot[b] = static_cast<uint16_t>(lo);
    }
    // the planner drew n_useful with ITS count of the useful offsets: the two must be one number
    if (threadIdx.x == 0 && s_cum[n_u] != T.u_len) asm volatile("trap;");
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  StagedTile S;
  S.SV = SharedView{opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_rec))),
                    opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_depth))),
                    opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_alt))), F.inst, T.r0};
  S.dir = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_dir)));
  S.ent = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_ent)));
  S.carried = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_carried)));
  S.carried_n = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_carried_n)));
  S.n = n;
  S.stage_end = T.begin + T.len + M.reach;  // first position whose loci are not staged
  S.chr_l1 = __ldg(F.chr_locus_off + T.chr + 1);
  HitQueue Q;
  Q.base = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_queue + warp * kQueueSlots)));
  Q.tail = Q.base;
  const uint32_t R = M.read_size;
  const uint32_t safe = s_safe, len = T.len;
  uint32_t dropped = 0;  // templates that fell off their molecule

  uint32_t lanes_below;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanes_below));
  // does a template of tlen bases starting at offset off fit the fragment its haplotype draw selects?
  auto fits = [&](uint32_t u_hap, uint32_t off, uint32_t tlen) {
    uint32_t base, lo_addr;
    return T.begin + off + (tlen - 1u) <= S.entry_of(u_hap, base, lo_addr).w;
  };
  // probe: the directory gives index and offset of the first staged locus at or after the read's bucket in one
  // load; the read goes to the queue if that locus lies before its end (the sentinel covers empty buckets)
  auto flush_carried = [&](bool all) {
    if (!ERRORS) return;
    __syncwarp();  // the walks' pushes (shared atomics of other lanes) are visible
    const uint32_t waiting = lds32(S.carried_n + warp * 4u);
    if (waiting >= 32u || (all && waiting != 0u))
      settle_carried(err_model(M), S.carried + warp * (kCarriedSlots * 16u), S.carried_n + warp * 4u, S.SV.alt, T.id, all);
  };
  auto drain = [&]() {
    while (Q.tail >= Q.base + 32u * 16u) {
      __syncwarp();
      Q.tail -= 32u * 16u;
      const uint4 mine = lds128(Q.tail + lane * 16u);
      __syncwarp();
      staged_read_queued<ERRORS>(S, T, F, M, depth, alt, mine);
      flush_carried(false);
    }
  };

  if (thin) {
    // Thinned tile: only the templates whose read can span a locus are drawn -- T.n_useful of them, start uniform
    // over the useful offsets.  Every one of them is walked (no probe, no queue): the draw lands in the new part
    // of window i, so locus i is the first the read spans; a draw in the tail zone (the last read length of a
    // fragment) is tested for falling off its fragment and looks its first locus up.
    const uint32_t cum_a = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_cum)));
    const uint32_t slot_a = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_slot)));
#if PCS_THIN_STATE_IN_SMEM
    // the loop's invariants stay in shared memory and are read where they are used: with the error models the
    // walk needs the registers (a spilled invariant would be read back from local memory instead)
    __shared__ uint32_t s_thin[4];
    if (threadIdx.x == 0) {
      s_thin[0] = T.u_len;
      s_thin[1] = T.begin + T.len - 1u;
      s_thin[2] = T.begin + T.tail_off - 1u;
      s_thin[3] = s_safe;
    }
    __syncthreads();
    const uint32_t thin_a = opaque(static_cast<uint32_t>(__cvta_generic_to_shared(s_thin)));
#define THIN_U_LEN (ERRORS ? lds32(thin_a) : u_len_r)
#define THIN_LAST (ERRORS ? lds32(thin_a + 4u) : last_r)
#define THIN_LIMIT (ERRORS ? lds32(thin_a + 8u) : limit_r)
#define THIN_SAFE (ERRORS ? lds32(thin_a + 12u) : safe)
#else
#define THIN_U_LEN u_len_r
#define THIN_LAST last_r
#define THIN_LIMIT limit_r
#define THIN_SAFE safe
#endif
    const uint32_t u_len_r = T.u_len, n_useful = T.n_useful, last_r = T.begin + T.len - 1u, limit_r = T.begin + T.tail_off - 1u;
    const uint32_t n_blocks = (n_useful + 1u) >> 1;
    const uint32_t below_code = below_threshold_code(M);
    for (uint32_t j0 = warp * 32u; j0 < n_blocks; j0 += kStagedThreads) {  // warp-uniform trip count
      const uint32_t j = j0 + lane;
      const uint4 u = philox4x32_10(make_uint4(j, T.id, 0u, M.seed));
      uint32_t ecodes = 0;  // pair_error_codes of reads 2j and 2j + 1: all that stays of their error block
      if (ERRORS) ecodes = pair_error_codes(philox4x32_10(make_uint4(j, T.id, kPurposeErrorBlock, M.seed)), M.err_thr, below_code);
#pragma unroll (ERRORS ? kThinUnroll : 2)
      for (uint32_t k = 0; k < 2u; ++k) {
        const uint32_t t = __umulhi(k ? u.z : u.x, THIN_U_LEN), u_hap = k ? u.w : u.y;
        if (2u * j + k < n_useful) {
          uint32_t i;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(i) : "r"(slot_a + (t >> 6) * 2u));
          uint32_t c = lds32(cum_a + i * 4u);
          while (c <= t) {
            ++i;
            c = lds32(cum_a + i * 4u);
          }
          const uint32_t we = i < n_u ? min(lds32(S.SV.rec + i * 16u), THIN_LIMIT) : THIN_LAST;
          const uint32_t xs = useful_position(t, c, we), off = xs - T.begin;
          if (off >= THIN_SAFE && !fits(u_hap, off, R)) {
            // rare: the last read length of a fragment (error models: counted in shared memory, the walk needs the register)
            if (ERRORS) atomicAdd(&s_dropped, 1u); else ++dropped;
          } else {
            if (i == n_u) {  // a draw in the tail zone: the first locus at or after the read's start, if any
              uint32_t lo = 0, hi = n;
              while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (lds32(S.SV.rec + mid * 16u) < xs) lo = mid + 1u; else hi = mid;
              }
              i = lo;
            }
            staged_read<ERRORS>(S, T, F, M, depth, alt, make_uint4(off, u_hap, 2u * j + k, i), (ecodes >> (4u * k)) & 15u);
          }
        }
      }
      flush_carried(false);
    }
    flush_carried(true);
  } else {
  // Philox block j: two single-end templates (2j, 2j+1) or one paired template (mates 2j, 2j+1)
  const uint32_t n_blocks = PAIRED ? T.n_templates : (T.n_templates + 1u) >> 1;
  const uint32_t n_second = PAIRED ? T.n_templates : T.n_templates >> 1;  // blocks whose second read exists
  for (uint32_t j0 = warp * 32u; j0 < n_blocks; j0 += kStagedThreads) {  // warp-uniform trip count
    const uint32_t j = j0 + lane;
    const uint4 u = philox4x32_10(make_uint4(j, T.id, 0u, M.seed));
    if (PAIRED) {
      const uint32_t off = __umulhi(u.x, len), off2 = off + R + draw_insert(M, u.z);
      bool ok = j < n_blocks;
      if (off >= safe && ok && !fits(u.y, off, off2 - off + R)) {
        ok = false;
        ++dropped;
      }
      const uint2 d0 = S.first_locus(off >> shift), d1 = S.first_locus(off2 >> shift);
      Q.push(ok && d0.y < off + R, make_uint4(off, u.y, 2u * j, d0.x), lanes_below);
      Q.push(ok && d1.y < off2 + R, make_uint4(off2, u.y, 2u * j + 1u, d1.x), lanes_below);
    } else {
      const uint32_t off0 = __umulhi(u.x, len), off1 = __umulhi(u.z, len);
      bool ok0 = j < n_blocks, ok1 = j < n_second;
      if (max(off0, off1) >= safe) {  // last tile of a fragment only
        if (ok0 && !fits(u.y, off0, R)) {
          ok0 = false;
          ++dropped;
        }
        if (ok1 && !fits(u.w, off1, R)) {
          ok1 = false;
          ++dropped;
        }
      }
      const uint2 d0 = S.first_locus(off0 >> shift), d1 = S.first_locus(off1 >> shift);
      Q.push(ok0 && d0.y < off0 + R, make_uint4(off0, u.y, 2u * j, d0.x), lanes_below);
      Q.push(ok1 && d1.y < off1 + R, make_uint4(off1, u.w, 2u * j + 1u, d1.x), lanes_below);
    }
    drain();
  }
  __syncwarp();
  if (lane < (Q.tail - Q.base) / 16u) staged_read_queued<ERRORS>(S, T, F, M, depth, alt, lds128(Q.base + lane * 16u));
  flush_carried(true);
  }
  __syncthreads();

  // ---- flush: one reduction per touched counter, coalesced over consecutive loci / rows
  uint32_t* depth_s = depth + static_cast<size_t>(T.sample) * F.n_loci + T.l0;
  for (uint32_t i = threadIdx.x; i < n; i += kStagedThreads) {
    const uint32_t v = s_depth[i];
    if (v) atomicAdd(depth_s + i, v);
  }
  uint32_t* alt_s = alt + static_cast<size_t>(T.sample) * F.n_mut + T.r0;
  for (uint32_t r = threadIdx.x; r < T.n_rows; r += kStagedThreads) {
    const uint32_t v = s_alt[r];
    if (v) atomicAdd(alt_s + r, v);
  }
  // reads placed = every template of the tile but the dropped ones
  const uint32_t mates = PAIRED ? 2u : 1u;
  block_sub_u64(static_cast<unsigned long long>(T.n_templates) * mates, (dropped + (threadIdx.x == 0 ? s_dropped : 0u)) * mates, n_reads);
}

// ------------------------------------------- global-memory sampler (fallback, trace)
template <bool TRACE>
__device__ __forceinline__ void global_read(const Tile& T, const DevForest& F, const SeqModel& M, const GlobalView& GV,
                                            uint32_t chr_l1, uint32_t read_id, uint32_t xs, uint32_t h,
                                            uint32_t frag_end, DevPlacement* trace, uint32_t* trace_masks,
                                            unsigned long long trace_cap, unsigned long long* trace_n) {
  uint32_t mask[PCS_ERRMASK_WORDS];
  if (TRACE) {
#pragma unroll
    for (int i = 0; i < PCS_ERRMASK_WORDS; ++i) mask[i] = 0;
  }
  const ErrDraw err{M, read_id, T.id, TRACE ? mask : nullptr};
  Walk w;
  w.init(xs, M.read_size, frag_end);
  walk_global(GV, lower_bound_pos(F.locus_pos, T.l0, chr_l1, xs), chr_l1, h, M.read_size, frag_end, w, err);
  if (TRACE) {
    unsigned long long idx = atomicAdd(trace_n, 1ull);
    if (idx < trace_cap) {
      trace[idx] = DevPlacement{h, xs, frag_end, T.chr | (T.sample << 16)};
      if (trace_masks)
        for (int i = 0; i < PCS_ERRMASK_WORDS; ++i) trace_masks[idx * PCS_ERRMASK_WORDS + i] = mask[i];
    }
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(256)
sample_tiles_global_kernel(const Tile* __restrict__ tiles, const Entry* __restrict__ entries,
                           const uint32_t* __restrict__ entry_lo, DevForest F, SeqModel M,
                           uint32_t* __restrict__ depth, uint32_t* __restrict__ alt,
                           unsigned long long* __restrict__ n_reads, DevPlacement* __restrict__ trace,
                           uint32_t* __restrict__ trace_masks, unsigned long long trace_cap,
                           unsigned long long* __restrict__ trace_n) {
  const Tile T = tiles[blockIdx.x];
  const uint32_t chr_l1 = F.chr_locus_off[T.chr + 1];
  const GlobalView GV{F.locus_pos, F.locus_inst_off, F.inst,
                      TRACE ? nullptr : depth + static_cast<size_t>(T.sample) * F.n_loci,
                      TRACE ? nullptr : alt + static_cast<size_t>(T.sample) * F.n_mut};
  const uint32_t R = M.read_size;
  const Entry* ent = entries + T.entry_off;
  const uint32_t* ent_lo = entry_lo + T.entry_off;
  uint32_t placed = 0;

  if (M.paired) {
    for (uint32_t t = threadIdx.x; t < T.n_templates; t += blockDim.x) {
      const uint4 u = philox4x32_10(make_uint4(t, T.id, 0u, M.seed));
      const uint32_t ins = draw_insert(M, u.z);
      Template tp;
      if (!place(T, ent, ent_lo, F, u.x, u.y, 2u * R + ins, tp)) continue;
      global_read<TRACE>(T, F, M, GV, chr_l1, 2u * t, tp.x, tp.h, tp.frag_end, trace, trace_masks, trace_cap, trace_n);
      global_read<TRACE>(T, F, M, GV, chr_l1, 2u * t + 1u, tp.x + R + ins, tp.h, tp.frag_end, trace, trace_masks,
                         trace_cap, trace_n);
      placed += 2;
    }
  } else {
    // every read of the tile, the ones a thinned tile's sampler never draws included (TileReads)
    __shared__ uint32_t s_cum[kMaxThinLoci + 2];
    TileReads TR;
    TR.init(T, F, M, s_cum);
    for (uint32_t r = threadIdx.x; r < T.n_templates; r += blockDim.x) {
      uint32_t off, u_hap, first;
      TR.read(r, off, u_hap, first);
      Template tp;
      if (place_at(T, ent, ent_lo, F, off, u_hap, R, tp)) {
        global_read<TRACE>(T, F, M, GV, chr_l1, r, tp.x, tp.h, tp.frag_end, trace, trace_masks, trace_cap, trace_n);
        ++placed;
      }
    }
  }
  block_add_u64(placed, n_reads);
}

// ------------------------------------------------------------- SAM records
// Materialise reads: bases from the reference + carried SIDs, CIGAR, qualities, sequencing errors.
// Same templates, haplotypes and SID error outcomes as the counting kernels (same Philox counters),
// so the tables and the SAM files of one call describe the same reads.  Ordinary bases draw their
// error from block (read, tile, 0x80000000 | offset/4 [constant] or offset [random], seed).
struct CigarBuilder {
  uint32_t* ops;
  uint32_t n = 0;
  bool overflow = false;
  __device__ void push(uint32_t op, uint32_t len) {
    if (len == 0) return;
    if (n > 0 && (ops[n - 1] & 15u) == op) {
      ops[n - 1] += len << 4;
    } else if (n < kMaxCigar) {
      ops[n++] = (len << 4) | op;
    } else {
      overflow = true;
    }
  }
};

__device__ __forceinline__ uint8_t substitute_base(uint8_t base, uint32_t offset) {
  const char acgt[4] = {'A', 'C', 'G', 'T'};
  int code = base == 'A' ? 0 : base == 'C' ? 1 : base == 'G' ? 2 : base == 'T' ? 3 : -1;
  if (code < 0) return base;  // N stays N
  return static_cast<uint8_t>(acgt[(code + 1 + static_cast<int>(offset % 3u)) & 3]);
}

struct BaseWriter {
  const SeqModel& M;
  uint32_t read, tile;
  uint8_t* seq;
  uint8_t* qual;
  uint32_t* mask;
  uint32_t cached_block = 0xffffffffu;
  uint4 cached{};
  __device__ uint8_t phred(float e) const {
    float q = -10.0f * log10f(fmaxf(fminf(e, 1.0f), 1e-5f));
    int qi = static_cast<int>(q + 0.5f);
    qi = qi < 2 ? 2 : (qi > 41 ? 41 : qi);
    return static_cast<uint8_t>(33 + qi);
  }
  // base at read offset o; sid: it is base b of the hit-th carried SID (its error draw is the counting kernels')
  __device__ void put(uint32_t o, uint8_t base, bool sid, uint32_t hit, uint32_t b) {
    bool err = false;
    uint8_t q = 'I';
    if (sid && M.sequencer != PCS_SEQ_ERRORLESS) {
      float e;
      err = sid_base_error(err_model(M), read, tile, hit, b, o, &e);
      q = (err && M.sequencer == PCS_SEQ_BASIC_CONSTANT) ? '#' : phred(e);
    } else if (M.sequencer == PCS_SEQ_BASIC_CONSTANT) {
      const uint32_t block = 0x80000000u | (o >> 2);
      if (block != cached_block) {
        cached = philox4x32_10(make_uint4(read, tile, block, M.seed));
        cached_block = block;
      }
      const uint32_t word = (o & 3u) == 0 ? cached.x : (o & 3u) == 1 ? cached.y : (o & 3u) == 2 ? cached.z : cached.w;
      err = word < M.err_thr;
      q = err ? '#' : phred(M.error_rate);
    } else if (M.sequencer == PCS_SEQ_BASIC_RANDOM) {
      const uint4 u = philox4x32_10(make_uint4(read, tile, 0x80000000u | o, M.seed));
      const float z = sqrtf(-2.0f * __logf(u01(u.x))) * cospif(2.0f * u01(u.y));
      const float e = M.error_rate * ramp(o, M.read_size) * __expf(kQualSigma * z - 0.5f * kQualSigma * kQualSigma);
      err = u01(u.z) < fminf(e, 1.0f);
      q = phred(e);
    }
    if (err) {
      base = substitute_base(base, o);
      if (mask && o < 32u * PCS_ERRMASK_WORDS) mask[o >> 5] |= 1u << (o & 31);
    }
    seq[o] = base;
    qual[o] = q;
  }
};

__device__ void materialize_read(const DevForest& F, const SeqModel& M, const SeqData& D, uint32_t chr, uint32_t l_first,
                                 uint32_t l_end, uint32_t h, uint32_t xs, uint32_t frag_end, BaseWriter& W,
                                 CigarBuilder& C, uint32_t& len_out) {
  const uint32_t R = M.read_size;
  const uint8_t* ref = D.ref + D.chr_ref_off[chr] - 1;  // ref[p] = base at 1-based position p
  uint32_t q = xs, o = 0, hit = 0;
  uint32_t stop = min(xs + R, frag_end + 1u);
  for (uint32_t i = l_first; i < l_end && o < R; ++i) {
    const uint32_t p = __ldg(F.locus_pos + i);
    if (p >= stop) break;
    if (p < q) continue;
    const uint32_t k1 = __ldg(F.locus_inst_off + i + 1);
    for (uint32_t k = __ldg(F.locus_inst_off + i); k < k1; ++k) {
      const uint4 in = __ldg(F.inst + k);
      if (h - in.x >= in.y) continue;
      for (uint32_t pp = q; pp < p; ++pp) W.put(o++, ref[pp], false, 0u, 0u);
      C.push(0u, p - q);
      const uint32_t ref_len = in.w & 0xffu, alt_len = (in.w >> 8) & 0xffu;
      const uint32_t consumed = min(alt_len, R - o);
      const uint8_t* alt = D.alt + __ldg(D.alt_off + in.z);
      for (uint32_t b = 0; b < consumed; ++b) W.put(o++, alt[b], true, hit, b);
      const uint32_t m = min(consumed, ref_len);
      C.push(0u, m);
      C.push(1u, consumed - m);
      ++hit;
      q = p + ref_len;
      if (o < R && ref_len > alt_len) C.push(2u, ref_len - alt_len);
      stop = min(q + (R - o), frag_end + 1u);
      break;  // at most one SID per position on a haplotype
    }
  }
  if (o < R && q < stop) {
    const uint32_t last = min(q + (R - o), frag_end + 1u);
    for (uint32_t pp = q; pp < last; ++pp) W.put(o++, ref[pp], false, 0u, 0u);
    C.push(0u, last - q);
  }
  len_out = o;
}

__global__ void __launch_bounds__(128)
materialize_tiles_kernel(const Tile* __restrict__ tiles, const Entry* __restrict__ entries,
                         const uint32_t* __restrict__ entry_lo, DevForest F, SeqModel M, SeqData D,
                         SamHeader* __restrict__ hdr, uint32_t* __restrict__ masks,
                         uint8_t* __restrict__ seq, uint8_t* __restrict__ qual, unsigned long long cap,
                         unsigned long long* __restrict__ n_out) {
  const Tile T = tiles[blockIdx.x];
  const uint32_t chr_l1 = F.chr_locus_off[T.chr + 1];
  const uint32_t R = M.read_size;
  const Entry* ent = entries + T.entry_off;
  const uint32_t* ent_lo = entry_lo + T.entry_off;
  __shared__ uint32_t s_cum[kMaxThinLoci + 2];
  TileReads TR;
  TR.init(T, F, M, s_cum);  // single-end: every read of the tile, thinned or not
  // paired: Philox block j is template j (mates 2j, 2j+1); single-end: one read per trip
  const uint32_t n_trips = T.n_templates;
  for (uint32_t j = threadIdx.x; j < n_trips; j += blockDim.x) {
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (M.paired) u = philox4x32_10(make_uint4(j, T.id, 0u, M.seed));
    for (uint32_t k = 0; k < (M.paired ? 2u : 1u); ++k) {
      Template tp;
      uint32_t ins = 0, xs, mate_start = 0;
      int32_t tlen = 0;
      if (M.paired) {
        ins = draw_insert(M, u.z);
        if (!place(T, ent, ent_lo, F, u.x, u.y, 2u * R + ins, tp)) break;
        xs = tp.x + k * (R + ins);
        mate_start = tp.x + (1u - k) * (R + ins);
        tlen = static_cast<int32_t>(2u * R + ins) * (k == 0 ? 1 : -1);
      } else {
        uint32_t off, u_hap, first;
        TR.read(j, off, u_hap, first);
        if (!place_at(T, ent, ent_lo, F, off, u_hap, R, tp)) continue;
        xs = tp.x;
      }
      const unsigned long long idx = atomicAdd(n_out, 1ull);
      if (idx >= cap) continue;
      SamHeader H{};
      H.hap = tp.h; H.start = xs; H.frag_end = tp.frag_end; H.chr_sample = T.chr | (T.sample << 16);
      H.read_id = M.paired ? 2u * j + k : j; H.tile_id = T.id; H.flags = (M.paired ? 1u : 0u) | (M.paired && k ? 2u : 0u);
      H.mate_start = mate_start; H.tlen = tlen;
      uint32_t* mask = masks ? masks + idx * PCS_ERRMASK_WORDS : nullptr;  // the SAM writer does not need them
      if (mask)
        for (int i = 0; i < PCS_ERRMASK_WORDS; ++i) mask[i] = 0;
      BaseWriter W{M, H.read_id, T.id, seq + idx * R, qual + idx * R, mask};
      CigarBuilder C{H.cigar};
      materialize_read(F, M, D, T.chr, lower_bound_pos(F.locus_pos, T.l0, chr_l1, xs), chr_l1, tp.h, xs, tp.frag_end, W, C,
                       H.len);
      H.n_cigar = C.n;
      if (C.overflow) H.flags |= 4u;
      hdr[idx] = H;
    }
  }
}

cudaError_t launch_materialize_tiles(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                     const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, const SeqData& D, SamHeader* hdr,
                                     uint32_t* masks, uint8_t* seq, uint8_t* qual, unsigned long long cap,
                                     unsigned long long* n_out) {
  if (n_tiles == 0) return cudaSuccess;
  materialize_tiles_kernel<<<n_tiles, 128, 0, st>>>(tiles, entries, entry_lo, F, M, D, hdr, masks, seq, qual, cap, n_out);
  return cudaGetLastError();
}

// ----------------------------------------------------------- injected reads
__global__ void __launch_bounds__(256)
count_injected_kernel(const DevPlacement* __restrict__ rec, const uint32_t* __restrict__ masks,
                      unsigned long long n, DevForest F, uint32_t R, uint32_t* __restrict__ depth,
                      uint32_t* __restrict__ alt) {
  const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const DevPlacement p = rec[i];
  const uint32_t chr = p.chr_sample & 0xffffu, sample = p.chr_sample >> 16;
  const uint32_t l0 = F.chr_locus_off[chr], l1 = F.chr_locus_off[chr + 1];
  const ErrMaskLookup err{masks ? masks + i * PCS_ERRMASK_WORDS : nullptr};
  const GlobalView GV{F.locus_pos, F.locus_inst_off, F.inst, depth + static_cast<size_t>(sample) * F.n_loci,
                      alt + static_cast<size_t>(sample) * F.n_mut};
  Walk w;
  w.init(p.start, R, p.frag_end);
  walk_global(GV, lower_bound_pos(F.locus_pos, l0, l1, p.start), l1, p.hap, R, p.frag_end, w, err);
}

// ------------------------------------------------------------------ finalize
__global__ void finalize_kernel(const uint32_t* __restrict__ depth, const uint32_t* __restrict__ row_locus,
                                uint32_t n_samples, uint32_t n_loci, uint32_t n_mut,
                                uint32_t* __restrict__ coverage) {
  const size_t total = static_cast<size_t>(n_samples) * n_mut;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint32_t s = static_cast<uint32_t>(i / n_mut), row = static_cast<uint32_t>(i % n_mut);
    coverage[i] = depth[static_cast<size_t>(s) * n_loci + __ldg(row_locus + row)];
  }
}

__global__ void sum_u32_kernel(const uint32_t* __restrict__ v, size_t n, unsigned long long* __restrict__ out) {
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    acc += v[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// ------------------------------------------------------- result assembly
// get_result_dataframe()/get_active_mutations()/add_sample_statistics() (src/seq_simulation.cpp:92-181) on the
// device: a row is active iff some sample saw it (occurrences > 0) or -- include_non_sequenced_mutations -- some
// sequenced cell carries it (`carried`, host-made, may be null); active rows are compacted IN ROW ORDER (the
// std::map<SID, ...> order of the reference) into column-major per-sample columns: occurrences, coverage =
// depth at the row's locus (this replaces finalize_kernel on this path), VAF = occurrences / coverage as double
// (0 where the sample never covered the locus, :129-131).  HBM-bound: reads the occurrence tables twice, writes
// the compact columns once.
constexpr int kActiveThreads = 256, kActivePerThread = 8;
constexpr uint32_t kActiveRowsPerBlock = kActiveThreads * kActivePerThread;

__device__ __forceinline__ bool row_active(const uint32_t* __restrict__ occ, const uint8_t* __restrict__ carried,
                                           uint32_t S, uint32_t M, uint32_t m) {
  if (carried && carried[m]) return true;
  uint32_t any = 0;
  for (uint32_t s = 0; s < S; ++s) any |= __ldg(occ + static_cast<size_t>(s) * M + m);
  return any != 0;
}

__global__ void __launch_bounds__(kActiveThreads)
active_count_kernel(const uint32_t* __restrict__ occ, const uint8_t* __restrict__ carried, uint32_t S, uint32_t M,
                    uint32_t* __restrict__ block_count) {
  const uint32_t base = blockIdx.x * kActiveRowsPerBlock;
  uint32_t n = 0;
#pragma unroll
  for (int j = 0; j < kActivePerThread; ++j) {
    const uint32_t m = base + j * kActiveThreads + threadIdx.x;
    n += (m < M && row_active(occ, carried, S, M, m)) ? 1u : 0u;
  }
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  __shared__ uint32_t s_w[kActiveThreads / 32];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kActiveThreads / 32; ++w) t += s_w[w];
    block_count[blockIdx.x] = t;
  }
}

// exclusive scan of the block counts in place (one CTA: a few thousand values); total[0] = number of active rows
__global__ void __launch_bounds__(1024) active_scan_kernel(uint32_t* __restrict__ block_count, uint32_t n_blocks,
                                                           uint32_t* __restrict__ total) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t i0 = 0; i0 < n_blocks; i0 += 1024) {
    const uint32_t i = i0 + threadIdx.x;
    const uint32_t v = i < n_blocks ? block_count[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= static_cast<uint32_t>(o)) x += y;
    }
    if (lane == 31) s_w[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_w[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= static_cast<uint32_t>(o)) w += y;
      }
      s_w[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t before = s_carry + (warp ? s_w[warp - 1] : 0u) + x - v;
    if (i < n_blocks) block_count[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) total[0] = s_carry;
}

__global__ void __launch_bounds__(kActiveThreads)
active_scatter_kernel(const uint32_t* __restrict__ occ, const uint32_t* __restrict__ depth,
                      const uint32_t* __restrict__ row_locus, const uint8_t* __restrict__ carried, uint32_t S,
                      uint32_t M, uint32_t L, const uint32_t* __restrict__ block_off, uint32_t n_active,
                      uint32_t* __restrict__ rows_out, uint32_t* __restrict__ occ_c, uint32_t* __restrict__ cov_c,
                      double* __restrict__ vaf_c) {
  __shared__ uint32_t s_w[kActiveThreads / 32];
  const uint32_t base = blockIdx.x * kActiveRowsPerBlock;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t lanes_below;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanes_below));
  uint32_t running = block_off[blockIdx.x];
  for (int j = 0; j < kActivePerThread; ++j) {
    const uint32_t m = base + j * kActiveThreads + threadIdx.x;
    const bool on = m < M && row_active(occ, carried, S, M, m);
    const uint32_t ballot = __ballot_sync(0xffffffffu, on);
    if (lane == 0) s_w[warp] = __popc(ballot);
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kActiveThreads / 32; ++w) {
      const uint32_t c = s_w[w];
      before += w < static_cast<int>(warp) ? c : 0u;
      all += c;
    }
    if (on) {
      const uint32_t k = running + before + __popc(ballot & lanes_below);
      rows_out[k] = m;
      const uint32_t l = __ldg(row_locus + m);
      for (uint32_t s = 0; s < S; ++s) {
        const uint32_t o = __ldg(occ + static_cast<size_t>(s) * M + m);
        const uint32_t c = __ldg(depth + static_cast<size_t>(s) * L + l);
        const size_t at = static_cast<size_t>(s) * n_active + k;
        occ_c[at] = o;
        cov_c[at] = c;
        if (vaf_c) vaf_c[at] = c ? static_cast<double>(o) / static_cast<double>(c) : 0.0;
      }
    }
    running += all;
    __syncthreads();
  }
}

uint32_t active_blocks(uint32_t n_mut) { return (n_mut + kActiveRowsPerBlock - 1) / kActiveRowsPerBlock; }

cudaError_t launch_active_count(cudaStream_t st, const uint32_t* occ, const uint8_t* carried, uint32_t S, uint32_t M,
                                uint32_t* block_count, uint32_t* total) {
  const uint32_t nb = active_blocks(M);
  if (nb == 0) return cudaMemsetAsync(total, 0, sizeof(uint32_t), st);
  active_count_kernel<<<nb, kActiveThreads, 0, st>>>(occ, carried, S, M, block_count);
  active_scan_kernel<<<1, 1024, 0, st>>>(block_count, nb, total);
  return cudaGetLastError();
}

cudaError_t launch_active_scatter(cudaStream_t st, const uint32_t* occ, const uint32_t* depth, const uint32_t* row_locus,
                                  const uint8_t* carried, uint32_t S, uint32_t M, uint32_t L, const uint32_t* block_off,
                                  uint32_t n_active, uint32_t* rows_out, uint32_t* occ_c, uint32_t* cov_c,
                                  double* vaf_c) {
  const uint32_t nb = active_blocks(M);
  if (nb == 0 || n_active == 0) return cudaSuccess;
  active_scatter_kernel<<<nb, kActiveThreads, 0, st>>>(occ, depth, row_locus, carried, S, M, L, block_off, n_active,
                                                       rows_out, occ_c, cov_c, vaf_c);
  return cudaGetLastError();
}

// ------------------------------------------------ instance table on the device
// The flattener of an uploaded forest does not write the instance table (flat.hpp: FlatForest::inst_deferred): 99 %
// of it is the germline, and a germline instance is {interval of its allele mask on its chromosome, row, lengths}.
// Built here from one mask byte and one length pair per row, the somatic placements (sorted by row; a few tens of
// thousands) and eight words per chromosome -- same table, bit for bit, as the host merge
// (tests/test_gpu_genomes.py::test_device_built_instances_equal_the_host_table): inside a row the somatic
// placements first, then the germline one; locus_inst_off[l] = first instance of the locus' first row.
// HBM-bound: reads 7 bytes per row, writes 16 per instance.
__global__ void __launch_bounds__(kActiveThreads)
germline_count_kernel(const uint8_t* __restrict__ mask, uint32_t M, uint32_t* __restrict__ block_count) {
  const uint32_t base = blockIdx.x * kActiveRowsPerBlock;
  uint32_t n = 0;
#pragma unroll
  for (int j = 0; j < kActivePerThread; ++j) {
    const uint32_t m = base + j * kActiveThreads + threadIdx.x;
    n += (m < M && mask[m] != 0) ? 1u : 0u;
  }
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  __shared__ uint32_t s_w[kActiveThreads / 32];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kActiveThreads / 32; ++w) t += s_w[w];
    block_count[blockIdx.x] = t;
  }
}

// first index of the sorted-by-row somatic placements whose row is >= m
__device__ __forceinline__ uint32_t somatic_lower_bound(const uint4* __restrict__ som, uint32_t n_som, uint32_t m) {
  uint32_t lo = 0, hi = n_som;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(&som[mid].z) < m) lo = mid + 1u; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kActiveThreads)
build_instances_kernel(const uint8_t* __restrict__ mask, const uint16_t* __restrict__ meta,
                       const uint32_t* __restrict__ row_locus, const uint32_t* __restrict__ chr_row_off, uint32_t n_chr,
                       const uint32_t* __restrict__ germ_iv, const uint4* __restrict__ som, uint32_t n_som,
                       const uint32_t* __restrict__ block_off, uint32_t M, uint32_t L, uint32_t n_inst,
                       uint4* __restrict__ inst, uint32_t* __restrict__ locus_inst_off) {
  __shared__ uint32_t s_w[kActiveThreads / 32];
  const uint32_t base = blockIdx.x * kActiveRowsPerBlock;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t lanes_below;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lanes_below));
  if (blockIdx.x == 0 && threadIdx.x == 0) locus_inst_off[L] = n_inst;
  uint32_t running = block_off[blockIdx.x];  // germline rows before this block
  for (int j = 0; j < kActivePerThread; ++j) {
    const uint32_t m = base + j * kActiveThreads + threadIdx.x;
    const uint32_t k = m < M ? mask[m] : 0u;
    const uint32_t ballot = __ballot_sync(0xffffffffu, k != 0u);
    if (lane == 0) s_w[warp] = __popc(ballot);
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kActiveThreads / 32; ++w) {
      const uint32_t c = s_w[w];
      before += w < static_cast<int>(warp) ? c : 0u;
      all += c;
    }
    if (m < M) {
      uint32_t si = somatic_lower_bound(som, n_som, m);
      uint32_t at = running + before + __popc(ballot & lanes_below) + si;  // the instances of every earlier row
      if (m == 0u || __ldg(row_locus + m) != __ldg(row_locus + m - 1u)) locus_inst_off[__ldg(row_locus + m)] = at;
      for (; si < n_som && __ldg(&som[si].z) == m; ++si) inst[at++] = __ldg(som + si);
      if (k != 0u) {
        uint32_t lo = 0, hi = n_chr;  // the chromosome of the row: last c with chr_row_off[c] <= m
        while (hi - lo > 1u) {
          const uint32_t mid = (lo + hi) >> 1;
          if (__ldg(chr_row_off + mid) <= m) lo = mid; else hi = mid;
        }
        const uint32_t* iv = germ_iv + lo * 8u;
        inst[at] = make_uint4(__ldg(iv + (k & 3u)), __ldg(iv + 4u + (k & 3u)), m, meta[m]);
      }
    }
    running += all;
    __syncthreads();
  }
}

cudaError_t launch_build_instances(cudaStream_t st, const uint8_t* mask, const uint16_t* meta, const uint32_t* row_locus,
                                   const uint32_t* chr_row_off, uint32_t n_chr, const uint32_t* germ_iv, const uint4* som,
                                   uint32_t n_som, uint32_t* block_scratch, uint32_t M, uint32_t L, uint32_t n_inst,
                                   uint4* inst, uint32_t* locus_inst_off) {
  const uint32_t nb = active_blocks(M);
  if (nb == 0) return cudaMemsetAsync(locus_inst_off, 0, sizeof(uint32_t) * (static_cast<size_t>(L) + 1), st);
  germline_count_kernel<<<nb, kActiveThreads, 0, st>>>(mask, M, block_scratch);
  active_scan_kernel<<<1, 1024, 0, st>>>(block_scratch, nb, block_scratch + nb);
  build_instances_kernel<<<nb, kActiveThreads, 0, st>>>(mask, meta, row_locus, chr_row_off, n_chr, germ_iv, som, n_som,
                                                        block_scratch, M, L, n_inst, inst, locus_inst_off);
  return cudaGetLastError();
}

// ------------------------------------------------------- coverage tracks
// Binned depth along the genome (the tables behind depth-ratio plots,
// R/plot_genome_wide_mutations.R:75-108, want depth between the mutations too).  Reads are never materialised,
// so the track is made by drawing the templates of the plan ONCE MORE -- same Philox counters, same starts, same
// templates dropped at fragment ends as the counting kernels: the track and the tables describe the same reads --
// without haplotypes or loci: a read adds its R reference bases (its span in the frame it starts in) to the bins
// of 2^bin_shift bp it overlaps.  One CTA per tile, bins of the tile in shared memory, one red.global per touched
// bin.  About a third of the sampler's work, and only when a track is asked for.
template <bool PAIRED>
__global__ void __launch_bounds__(256)
coverage_track_kernel(const Tile* __restrict__ tiles, const Entry* __restrict__ entries, DevForest F, SeqModel M,
                      uint32_t bin_shift, const uint64_t* __restrict__ chr_bin_off, uint64_t n_bins,
                      uint32_t* __restrict__ track) {
  extern __shared__ uint32_t s_bins[];
  const Tile T = tiles[blockIdx.x];
  const uint32_t R = M.read_size;
  const uint32_t first_bin = T.begin >> bin_shift;
  const uint32_t n_local = ((T.begin + T.len + M.reach) >> bin_shift) - first_bin + 1u;
  for (uint32_t b = threadIdx.x; b < n_local; b += blockDim.x) s_bins[b] = 0;
  __shared__ uint32_t s_safe;
  if (threadIdx.x == 0) {
    uint32_t min_fe = 0xffffffffu;
    for (uint32_t e = 0; e < T.n_entries; ++e) min_fe = min(min_fe, __ldg(&entries[T.entry_off + e].frag_end));
    const long long lim = static_cast<long long>(min_fe) + 2 - static_cast<long long>(M.reach) - T.begin;
    s_safe = lim <= 0 ? 0u : (lim >= static_cast<long long>(T.len) ? T.len : static_cast<uint32_t>(lim));
  }
  __syncthreads();
  const uint32_t safe = s_safe;
  const Entry* ent = entries + T.entry_off;
  // does the template fit the fragment its haplotype draw selects?  (only asked past `safe`)
  auto fits = [&](uint32_t u_hap, uint32_t off, uint32_t tlen) {
    uint32_t e = 0;
    while (u_hap > __ldg(&ent[e].thr)) ++e;
    return T.begin + off + (tlen - 1u) <= __ldg(&ent[e].frag_end);
  };
  auto add = [&](uint32_t off) {  // R bases from tile offset `off` on
    const uint32_t x = T.begin + off, b0 = (x >> bin_shift) - first_bin;
    const uint32_t in_first = min(R, ((x >> bin_shift) + 1u << bin_shift) - x);
    atomicAdd(&s_bins[b0], in_first);
    for (uint32_t left = R - in_first, b = b0 + 1u; left != 0u; ++b) {  // R may span several bins
      const uint32_t n = min(left, 1u << bin_shift);
      atomicAdd(&s_bins[b], n);
      left -= n;
    }
  };
  __shared__ uint32_t s_cum[kMaxThinLoci + 2];
  TileReads TR;
  if (!PAIRED) TR.init(T, F, M, s_cum);  // single-end: every read of the tile, thinned or not
  for (uint32_t j = threadIdx.x; j < T.n_templates; j += blockDim.x) {
    if (PAIRED) {
      const uint4 u = philox4x32_10(make_uint4(j, T.id, 0u, M.seed));
      const uint32_t off = __umulhi(u.x, T.len), off2 = off + R + draw_insert(M, u.z);
      if (off >= safe && !fits(u.y, off, off2 - off + R)) continue;
      add(off);
      add(off2);
    } else {
      uint32_t off, u_hap, first;
      TR.read(j, off, u_hap, first);
      if (off < safe || fits(u_hap, off, R)) add(off);
    }
  }
  __syncthreads();
  uint32_t* out = track + static_cast<size_t>(T.sample) * n_bins + chr_bin_off[T.chr] + first_bin;
  const uint32_t chr_bins = static_cast<uint32_t>(chr_bin_off[T.chr + 1] - chr_bin_off[T.chr]);
  for (uint32_t b = threadIdx.x; b < n_local; b += blockDim.x) {
    const uint32_t v = s_bins[b];
    if (v && first_bin + b < chr_bins) atomicAdd(out + b, v);
  }
}

cudaError_t launch_coverage_track(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                  const DevForest& F, const SeqModel& M, uint32_t bin_shift, uint32_t max_tile_len,
                                  const uint64_t* chr_bin_off, uint64_t n_bins, uint32_t* track) {
  if (n_tiles == 0) return cudaSuccess;
  const size_t smem = (((static_cast<size_t>(max_tile_len) + M.reach) >> bin_shift) + 2) * sizeof(uint32_t);
  if (smem > 200u * 1024u) return cudaErrorInvalidValue;  // bins this fine do not fit: the caller picks a wider bin
  auto kern = M.paired ? coverage_track_kernel<true> : coverage_track_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kern<<<n_tiles, 256, smem, st>>>(tiles, entries, F, M, bin_shift, chr_bin_off, n_bins, track);
  return cudaGetLastError();
}

// ----------------------------------------------------------------- launchers
size_t staged_smem_bytes(const StageDims& D, bool errors) {
  size_t b = static_cast<size_t>(D.max_loci) * (sizeof(uint4) + sizeof(uint32_t));
  b += (1 + static_cast<size_t>(kStagedThreads / 32) * (kQueueSlots + (errors ? kCarriedSlots : 0u)) + kMaxStagedEntries) *
       sizeof(uint4);
  b += kMaxStagedEntries * sizeof(uint32_t);  // low words of the entries' leaf scales
  b += static_cast<size_t>(D.max_rows) * sizeof(uint32_t);
  b += static_cast<size_t>(D.max_buckets) * sizeof(uint2);
  return (b + 15) & ~static_cast<size_t>(15);
}

// CTAs per SM the kernel is compiled for: PCS_MIN_CTAS overrides.  The error-model variants need ~78
// registers uncapped; capping them at 64 (4 CTAs) spills a little and still wins on occupancy.
static int staged_min_ctas(bool errors, bool paired) {
  static const int v = [] {
    const char* s = std::getenv("PCS_MIN_CTAS");
    const int x = s ? std::atoi(s) : 0;
    return (x >= 3 && x <= 8) ? x : 0;
  }();
  // measured on C3 (profiles/r02_*): errorless single-end (thinned tiles) 8.85 / 8.85 / 8.68 / 9.11 ms at 3 / 4 / 5 / 6;
  // the unthinned loop (paired reads) is fastest at 3
  return v ? v : (errors ? 4 : (paired ? kDefaultMinCtas : 5));
}

template <bool PAIRED, bool ERRORS, int MIN_CTAS>
static cudaError_t launch_staged_occ(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                     const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, const StageDims& D, uint32_t* depth,
                                     uint32_t* alt, unsigned long long* n_reads) {
  const size_t smem = staged_smem_bytes(D, ERRORS);
  auto kern = sample_tiles_staged_kernel<PAIRED, ERRORS, MIN_CTAS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kern<<<n_tiles, kStagedThreads, smem, st>>>(tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
  return cudaGetLastError();
}

template <bool PAIRED, bool ERRORS>
static cudaError_t launch_staged(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                 const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, const StageDims& D, uint32_t* depth,
                                 uint32_t* alt, unsigned long long* n_reads) {
  switch (staged_min_ctas(ERRORS, PAIRED)) {
    case 3: return launch_staged_occ<PAIRED, ERRORS, 3>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
    case 5: return launch_staged_occ<PAIRED, ERRORS, 5>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
    case 6: return launch_staged_occ<PAIRED, ERRORS, 6>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
    case 8: return launch_staged_occ<PAIRED, ERRORS, 8>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
    default: return launch_staged_occ<PAIRED, ERRORS, 4>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
  }
}

cudaError_t launch_sample_tiles_staged(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                       const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, const StageDims& D, uint32_t* depth,
                                       uint32_t* alt, unsigned long long* n_reads) {
  if (n_tiles == 0) return cudaSuccess;
  const bool errors = M.sequencer != PCS_SEQ_ERRORLESS;
  if (M.paired) {
    return errors ? launch_staged<true, true>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads)
                  : launch_staged<true, false>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
  }
  return errors ? launch_staged<false, true>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads)
                : launch_staged<false, false>(st, tiles, n_tiles, entries, entry_lo, F, M, D, depth, alt, n_reads);
}

cudaError_t launch_sample_tiles_global(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                       const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, uint32_t* depth, uint32_t* alt,
                                       unsigned long long* n_reads) {
  if (n_tiles == 0) return cudaSuccess;
  sample_tiles_global_kernel<false><<<n_tiles, 256, 0, st>>>(tiles, entries, entry_lo, F, M, depth, alt, n_reads, nullptr,
                                                              nullptr, 0ull, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_trace_tiles(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                               const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, unsigned long long* n_reads,
                               DevPlacement* trace, uint32_t* trace_masks, unsigned long long cap,
                               unsigned long long* trace_n) {
  if (n_tiles == 0) return cudaSuccess;
  sample_tiles_global_kernel<true><<<n_tiles, 256, 0, st>>>(tiles, entries, entry_lo, F, M, nullptr, nullptr, n_reads, trace,
                                                             trace_masks, cap, trace_n);
  return cudaGetLastError();
}

cudaError_t launch_count_injected(cudaStream_t st, const DevPlacement* rec, const uint32_t* masks,
                                  unsigned long long n, const DevForest& F, uint32_t R, uint32_t* depth,
                                  uint32_t* alt) {
  if (n == 0) return cudaSuccess;
  const unsigned long long blocks = (n + 255) / 256;
  count_injected_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(rec, masks, n, F, R, depth, alt);
  return cudaGetLastError();
}

cudaError_t launch_finalize(cudaStream_t st, const uint32_t* depth, const uint32_t* row_locus, uint32_t n_samples,
                            uint32_t n_loci, uint32_t n_mut, uint32_t* coverage) {
  const size_t total = static_cast<size_t>(n_samples) * n_mut;
  if (total == 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>((total + 255) / 256, 148 * 16));
  finalize_kernel<<<blocks, 256, 0, st>>>(depth, row_locus, n_samples, n_loci, n_mut, coverage);
  return cudaGetLastError();
}

cudaError_t launch_sum_u32(cudaStream_t st, const uint32_t* v, size_t n, unsigned long long* out) {
  if (n == 0) return cudaSuccess;
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 148 * 16));
  sum_u32_kernel<<<blocks, 256, 0, st>>>(v, n, out);
  return cudaGetLastError();
}

}  // namespace pcs
