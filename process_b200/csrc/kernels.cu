// kernels.cu -- sm_100a kernels of the read sampler.
//
//   sample_tiles_staged_kernel  the hot kernel.  One CTA per tile: the tile's sorted
//                         locus positions, SID instances and a bucket directory are
//                         staged in shared memory; every thread draws templates with
//                         Philox4x32-10 (start, haplotype, insert), walks the loci its
//                         reads span and counts depth / occurrences with shared-memory
//                         atomics; one coalesced red.global per touched counter at the end.
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

#include "../../include/pcs_seq.h"
#include "dev.hpp"
#include "kernels.hpp"

namespace pcs {

// ------------------------------------------------------------------- Philox
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
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

// is any of the `n` read bases starting at `off` a sequencing error?  One Philox
// block per tested base, keyed by (template, mate, hit, base) so the outcome does
// not depend on scheduling.
struct ErrDraw {
  const SeqModel& M;
  uint2 key;
  uint32_t tmpl, mate;
  uint32_t* mask;  // trace mode: error bits found, else nullptr
  __device__ bool operator()(uint32_t hit, uint32_t off, uint32_t n) const {
    if (M.sequencer == PCS_SEQ_ERRORLESS) return false;
    bool any = false;
    for (uint32_t b = 0; b < n; ++b) {
      uint4 w = philox4x32_10(make_uint4(tmpl, 1u + mate, hit, b), key);
      bool e;
      if (M.sequencer == PCS_SEQ_BASIC_CONSTANT) {
        e = w.x < M.err_thr;
      } else {
        float z = sqrtf(-2.0f * __logf(u01(w.x))) * cospif(2.0f * u01(w.y));
        float p = M.error_rate * ramp(off + b, M.read_size) * __expf(kQualSigma * z - 0.5f * kQualSigma * kQualSigma);
        e = u01(w.z) < fminf(p, 1.0f);
      }
      if (e) {
        any = true;
        uint32_t i = off + b;
        if (mask && i < 32u * PCS_ERRMASK_WORDS) mask[i >> 5] |= 1u << (i & 31);
      }
    }
    return any;
  }
};

struct NoErr {
  __device__ __forceinline__ bool operator()(uint32_t, uint32_t, uint32_t) const { return false; }
};

struct ErrMaskLookup {
  const uint32_t* mask;  // nullptr: no errors
  __device__ bool operator()(uint32_t, uint32_t off, uint32_t n) const {
    if (!mask) return false;
    for (uint32_t i = off; i < off + n; ++i)
      if (i < 32u * PCS_ERRMASK_WORDS && ((mask[i >> 5] >> (i & 31)) & 1u)) return true;
    return false;
  }
};

// ---------------------------------------------------------------- locus walk
// State of one read of R bases from haplotype h: next reference position q,
// bases still to place, carried SIDs met so far.
struct Walk {
  uint32_t q, rem, hit;
};

// where the walk reads loci from / counts into
struct GlobalView {
  const uint32_t* pos;       // locus_pos
  const uint32_t* ioff;      // locus_inst_off
  const uint4* inst;
  uint32_t* depth;           // [L] of the sample (nullptr: trace mode)
  uint32_t* alt;             // [M] of the sample
  __device__ __forceinline__ uint32_t position(uint32_t i) const { return __ldg(pos + i); }
  __device__ __forceinline__ uint32_t inst_begin(uint32_t i) const { return __ldg(ioff + i); }
  __device__ __forceinline__ uint32_t inst_end(uint32_t i) const { return __ldg(ioff + i + 1); }
  __device__ __forceinline__ uint4 instance(uint32_t k) const { return __ldg(inst + k); }
  __device__ __forceinline__ void add_depth(uint32_t i) const { if (depth) atomicAdd(depth + i, 1u); }
  __device__ __forceinline__ void add_alt(uint32_t row) const { if (alt) atomicAdd(alt + row, 1u); }
};

struct SharedView {
  const uint32_t* pos;   // [n] staged positions
  const uint32_t* ioff;  // [n+1] instance offsets relative to the tile's first instance
  const uint4* inst;     // staged instances; .z is the row relative to the tile's first row
  uint32_t* depth;       // [n]
  uint32_t* alt;         // [rows]
  __device__ __forceinline__ uint32_t position(uint32_t i) const { return pos[i]; }
  __device__ __forceinline__ uint32_t inst_begin(uint32_t i) const { return ioff[i]; }
  __device__ __forceinline__ uint32_t inst_end(uint32_t i) const { return ioff[i + 1]; }
  __device__ __forceinline__ uint4 instance(uint32_t k) const { return inst[k]; }
  __device__ __forceinline__ void add_depth(uint32_t i) const { atomicAdd(depth + i, 1u); }
  __device__ __forceinline__ void add_alt(uint32_t row) const { atomicAdd(alt + row, 1u); }
};

// Walk loci [i, end) of view V.  Returns the index it stopped at: `end` means the
// view ran out before the read did (the caller may continue in another view).
template <class View, class Err>
__device__ __forceinline__ uint32_t walk_loci(const View& V, uint32_t i, uint32_t end, uint32_t h, uint32_t R,
                                              uint32_t frag_end, Walk& w, const Err& err, bool& done) {
  done = true;
  for (; i < end; ++i) {
    const uint32_t p = V.position(i);
    if (p > frag_end) return i;
    if (p < w.q) continue;  // inside the reference bases a carried SID replaced
    const uint32_t gap = p - w.q;
    if (gap >= w.rem) return i;
    w.rem -= gap;
    w.q = p;
    V.add_depth(i);
    const uint32_t k1 = V.inst_end(i);
    for (uint32_t k = V.inst_begin(i); k < k1; ++k) {
      const uint4 in = V.instance(k);
      if (h - in.x < in.y) {
        const uint32_t ref_len = in.w & 0xffu, alt_len = (in.w >> 8) & 0xffu;
        const uint32_t consumed = min(alt_len, w.rem);
        if (!err(w.hit, R - w.rem, consumed)) V.add_alt(in.z);
        ++w.hit;
        if (ref_len != 1u || alt_len != 1u) {
          w.rem -= consumed;
          w.q = p + ref_len;
        }
      }
    }
    if (w.rem == 0) return i;
  }
  done = false;
  return end;
}

__device__ __forceinline__ uint32_t lower_bound_pos(const uint32_t* pos, uint32_t lo, uint32_t hi, uint32_t x) {
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(pos + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// draw of one template: start, haplotype, insert; false if it falls off its molecule
struct Template {
  uint32_t x, h, ins, frag_end;
};

template <bool PAIRED>
__device__ __forceinline__ bool draw_template(const Tile& T, const Entry* __restrict__ ent, const DevForest& F,
                                              const SeqModel& M, uint2 key, uint32_t t, Template& out) {
  const uint4 w = philox4x32_10(make_uint4(t, 0u, 0u, 0u), key);
  out.x = T.begin + __umulhi(w.x, T.len);
  uint32_t e = 0;
  while (e + 1 < T.n_entries && w.y > ent[e].thr) ++e;
  const Entry E = ent[e];
  out.h = __ldg(F.hap_list + E.list_off + __umulhi(w.z, E.list_n));
  out.frag_end = E.frag_end;
  out.ins = 0;
  uint32_t tlen = M.read_size;
  if (PAIRED) {
    uint32_t lo = 0, hi = M.insert_n - 1;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (w.w > __ldg(M.insert_cdf + mid)) lo = mid + 1; else hi = mid;
    }
    out.ins = M.insert_min + lo;
    tlen = 2u * M.read_size + out.ins;
  }
  // 64-bit: x + tlen may pass 2^32 only for absurd inputs, but stay exact
  return static_cast<uint64_t>(out.x) + tlen - 1 <= E.frag_end;
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

// ---------------------------------------------------- staged sampler kernel
constexpr int kStagedThreads = 256;

template <bool PAIRED, bool ERRORS>
__global__ void __launch_bounds__(kStagedThreads, 4)
sample_tiles_staged_kernel(const Tile* __restrict__ tiles, const Entry* __restrict__ entries, DevForest F, SeqModel M,
                           StageDims D, uint32_t* __restrict__ depth, uint32_t* __restrict__ alt,
                           unsigned long long* __restrict__ n_reads) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint4* s_inst = reinterpret_cast<uint4*>(smem);
  uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_inst + D.max_inst);
  uint32_t* s_ioff = s_pos + D.max_loci;
  uint32_t* s_depth = s_ioff + D.max_loci + 1;
  uint32_t* s_alt = s_depth + D.max_loci;
  uint16_t* s_dir = reinterpret_cast<uint16_t*>(s_alt + D.max_rows);
  __shared__ Entry s_ent[8];

  const Tile T = tiles[blockIdx.x];
  const uint32_t n = T.l1 - T.l0;
  const uint32_t i0 = __ldg(F.locus_inst_off + T.l0);
  const uint32_t n_inst = __ldg(F.locus_inst_off + T.l1) - i0;
  const uint32_t shift = M.dir_shift;
  const uint32_t n_buckets = ((T.len + M.reach) >> shift) + 1;

  // ---- stage the tile
  for (uint32_t i = threadIdx.x; i < n; i += kStagedThreads) {
    s_pos[i] = __ldg(F.locus_pos + T.l0 + i);
    s_depth[i] = 0;
  }
  for (uint32_t i = threadIdx.x; i <= n; i += kStagedThreads) s_ioff[i] = __ldg(F.locus_inst_off + T.l0 + i) - i0;
  for (uint32_t k = threadIdx.x; k < n_inst; k += kStagedThreads) {
    uint4 in = __ldg(F.inst + i0 + k);
    in.z -= T.r0;
    s_inst[k] = in;
  }
  for (uint32_t r = threadIdx.x; r < T.n_rows; r += kStagedThreads) s_alt[r] = 0;
  if (threadIdx.x < T.n_entries && threadIdx.x < 8) s_ent[threadIdx.x] = entries[T.entry_off + threadIdx.x];
  __syncthreads();
  // directory: first staged locus at or after the start of each bucket
  for (uint32_t b = threadIdx.x; b < n_buckets; b += kStagedThreads) {
    const uint32_t x = T.begin + (b << shift);
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (s_pos[mid] < x) lo = mid + 1; else hi = mid;
    }
    s_dir[b] = static_cast<uint16_t>(lo);
  }
  __syncthreads();

  const SharedView SV{s_pos, s_ioff, s_inst, s_depth, s_alt};
  const uint32_t chr_l1 = __ldg(F.chr_locus_off + T.chr + 1);
  const uint2 key = make_uint2(M.seed, T.id);
  const uint32_t R = M.read_size;
  const uint32_t stage_end = T.begin + T.len + M.reach;  // first position whose loci are not staged
  const Entry* ent = T.n_entries <= 8 ? s_ent : entries + T.entry_off;
  uint32_t placed = 0;

  for (uint32_t t = threadIdx.x; t < T.n_templates; t += kStagedThreads) {
    Template tp;
    if (!draw_template<PAIRED>(T, ent, F, M, key, t, tp)) continue;
#pragma unroll
    for (uint32_t mate = 0; mate < (PAIRED ? 2u : 1u); ++mate) {
      const uint32_t xs = mate == 0 ? tp.x : tp.x + R + tp.ins;
      uint32_t i = s_dir[(xs - T.begin) >> shift];
      while (i < n && s_pos[i] < xs) ++i;
      Walk w{xs, R, 0u};
      bool done;
      if (ERRORS) {
        const ErrDraw err{M, key, t, mate, nullptr};
        i = walk_loci(SV, i, n, tp.h, R, tp.frag_end, w, err, done);
        if (!done && T.l1 < chr_l1 && w.q + w.rem > stage_end) {  // a carried deletion stretched the read past the staged loci
          const GlobalView GV{F.locus_pos, F.locus_inst_off, F.inst, depth + static_cast<size_t>(T.sample) * F.n_loci,
                              alt + static_cast<size_t>(T.sample) * F.n_mut};
          walk_loci(GV, T.l1, chr_l1, tp.h, R, tp.frag_end, w, err, done);
        }
      } else {
        const NoErr err;
        i = walk_loci(SV, i, n, tp.h, R, tp.frag_end, w, err, done);
        if (!done && T.l1 < chr_l1 && w.q + w.rem > stage_end) {
          const GlobalView GV{F.locus_pos, F.locus_inst_off, F.inst, depth + static_cast<size_t>(T.sample) * F.n_loci,
                              alt + static_cast<size_t>(T.sample) * F.n_mut};
          walk_loci(GV, T.l1, chr_l1, tp.h, R, tp.frag_end, w, err, done);
        }
      }
    }
    placed += PAIRED ? 2u : 1u;
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
  block_add_u64(placed, n_reads);
}

// ------------------------------------------- global-memory sampler (fallback, trace)
template <bool TRACE>
__global__ void __launch_bounds__(256)
sample_tiles_global_kernel(const Tile* __restrict__ tiles, const Entry* __restrict__ entries, DevForest F, SeqModel M,
                           uint32_t* __restrict__ depth, uint32_t* __restrict__ alt,
                           unsigned long long* __restrict__ n_reads, DevPlacement* __restrict__ trace,
                           uint32_t* __restrict__ trace_masks, unsigned long long trace_cap,
                           unsigned long long* __restrict__ trace_n) {
  const Tile T = tiles[blockIdx.x];
  const uint32_t chr_l1 = F.chr_locus_off[T.chr + 1];
  const GlobalView GV{F.locus_pos, F.locus_inst_off, F.inst,
                      TRACE ? nullptr : depth + static_cast<size_t>(T.sample) * F.n_loci,
                      TRACE ? nullptr : alt + static_cast<size_t>(T.sample) * F.n_mut};
  const uint2 key = make_uint2(M.seed, T.id);
  const uint32_t R = M.read_size;
  const uint32_t mates = M.paired ? 2u : 1u;
  const Entry* ent = entries + T.entry_off;
  uint32_t placed = 0;

  for (uint32_t t = threadIdx.x; t < T.n_templates; t += blockDim.x) {
    Template tp;
    const bool ok = M.paired ? draw_template<true>(T, ent, F, M, key, t, tp) : draw_template<false>(T, ent, F, M, key, t, tp);
    if (!ok) continue;
    for (uint32_t mate = 0; mate < mates; ++mate) {
      const uint32_t xs = mate == 0 ? tp.x : tp.x + R + tp.ins;
      uint32_t mask[PCS_ERRMASK_WORDS];
      if (TRACE) {
#pragma unroll
        for (int i = 0; i < PCS_ERRMASK_WORDS; ++i) mask[i] = 0;
      }
      const ErrDraw err{M, key, t, mate, TRACE ? mask : nullptr};
      Walk w{xs, R, 0u};
      bool done;
      walk_loci(GV, lower_bound_pos(F.locus_pos, T.l0, chr_l1, xs), chr_l1, tp.h, R, tp.frag_end, w, err, done);
      if (TRACE) {
        unsigned long long idx = atomicAdd(trace_n, 1ull);
        if (idx < trace_cap) {
          trace[idx] = DevPlacement{tp.h, xs, tp.frag_end, T.chr | (T.sample << 16)};
          if (trace_masks)
            for (int i = 0; i < PCS_ERRMASK_WORDS; ++i) trace_masks[idx * PCS_ERRMASK_WORDS + i] = mask[i];
        }
      }
    }
    placed += mates;
  }
  block_add_u64(placed, n_reads);
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
  Walk w{p.start, R, 0u};
  bool done;
  walk_loci(GV, lower_bound_pos(F.locus_pos, l0, l1, p.start), l1, p.hap, R, p.frag_end, w, err, done);
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

// ----------------------------------------------------------------- launchers
size_t staged_smem_bytes(const StageDims& D) {
  size_t b = static_cast<size_t>(D.max_inst) * sizeof(uint4);
  b += (static_cast<size_t>(D.max_loci) * 3 + 1 + D.max_rows) * sizeof(uint32_t);
  b += static_cast<size_t>(D.max_buckets) * sizeof(uint16_t);
  return (b + 15) & ~static_cast<size_t>(15);
}

template <bool PAIRED, bool ERRORS>
static cudaError_t launch_staged(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                 const DevForest& F, const SeqModel& M, const StageDims& D, uint32_t* depth,
                                 uint32_t* alt, unsigned long long* n_reads) {
  const size_t smem = staged_smem_bytes(D);
  auto kern = sample_tiles_staged_kernel<PAIRED, ERRORS>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kern<<<n_tiles, kStagedThreads, smem, st>>>(tiles, entries, F, M, D, depth, alt, n_reads);
  return cudaGetLastError();
}

cudaError_t launch_sample_tiles_staged(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                       const DevForest& F, const SeqModel& M, const StageDims& D, uint32_t* depth,
                                       uint32_t* alt, unsigned long long* n_reads) {
  if (n_tiles == 0) return cudaSuccess;
  const bool errors = M.sequencer != PCS_SEQ_ERRORLESS;
  if (M.paired) {
    return errors ? launch_staged<true, true>(st, tiles, n_tiles, entries, F, M, D, depth, alt, n_reads)
                  : launch_staged<true, false>(st, tiles, n_tiles, entries, F, M, D, depth, alt, n_reads);
  }
  return errors ? launch_staged<false, true>(st, tiles, n_tiles, entries, F, M, D, depth, alt, n_reads)
                : launch_staged<false, false>(st, tiles, n_tiles, entries, F, M, D, depth, alt, n_reads);
}

cudaError_t launch_sample_tiles_global(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                       const DevForest& F, const SeqModel& M, uint32_t* depth, uint32_t* alt,
                                       unsigned long long* n_reads) {
  if (n_tiles == 0) return cudaSuccess;
  sample_tiles_global_kernel<false><<<n_tiles, 256, 0, st>>>(tiles, entries, F, M, depth, alt, n_reads, nullptr,
                                                              nullptr, 0ull, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_trace_tiles(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                               const DevForest& F, const SeqModel& M, unsigned long long* n_reads,
                               DevPlacement* trace, uint32_t* trace_masks, unsigned long long cap,
                               unsigned long long* trace_n) {
  if (n_tiles == 0) return cudaSuccess;
  sample_tiles_global_kernel<true><<<n_tiles, 256, 0, st>>>(tiles, entries, F, M, nullptr, nullptr, n_reads, trace,
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
