// kernels.cu -- sm_100a kernels of the read sampler.
//
//   sample_tiles_kernel   one CTA per tile: Philox4x32-10 per template -> start,
//                         haplotype, insert -> locus walk -> depth / alt counts
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
// One read of R bases from haplotype `h` starting at reference position x.
// depth_s / alt_s: count tables of the read's sample (nullptr in trace mode).
template <class Err>
__device__ __forceinline__ void walk_read(const DevForest& F, uint32_t l_first, uint32_t l_end, uint32_t h,
                                          uint32_t x, uint32_t R, uint32_t frag_end, uint32_t* depth_s,
                                          uint32_t* alt_s, const Err& err) {
  uint32_t q = x, rem = R, hit = 0;
  for (uint32_t i = l_first; i < l_end; ++i) {
    const uint32_t p = __ldg(F.locus_pos + i);
    if (p > frag_end) break;
    if (p < q) continue;  // inside the reference bases a carried SID replaced
    const uint32_t gap = p - q;
    if (gap >= rem) break;
    rem -= gap;
    q = p;
    if (depth_s) atomicAdd(depth_s + i, 1u);
    const uint32_t k1 = __ldg(F.locus_inst_off + i + 1);
    for (uint32_t k = __ldg(F.locus_inst_off + i); k < k1; ++k) {
      const uint4 in = __ldg(F.inst + k);
      if (h - in.x < in.y) {
        const uint32_t ref_len = in.w & 0xffu, alt_len = (in.w >> 8) & 0xffu;
        const uint32_t consumed = min(alt_len, rem);
        const bool bad = err(hit, R - rem, consumed);
        if (!bad && alt_s) atomicAdd(alt_s + in.z, 1u);
        ++hit;
        if (ref_len != 1u || alt_len != 1u) {
          rem -= consumed;
          q = p + ref_len;
        }
      }
    }
    if (rem == 0) break;
  }
}

__device__ __forceinline__ uint32_t lower_bound_pos(const uint32_t* pos, uint32_t lo, uint32_t hi, uint32_t x) {
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(pos + mid) < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------ sampler kernel
template <bool TRACE>
__global__ void __launch_bounds__(256)
sample_tiles_kernel(const Tile* __restrict__ tiles, const Entry* __restrict__ entries, DevForest F, SeqModel M,
                    uint32_t* __restrict__ depth, uint32_t* __restrict__ alt,
                    unsigned long long* __restrict__ n_reads, DevPlacement* __restrict__ trace,
                    uint32_t* __restrict__ trace_masks, unsigned long long trace_cap,
                    unsigned long long* __restrict__ trace_n) {
  const Tile T = tiles[blockIdx.x];
  const uint32_t chr_l1 = F.chr_locus_off[T.chr + 1];
  uint32_t* depth_s = TRACE ? nullptr : depth + static_cast<size_t>(T.sample) * F.n_loci;
  uint32_t* alt_s = TRACE ? nullptr : alt + static_cast<size_t>(T.sample) * F.n_mut;
  const uint2 key = make_uint2(M.seed, T.id);
  const uint32_t R = M.read_size;
  const uint32_t mates = M.paired ? 2u : 1u;
  uint32_t placed = 0;

  for (uint32_t t = threadIdx.x; t < T.n_templates; t += blockDim.x) {
    const uint4 w = philox4x32_10(make_uint4(t, 0u, 0u, 0u), key);
    const uint32_t x = T.begin + __umulhi(w.x, T.len);
    uint32_t e = 0;
    while (e + 1 < T.n_entries && w.y > entries[T.entry_off + e].thr) ++e;
    const Entry E = entries[T.entry_off + e];
    const uint32_t h = __ldg(F.hap_list + E.list_off + __umulhi(w.z, E.list_n));
    uint32_t ins = 0;
    if (M.paired) {
      uint32_t lo = 0, hi = M.insert_n - 1;
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (w.w > __ldg(M.insert_cdf + mid)) lo = mid + 1; else hi = mid;
      }
      ins = M.insert_min + lo;
    }
    const uint64_t tlen = M.paired ? 2ull * R + ins : R;
    if (static_cast<uint64_t>(x) + tlen - 1 > E.frag_end) continue;  // falls off the molecule
    for (uint32_t mate = 0; mate < mates; ++mate) {
      const uint32_t xs = mate == 0 ? x : x + R + ins;
      uint32_t mask[PCS_ERRMASK_WORDS];
      if (TRACE) {
#pragma unroll
        for (int i = 0; i < PCS_ERRMASK_WORDS; ++i) mask[i] = 0;
      }
      ErrDraw err{M, key, t, mate, TRACE ? mask : nullptr};
      const uint32_t l_first = lower_bound_pos(F.locus_pos, T.l0, chr_l1, xs);
      walk_read(F, l_first, chr_l1, h, xs, R, E.frag_end, depth_s, alt_s, err);
      if (TRACE) {
        unsigned long long idx = atomicAdd(trace_n, 1ull);
        if (idx < trace_cap) {
          trace[idx] = DevPlacement{h, xs, E.frag_end, T.chr | (T.sample << 16)};
          if (trace_masks)
            for (int i = 0; i < PCS_ERRMASK_WORDS; ++i) trace_masks[idx * PCS_ERRMASK_WORDS + i] = mask[i];
        }
      }
    }
    placed += mates;
  }

  // reads placed by this CTA
  for (int o = 16; o > 0; o >>= 1) placed += __shfl_xor_sync(0xffffffffu, placed, o);
  __shared__ uint32_t s_placed[8];
  if ((threadIdx.x & 31) == 0) s_placed[threadIdx.x >> 5] = placed;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0;
    for (uint32_t i = 0; i < (blockDim.x + 31) / 32; ++i) tot += s_placed[i];
    if (tot) atomicAdd(n_reads, tot);
  }
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
  ErrMaskLookup err{masks ? masks + i * PCS_ERRMASK_WORDS : nullptr};
  const uint32_t l_first = lower_bound_pos(F.locus_pos, l0, l1, p.start);
  walk_read(F, l_first, l1, p.hap, p.start, R, p.frag_end, depth + static_cast<size_t>(sample) * F.n_loci,
            alt + static_cast<size_t>(sample) * F.n_mut, err);
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
cudaError_t launch_sample_tiles(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                const DevForest& F, const SeqModel& M, uint32_t* depth, uint32_t* alt,
                                unsigned long long* n_reads) {
  if (n_tiles == 0) return cudaSuccess;
  sample_tiles_kernel<false><<<n_tiles, 256, 0, st>>>(tiles, entries, F, M, depth, alt, n_reads, nullptr,
                                                       nullptr, 0ull, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_trace_tiles(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                               const DevForest& F, const SeqModel& M, unsigned long long* n_reads,
                               DevPlacement* trace, uint32_t* trace_masks, unsigned long long cap,
                               unsigned long long* trace_n) {
  if (n_tiles == 0) return cudaSuccess;
  sample_tiles_kernel<true><<<n_tiles, 256, 0, st>>>(tiles, entries, F, M, nullptr, nullptr, n_reads, trace,
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
