// kernels.hpp -- host-callable launchers of kernels.cu
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "dev.hpp"

namespace pcs {

size_t staged_smem_bytes(const StageDims& D, bool errors);
cudaError_t launch_sample_tiles_staged(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                       const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, const StageDims& D, uint32_t* depth,
                                       uint32_t* alt, unsigned long long* n_reads);
cudaError_t launch_sample_tiles_global(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                       const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, uint32_t* depth, uint32_t* alt,
                                       unsigned long long* n_reads);
cudaError_t launch_trace_tiles(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                               const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, unsigned long long* n_reads,
                               DevPlacement* trace, uint32_t* trace_masks, unsigned long long cap,
                               unsigned long long* trace_n);
cudaError_t launch_materialize_tiles(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                     const uint32_t* entry_lo, const DevForest& F, const SeqModel& M, const SeqData& D, SamHeader* hdr,
                                     uint32_t* masks, uint8_t* seq, uint8_t* qual, unsigned long long cap,
                                     unsigned long long* n_out);
cudaError_t launch_count_injected(cudaStream_t st, const DevPlacement* rec, const uint32_t* masks,
                                  unsigned long long n, const DevForest& F, uint32_t R, uint32_t* depth,
                                  uint32_t* alt);
cudaError_t launch_finalize(cudaStream_t st, const uint32_t* depth, const uint32_t* row_locus, uint32_t n_samples,
                            uint32_t n_loci, uint32_t n_mut, uint32_t* coverage);
// result assembly: active rows compacted in row order into column-major per-sample columns
uint32_t active_blocks(uint32_t n_mut);  // length of the block_count scratch array
cudaError_t launch_active_count(cudaStream_t st, const uint32_t* occ, const uint8_t* carried, uint32_t S, uint32_t M,
                                uint32_t* block_count, uint32_t* total);
cudaError_t launch_active_scatter(cudaStream_t st, const uint32_t* occ, const uint32_t* depth, const uint32_t* row_locus,
                                  const uint8_t* carried, uint32_t S, uint32_t M, uint32_t L, const uint32_t* block_off,
                                  uint32_t n_active, uint32_t* rows_out, uint32_t* occ_c, uint32_t* cov_c,
                                  double* vaf_c);
// the instance table of a forest flattened with deferred instances (flat.hpp), built on the device;
// block_scratch: active_blocks(M) + 1 words
cudaError_t launch_build_instances(cudaStream_t st, const uint8_t* mask, const uint16_t* meta, const uint32_t* row_locus,
                                   const uint32_t* chr_row_off, uint32_t n_chr, const uint32_t* germ_iv, const uint4* som,
                                   uint32_t n_som, uint32_t* block_scratch, uint32_t M, uint32_t L, uint32_t n_inst,
                                   uint4* inst, uint32_t* locus_inst_off);
// binned depth track of the plan's reads: track[s][chr_bin_off[chr] + (pos >> bin_shift)] += bases
cudaError_t launch_coverage_track(cudaStream_t st, const Tile* tiles, uint32_t n_tiles, const Entry* entries,
                                  const DevForest& F, const SeqModel& M, uint32_t bin_shift, uint32_t max_tile_len,
                                  const uint64_t* chr_bin_off, uint64_t n_bins, uint32_t* track);
cudaError_t launch_sum_u32(cudaStream_t st, const uint32_t* v, size_t n, unsigned long long* out);

}  // namespace pcs
