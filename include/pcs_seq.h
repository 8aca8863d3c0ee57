/*
 * pcs_seq.h -- C ABI of libpcs_seq: the B200-native read sampler that replaces
 * RACES::Mutations::SequencingSimulations::ReadSimulator<>::operator() on the
 * simulate_seq() / simulate_normal_seq() path of caravagnalab/ProCESS.
 *
 * Plain C: pointers and sizes only, no C++ types, no exceptions, no torch types.
 * Every function returns 0 (PCS_OK) or a negative pcs_status; the text of the
 * last error on the calling thread is available from pcs_last_error().
 *
 * Reference interfaces each entry point replaces (paths relative to the
 * ProCESS source tree):
 *   pcs_forest_upload        <- PhylogeneticForest::get_sample_mutations_list(),
 *                               get_normal_sample()      src/seq_simulation.cpp:566,572,651
 *                               (per-cell genome traversal idiom
 *                               src/phylogenetic_forest.cpp:279-376)
 *   pcs_forest_set_groups    <- apply_FACS_labels()/split_by_labels()
 *                                                        src/seq_simulation.cpp:183-243
 *   pcs_plan_create          <- ReadSimulator<> construction + get_bin_dist() +
 *                               get_relevant_chr_set()   src/seq_simulation.cpp:551-564,431-451,299-352
 *   pcs_plan_run / pcs_simulate
 *                            <- simulator(sequencer, mutations_list, chr_ids,
 *                               coverage, normal_sample, purity, ...)
 *                                                        src/seq_simulation.cpp:371,413,423
 *   occurrences[] / coverage[] output tables
 *                            <- SampleStatistics::get_data() / get_coverage()
 *                               as consumed by add_sample_statistics()
 *                                                        src/seq_simulation.cpp:92-140
 *   pcs_active_rows          <- get_active_mutations()   src/seq_simulation.cpp:142-167
 *   pcs_count_injected       <- the counting half of the simulator applied to a
 *                               caller-supplied read-placement list (parity mode;
 *                               no counterpart is exported by the reference)
 *
 * Positions are 1-based chromosome coordinates (the `chr_pos` column of the
 * reference's data frame, src/seq_simulation.cpp:68).
 */
#ifndef PCS_SEQ_H
#define PCS_SEQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCS_ABI_VERSION 1

typedef enum pcs_status {
  PCS_OK = 0,
  PCS_ERR_INVALID = -1,   /* bad argument / malformed forest  (std::domain_error in the shim) */
  PCS_ERR_CUDA = -2,      /* CUDA runtime failure              (std::runtime_error) */
  PCS_ERR_NOMEM = -3,
  PCS_ERR_UNSUPPORTED = -4,
  PCS_ERR_INTERNAL = -5
} pcs_status;

/* kinds of genomic events labelling a forest node, applied in array order */
enum {
  PCS_EV_SID = 0,      /* SNV / indel (RACES::Mutations::SID) placed on `allele` */
  PCS_EV_CNA_AMP = 1,  /* amplification: copy [pos,pos+len) of `allele` into new allele `dest` */
  PCS_EV_CNA_DEL = 2,  /* deletion of [pos,pos+len) from `allele` */
  PCS_EV_WGD = 3       /* whole genome doubling: every allele of every chromosome is copied */
};

/* Mutation::Nature, in the order of the reference's class strings
 * (src/sequencing.cpp:333-335): "driver", "passenger", "germinal", "preneoplastic" */
enum {
  PCS_NATURE_DRIVER = 0,
  PCS_NATURE_PASSENGER = 1,
  PCS_NATURE_GERMINAL = 2,
  PCS_NATURE_PRENEOPLASTIC = 3
};

/* sequencer models (src/seq_simulation.cpp:386-428) */
enum {
  PCS_SEQ_ERRORLESS = 0,        /* ErrorlessIlluminaSequencer or sequencer=NULL */
  PCS_SEQ_BASIC_CONSTANT = 1,   /* BasicSequencer<ConstantQualityScoreModel>    */
  PCS_SEQ_BASIC_RANDOM = 2      /* BasicSequencer<QualityScoreModel>            */
};

/*
 * Flat description of a phylogenetic forest whose nodes are labelled by the
 * genomic events arising in them.  All pointers are HOST memory owned by the
 * caller; the library never keeps them after pcs_forest_upload() returns.
 *
 * Mutation table: the distinct SIDs (chr, pos, ref, alt) of the forest and of
 * the germline, sorted by (chr, pos) with ties in the caller's (ref, alt) order
 * -- i.e. std::map<SID, ...> order (src/seq_simulation.cpp:148,176).  The row
 * index in this table is the row index of every output table.
 */
typedef struct pcs_forest_desc {
  /* chromosomes, in ChromosomeId order */
  uint32_t n_chr;
  const uint32_t* chr_len;        /* [n_chr] length in bp                                     */
  const uint8_t*  chr_n_alleles;  /* [n_chr] alleles in the germline genome (2, or 1 for X/Y) */

  /* cell tree: parent index < child index; -1 marks a root */
  uint32_t n_nodes;
  const int32_t* node_parent;     /* [n_nodes] */

  /* sampled cells (forest leaves) and the sample each comes from */
  uint32_t n_samples;
  uint32_t n_leaves;
  const uint32_t* leaf_node;      /* [n_leaves] node index; the node must have no children */
  const uint32_t* leaf_sample;    /* [n_leaves] in [0, n_samples)                          */

  /* events, CSR by node, in application order inside a node; pre-neoplastic
   * SIDs must come first in a root's list */
  uint64_t n_events;
  const uint64_t* node_event_off; /* [n_nodes+1] */
  const uint8_t*  ev_kind;        /* [n_events] PCS_EV_*                                   */
  const uint16_t* ev_chr;         /* [n_events] chromosome index (ignored for WGD)         */
  const uint32_t* ev_pos;         /* [n_events] CNA first position (SID: ignored)          */
  const uint32_t* ev_len;         /* [n_events] CNA length (SID: ignored)                  */
  const uint16_t* ev_allele;      /* [n_events] SID/DEL: target allele id; AMP: source     */
  const uint16_t* ev_dest;        /* [n_events] AMP: id of the new allele                  */
  const uint32_t* ev_mut;         /* [n_events] SID: row in the mutation table             */
  const uint8_t*  ev_nature;      /* [n_events] PCS_NATURE_*                               */

  /* mutation table */
  uint32_t n_mut;
  const uint16_t* mut_chr;        /* [n_mut] */
  const uint32_t* mut_pos;        /* [n_mut] 1-based */
  const uint8_t*  mut_ref_len;    /* [n_mut] length of `ref` (>= 1) */
  const uint8_t*  mut_alt_len;    /* [n_mut] length of `alt` (>= 1) */

  /* germline SIDs: row + bit mask of the germline alleles carrying it */
  uint64_t n_germline;
  const uint32_t* germ_mut;         /* [n_germline] */
  const uint8_t*  germ_allele_mask; /* [n_germline] bit a set: allele a carries the SID */
} pcs_forest_desc;

/*
 * The same forest as EXPLICIT PER-CELL GENOMES -- what the seam of the reference really hands over:
 * forest.get_sample_mutations_list() (per sample a list of CellGenomeMutations, src/seq_simulation.cpp:566) and
 * forest.get_normal_sample("normal_sample", true) (src/seq_simulation.cpp:572), each genome walked as
 * chromosome -> allele -> fragment -> SID (src/phylogenetic_forest.cpp:279-290, 332-335).  No tree is needed: the
 * library recovers the haplotype intervals from the genomes' contents.  Cells 0 .. n_cells-1 are the sampled
 * tumour cells; cells n_cells .. n_cells+n_normal_preneo-1 are the normal cells that carry the germline plus the
 * pre-neoplastic SIDs (needed only for preneoplastic_in_normal; the plain normal cell is the germline itself).
 * Alleles are listed in any order; an allele's fragments sorted by position; its SIDs in any order.
 */
typedef struct pcs_cell_genomes_desc {
  uint32_t n_chr;
  const uint32_t* chr_len;          /* [n_chr] */
  const uint8_t*  chr_n_alleles;    /* [n_chr] alleles of the germline genome */
  uint32_t n_samples;
  uint32_t n_cells;
  const uint32_t* cell_sample;      /* [n_cells] */
  uint32_t n_normal_preneo;
  uint64_t n_alleles;
  const uint32_t* allele_cell;      /* [n_alleles] */
  const uint16_t* allele_chr;       /* [n_alleles] */
  const uint16_t* allele_id;        /* [n_alleles] AlleleId inside the cell */
  const uint8_t*  allele_origin;    /* [n_alleles] germline allele it descends from (the source chain of its CNAs) */
  const uint64_t* allele_frag_off;  /* [n_alleles+1] */
  const uint32_t* frag_begin;       /* inclusive, 1-based */
  const uint32_t* frag_end;
  const uint64_t* allele_sid_off;   /* [n_alleles+1] */
  const uint32_t* sid_row;          /* somatic SIDs the allele carries: rows of the mutation table */
  /* mutation table and germline SIDs: as in pcs_forest_desc */
  uint32_t n_mut;
  const uint16_t* mut_chr;
  const uint32_t* mut_pos;
  const uint8_t*  mut_ref_len;
  const uint8_t*  mut_alt_len;
  uint64_t n_germline;
  const uint32_t* germ_mut;
  const uint8_t*  germ_allele_mask;
} pcs_cell_genomes_desc;

/* arguments of one simulate_seq()/simulate_normal_seq() call
 * (src/seq_simulation.hpp:28-54; defaults src/sequencing.cpp:193-209) */
typedef struct pcs_seq_params {
  int32_t  seed;                  /* resolved seed (utility.hpp:41-64)                     */
  double   coverage;
  double   purity;                /* in [0,1]; ignored when normal_only                    */
  uint32_t read_size;
  uint32_t insert_size_mean;      /* 0: single-end; >0: paired-end                         */
  uint32_t insert_size_stddev;
  uint32_t sequencer;             /* PCS_SEQ_*                                             */
  double   error_rate;
  uint8_t  with_normal_sample;    /* append "normal_sample" as last output sample          */
  uint8_t  preneoplastic_in_normal;
  uint8_t  normal_only;           /* simulate_normal_seq(): only the normal sample         */
  uint8_t  reserved0;
  const uint8_t* chr_mask;        /* [n_chr] 1 = sequence the chromosome; NULL = all       */
  /* work sharding (one process per GPU): this call handles tiles of
   * shard `shard_rank` out of `shard_count`; counts of all shards add up to
   * the single-shard result bit for bit */
  uint32_t shard_rank;
  uint32_t shard_count;           /* 0 is read as 1 */
} pcs_seq_params;

/* one injected read placement (parity mode) */
typedef struct pcs_read_placement {
  uint32_t cell;     /* leaf index; for normal cells: root ordinal (PRENEO) or 0 (PLAIN) */
  uint32_t start;    /* 1-based reference position of the first base of the read        */
  uint16_t chr;
  uint16_t allele;   /* allele id inside the cell                                       */
  uint16_t sample;   /* output sample index                                             */
  uint16_t flags;    /* PCS_PLACE_*                                                     */
} pcs_read_placement;

enum {
  PCS_PLACE_TUMOUR = 0,
  PCS_PLACE_NORMAL_PLAIN = 1,   /* contaminant/normal cell carrying the germline only */
  PCS_PLACE_NORMAL_PRENEO = 2   /* normal cell carrying germline + pre-neoplastic SIDs */
};

#define PCS_ERRMASK_WORDS 8     /* 256 read offsets; bit i set = base i is a sequencing error */

typedef struct pcs_plan_info {
  uint32_t n_out_samples;   /* groups (+1 if with_normal_sample), or 1 if normal_only */
  uint32_t n_mut;           /* rows of the output tables                              */
  uint32_t n_loci;
  uint64_t n_tiles;         /* tiles of THIS shard                                    */
  uint64_t n_tiles_total;
  uint64_t n_templates;     /* read templates of THIS shard                           */
  uint64_t n_templates_total;
  uint32_t reads_per_template; /* 1 or 2 */
  uint32_t read_size;
} pcs_plan_info;

typedef struct pcs_run_stats {
  double   kernel_ms;       /* device time of the sampler kernel(s), CUDA events on the launch stream */
  double   total_ms;        /* zero + sample + finalise (+ copies when host output)                   */
  uint64_t kernel_launches; /* kernels launched by this call                                          */
  uint64_t n_templates;
  uint64_t n_reads;         /* reads that were placed (templates not rejected x mates)                */
  uint64_t sum_depth;       /* sum over samples and loci of depth  (k_bar   = sum_depth / n_reads)    */
  uint64_t sum_occurrences; /* sum over samples and rows of occ.   (k_alt   = sum_occ   / n_reads)    */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
} pcs_run_stats;

typedef struct pcs_ctx pcs_ctx;
typedef struct pcs_forest pcs_forest;
typedef struct pcs_plan pcs_plan;

/* flags of pcs_plan_run */
enum {
  PCS_RUN_HOST_OUTPUT = 0,    /* occ/cov are host pointers  */
  PCS_RUN_DEVICE_OUTPUT = 1,  /* occ/cov are device pointers on the context's device */
  PCS_RUN_NO_CHECKSUMS = 2,   /* skip the two table checksums (sum_depth / sum_occurrences stay 0) */
  PCS_RUN_ASYNC = 4           /* with DEVICE_OUTPUT: return once the kernels are queued; stats are not filled,
                                 the plan's counters accumulate until pcs_plan_counters() reads them */
};

int pcs_abi_version(void);
const char* pcs_last_error(void);

/* one context per GPU; `stream` is a cudaStream_t (NULL = a stream owned by the context).
 * Lifetime: free every forest (pcs_forest_free) and plan (pcs_plan_free) made on a context before
 * pcs_destroy; a forest hands the pinned block its tables were built in back to its context. */
int pcs_create(pcs_ctx** ctx, int device_id, void* stream);
int pcs_destroy(pcs_ctx* ctx);
int pcs_device_name(pcs_ctx* ctx, char* buf, size_t buf_len);

/* flatten the forest (haplotype numbering, loci, fragment sets) and copy it to HBM */
int pcs_forest_upload(pcs_ctx* ctx, const pcs_forest_desc* desc, pcs_forest** forest);
/* the same from explicit per-cell genomes (pcs_cell_genomes_desc): identical tables for identical reads */
int pcs_forest_upload_genomes(pcs_ctx* ctx, const pcs_cell_genomes_desc* genomes, pcs_forest** forest);
int pcs_forest_free(pcs_forest* forest);
/* repartition the sampled cells into `n_groups` output samples (FACS labelling);
 * leaf_group == NULL restores the forest's own samples */
int pcs_forest_set_groups(pcs_forest* forest, const uint32_t* leaf_group, uint32_t n_groups);
/* sizes of the flattened view: out[0]=n_loci out[1]=n_instances out[2]=n_haplotypes
 * out[3]=n_fragment_sets out[4]=n_pieces out[5]=device bytes */
int pcs_forest_info(const pcs_forest* forest, uint64_t out[6]);
/* the forest's instance table as it lies on the device: inst[4 * n_instances] = {first haplotype, haplotypes, row,
 * ref_len | alt_len << 8} sorted by row, locus_inst_off[n_loci + 1].  An uploaded forest builds the table on the
 * device from the germline masks (PCS_DEVICE_INSTANCES=0: on the host); for tests -- pcs_flat_instances is the
 * host-built twin.  Host pointers. */
int pcs_forest_instances(pcs_forest* forest, uint32_t* inst, uint32_t* locus_inst_off);

/* tile grid, per-tile template counts (host multinomial), sampling tables -> HBM */
int pcs_plan_create(pcs_forest* forest, const pcs_seq_params* params, pcs_plan** plan);
int pcs_plan_info_get(const pcs_plan* plan, pcs_plan_info* info);
int pcs_plan_free(pcs_plan* plan);

/* run the sampler: occurrences[s*n_mut + row], coverage[s*n_mut + row]
 * (uint32, n_out_samples x n_mut each).  stats may be NULL. */
int pcs_plan_run(pcs_plan* plan, int flags, uint32_t* occurrences, uint32_t* coverage,
                 pcs_run_stats* stats);

/* ---- multi-GPU: accumulate every shard straight into ONE pair of tables over NVLink peer memory ----
 * The sampler's flush is a red.global.add per touched counter, so the tables may live on another GPU:
 * the reduction of src/seq_simulation.cpp's per-sample statistics is fused into the kernel epilogue.
 *   owner:  pcs_shared_alloc (+ ipc handle for other processes), pcs_memset_u32, ..., pcs_plan_finalize
 *   others: pcs_shared_open (other process) or pcs_enable_peer (other device of this process),
 *           pcs_plan_accumulate
 * depth is [n_out_samples][n_loci], occurrences [n_out_samples][n_mut] (pcs_plan_info). */
int pcs_shared_alloc(pcs_ctx* ctx, size_t bytes, void** dev_ptr, unsigned char ipc_handle[64]);
int pcs_shared_free(pcs_ctx* ctx, void* dev_ptr);
int pcs_shared_open(pcs_ctx* ctx, const unsigned char ipc_handle[64], void** dev_ptr);
int pcs_shared_close(pcs_ctx* ctx, void* dev_ptr);
int pcs_enable_peer(pcs_ctx* ctx, int peer_device);
int pcs_memset_u32(pcs_ctx* ctx, uint32_t* dev_ptr, size_t count);   /* async on the context's stream */
int pcs_memset_u32_stream(pcs_ctx* ctx, uint32_t* dev_ptr, size_t count, void* stream); /* async on a cudaStream_t */
int pcs_memcpy_d2h(pcs_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes); /* synchronous */
/* run this shard's sampler adding into the given tables: no zeroing, no finalize.  stats == NULL: asynchronous --
 * the call returns once the kernels are queued on the context's stream and the plan's counters accumulate */
int pcs_plan_accumulate(pcs_plan* plan, uint32_t* depth, uint32_t* occurrences, pcs_run_stats* stats);
/* coverage[s][row] = depth[s][locus(row)] (device pointers) + table checksums in stats; stats == NULL:
 * asynchronous, no checksums */
int pcs_plan_finalize(pcs_plan* plan, const uint32_t* depth, const uint32_t* occurrences, uint32_t* coverage,
                      pcs_run_stats* stats);
/* the coverage gather alone, queued on the given cudaStream_t (not the context's): the owner of shared tables
 * runs it beside the next step's sampler instead of in front of it */
int pcs_plan_finalize_stream(pcs_plan* plan, const uint32_t* depth, uint32_t* coverage, void* stream);
/* wait for the plan's queued work, read and reset its counters: n_reads / checksums accumulated by the asynchronous
 * calls since the last read, kernel_ms of the last sampler launch */
int pcs_plan_counters(pcs_plan* plan, pcs_run_stats* stats);

/* ---- one process, several GPUs (what the single-threaded R session needs) ----
 * pcs_forest_replicate: a second device copy of an uploaded forest, sharing its flattened host view.
 * pcs_simulate_multi:   forests[i] lives on device i (replicas of forests[0]); shard i of n runs on it from its
 *                       own host thread, all samplers flush into forests[0]'s device, host tables come back. */
int pcs_forest_replicate(pcs_forest* src, pcs_ctx* ctx, pcs_forest** replica);
int pcs_simulate_multi(pcs_forest* const* forests, uint32_t n, const pcs_seq_params* params,
                       uint32_t* occurrences, uint32_t* coverage, pcs_run_stats* stats);

/* ---- SAM output (write_SAM = TRUE; src/seq_simulation.cpp:540-564, vignettes/sequencing.Rmd:137-177, 272-309) ----
 * Reads are materialised on the GPU (bases from the reference + carried SIDs, CIGAR, qualities, errors) with the
 * Philox counters of the counting kernels: the tables and the SAM files of one plan describe the same reads. */
int pcs_forest_set_reference(pcs_forest* forest, uint32_t chr, const char* bases, uint64_t len);
int pcs_forest_load_fasta(pcs_forest* forest, const char* path, const char* const* chr_names, uint32_t* n_loaded);
int pcs_forest_set_alt(pcs_forest* forest, const uint32_t* alt_off /*[n_mut+1]*/, const char* alt_bytes);

typedef struct pcs_sam_options {
  const char* output_dir;
  const char* filename_prefix;        /* "chr_"  */
  const char* template_name_prefix;   /* "r"     */
  const char* const* chr_names;       /* [n_chr] */
  const char* const* sample_names;    /* [n_out_samples]: read group ids */
  uint8_t update;                     /* 0: the directory must not exist (Mode::CREATE); 1: Mode::UPDATE */
} pcs_sam_options;

/* one file <output_dir>/<filename_prefix><chr>.sam per sequenced chromosome, every sample a read group;
 * in update mode an existing file name gets the first free suffix _<n> */
int pcs_plan_write_sam(pcs_plan* plan, const pcs_sam_options* options, uint64_t* n_reads_written);

#define PCS_MAX_CIGAR 16
/* parity/debug: the reads of the plan as binary records.  seq/qual: [cap][read_size]; cigar: [cap][PCS_MAX_CIGAR]
 * words (length << 4 | op, op 0 M 1 I 2 D); lengths: bases written per read */
int pcs_plan_materialize(pcs_plan* plan, uint64_t cap, pcs_read_placement* placements, uint32_t* err_masks,
                         uint8_t* seq, uint8_t* qual, uint32_t* cigar, uint32_t* n_cigar, uint32_t* lengths,
                         uint64_t* n_out);

/* ---- binned depth tracks (the depth behind R/plot_genome_wide_mutations.R:75-108 between the
 * mutations as well): track[s * n_bins + chr_bin_off[chr] + (pos / bin_bp)] = reference bases the reads of sample s
 * lay on that bin -- the reads of THIS plan (same Philox counters as pcs_plan_run: the track and the tables describe
 * the same reads; a read counts the R bases from its start, in the frame it starts in).  bin_bp: a power of two.
 * chr_bin_off: [n_chr + 1] (filled).  track == NULL: only chr_bin_off and *n_bins are filled (to size the buffer of
 * n_out_samples * n_bins uint32). */
int pcs_plan_coverage_track(pcs_plan* plan, uint32_t bin_bp, uint64_t* chr_bin_off, uint32_t* track,
                            uint64_t cap_bins, uint64_t* n_bins);

/* debug/parity: re-run the plan emitting every placed read as a placement
 * record (+ its error mask when the sequencer has errors) instead of counting.
 * Host buffers of capacity `cap` records; *n_out receives the number written. */
int pcs_plan_trace(pcs_plan* plan, pcs_read_placement* placements, uint32_t* err_masks,
                   uint64_t cap, uint64_t* n_out);

/* plan + run + free with host outputs: the call the Rcpp shim makes */
int pcs_simulate(pcs_forest* forest, const pcs_seq_params* params,
                 uint32_t* occurrences, uint32_t* coverage, pcs_run_stats* stats);

/* parity mode: count a caller-supplied list of read placements.
 * err_masks: NULL or n * PCS_ERRMASK_WORDS words.  Host pointers. */
int pcs_count_injected(pcs_forest* forest, uint32_t n_out_samples, uint32_t read_size,
                       const pcs_read_placement* placements, const uint32_t* err_masks,
                       uint64_t n, uint32_t* occurrences, uint32_t* coverage,
                       pcs_run_stats* stats);

/* rows with occurrences > 0 in at least one sample (get_active_mutations, src/seq_simulation.cpp:142-167);
 * with include_non_sequenced != 0 also the rows some sequenced cell carries (params says which cells were
 * sequenced; NULL = all).  rows_out has capacity n_mut; *n_rows receives the count. */
int pcs_active_rows(pcs_forest* forest, const uint32_t* occurrences, uint32_t n_out_samples,
                    int include_non_sequenced, const pcs_seq_params* params, uint32_t* rows_out, uint32_t* n_rows);

/* ---- result assembly on the device (get_result_dataframe / get_active_mutations / add_sample_statistics,
 * src/seq_simulation.cpp:92-181): the rows of the data frame -- occurrences > 0 in some sample, or carried by a
 * sequenced cell with include_non_sequenced -- compacted in row order into one column per sample:
 * occurrences (int), coverage (int), VAF (double, occurrences / coverage; 0 where coverage is 0, :129-131).
 * Only these columns cross the link.  pcs_simulate_result = plan + sample + assemble (the call the Rcpp shim
 * makes); pcs_plan_result assembles the tables the last pcs_plan_run of the plan left on the device.
 * pcs_result_fetch copies into caller-owned columns (R vectors): rows[n_rows], and per sample s the pointers
 * occ_cols[s] / cov_cols[s] / vaf_cols[s] to n_rows elements each (any array or pointer may be NULL = skip). */
typedef struct pcs_result pcs_result;
int pcs_simulate_result(pcs_forest* forest, const pcs_seq_params* params, int include_non_sequenced, int with_vaf,
                        pcs_result** result, pcs_run_stats* stats);
int pcs_plan_result(pcs_plan* plan, int include_non_sequenced, const pcs_seq_params* params, int with_vaf,
                    pcs_result** result);
int pcs_result_info(const pcs_result* result, uint32_t* n_rows, uint32_t* n_samples, int* has_vaf);
int pcs_result_fetch(pcs_result* result, uint32_t* rows, int32_t* const* occ_cols, int32_t* const* cov_cols,
                     double* const* vaf_cols, uint64_t* d2h_bytes);
int pcs_result_free(pcs_result* result);

/* ---- host-side column builders for the annotation columns of the data frame (chr, chr_pos, ref, alt, causes,
 * classes; src/seq_simulation.cpp:52-90), parallel over the host cores; no GPU involved.
 * pcs_host_gather: dst[i] = src[rows[i]] for elements of 1, 2, 4 or 8 bytes.
 * pcs_host_string_column: the strings table[codes[rows[i]]] as an Arrow large_string column -- offsets[n + 1]
 * (int64), the concatenated bytes in `data` (capacity data_cap; NULL: only offsets, *data_len and the validity
 * bits are produced, so that the caller can size `data` and call again), and, when `validity` is not NULL, one
 * bit per row (LSB first, ceil(n / 8) bytes; 0 = the table entry is NULL = NA). */
int pcs_host_gather(const uint32_t* rows, uint64_t n, const void* src, uint32_t elem_bytes, void* dst);
int pcs_host_string_column(const uint32_t* rows, uint64_t n, const uint16_t* codes, const char* const* table,
                           uint32_t n_table, int64_t* offsets, char* data, uint64_t data_cap, uint8_t* validity,
                           uint64_t* data_len, uint64_t* null_count);

/* ---- host-only introspection of the flattened view (no GPU needed).  Used by the
 * CPU test-suite to check the haplotype-interval view against explicit per-cell
 * genomes and the planner's shard partition; not needed by the Rcpp shim. ---- */
typedef struct pcs_flat pcs_flat;
int pcs_flat_create(const pcs_forest_desc* desc, pcs_flat** flat);
int pcs_flat_create_genomes(const pcs_cell_genomes_desc* genomes, pcs_flat** flat);
int pcs_flat_free(pcs_flat* flat);
int pcs_flat_set_groups(pcs_flat* flat, const uint32_t* leaf_group, uint32_t n_groups);
int pcs_flat_info(const pcs_flat* flat, uint64_t out[6]);
/* haplotypes (alleles) of one cell on one chromosome; kind: PCS_PLACE_* */
int pcs_flat_cell_haps(pcs_flat* flat, uint32_t kind, uint32_t cell, uint32_t chr, uint32_t cap,
                       uint16_t* allele, uint32_t* hap, uint32_t* fragset, uint32_t* n);
int pcs_flat_fragset(const pcs_flat* flat, uint32_t fragset, uint32_t cap, uint32_t* begin,
                     uint32_t* end, uint32_t* n);
/* mutation rows carried by haplotype `hap` of chromosome `chr` (germline included) */
int pcs_flat_hap_rows(const pcs_flat* flat, uint32_t chr, uint32_t hap, uint32_t cap, uint32_t* rows,
                      uint32_t* n);
/* the haplotypes of chromosome(fragset) the sampler draws from for output group `group` and fragment set
 * `fragset`: group < n_groups = tumour cells of that group, n_groups = normal cells (germline only),
 * n_groups + 1 = normal cells with the pre-neoplastic SIDs.  *offset: position of the list in the device's
 * hap_list.  Lists are in increasing haplotype order and tile hap_list in (group, fragset) order. */
int pcs_flat_group_list(const pcs_flat* flat, uint32_t group, uint32_t fragset, uint32_t cap, uint32_t* haps,
                        uint32_t* offset, uint32_t* n);
int pcs_flat_instances(const pcs_flat* flat, uint32_t* inst, uint32_t* locus_inst_off);
/* host half of pcs_plan_create: tile grid of the shard named in params */
int pcs_flat_plan(const pcs_flat* flat, const pcs_seq_params* params, pcs_plan_info* info, uint64_t cap,
                  uint32_t* tile_id, uint32_t* tile_templates, uint32_t* tile_sample, uint32_t* tile_chr,
                  uint32_t* tile_begin, uint32_t* tile_len);

/* thinning of the plan's tiles (single-end reads; dev.hpp: Tile), in the order of pcs_flat_plan's heaviest-first
 * list is NOT guaranteed -- match by tile_id: thin (0/1), the number of start offsets from which a read can span
 * a locus or run past its fragment (u_len), the first offset of the tail zone, and how many of the tile's
 * templates start at such offsets (n_useful <= templates; == templates when the tile is not thinned) */
int pcs_flat_plan_thinning(const pcs_flat* flat, const pcs_seq_params* params, uint64_t cap, uint32_t* tile_id,
                           uint32_t* thin, uint32_t* u_len, uint32_t* tail_off, uint32_t* n_useful);
/* hap_list[offset, offset + n): the haplotype indices of a sampling list */
int pcs_flat_hap_list(const pcs_flat* flat, uint32_t offset, uint32_t n, uint32_t* haps);
/* the sampling entries of one tile of the (single-shard) plan: entry e owns the haplotype draw words
 * (thr[e-1], thr[e]] (from 0 for e = 0) and draws uniformly from the list of list_n[e] haplotypes at
 * hap_list[list_off[e]] -- every cell / allele of a class equiprobable (src/sequencing.cpp:155-163) */
int pcs_flat_tile_entries(const pcs_flat* flat, const pcs_seq_params* params, uint32_t tile_id, uint32_t cap,
                          uint32_t* thr, uint32_t* list_off, uint32_t* list_n, uint32_t* frag_end, uint32_t* n);
/* the haplotype draw of the sampler kernels evaluated on the host (same inline function, dev.hpp::exact_leaf):
 * haplotype index (inside the tile's chromosome) and entry index each 32-bit draw word u[i] selects */
int pcs_flat_draw(const pcs_flat* flat, const pcs_seq_params* params, uint32_t tile_id, uint64_t n_draws,
                  const uint32_t* u, uint32_t* hap, uint32_t* entry);
/* `count` draws of the planner's Binomial(n, p) sampler (csrc/plan_rng.hpp) from the stream of `seed`: the
 * templates of a (sample, chromosome) are split over its tiles with chains of these (a multinomial), as the
 * reference draws its reads' positions one by one (src/sequencing.cpp:155-163).  Host only; for tests. */
int pcs_host_binomial(uint32_t seed, uint64_t n, double p, uint64_t count, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* PCS_SEQ_H */
